#!/usr/bin/env python
"""BASELINE config 5: subset-size sweep k in {1k .. 50k} x population {100, 1 000, 4 000} on the headline data set
(5 000 x 50 000), one GPU per process (run under torchrun for more; every rank then evaluates P / world genomes).
Per point: evals/s with the genomes resident, per-stage milliseconds, and which stage dominates -- the k at which the
Gram overtakes the Cholesky (update + panel) is the GEMM- vs Cholesky-bound crossover SURVEY 8(d) asks for.
Parity: at every k one batch of 8 genomes is compared with the fp64 path and 2 genomes with the exact oracle.

    python scripts/sweep.py --out profiles/r02_sweep.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.json"))
    ap.add_argument("--ks", default="1000,5001,10000,25000,50000")
    ap.add_argument("--pops", default="100,1000,4000")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--budget_s", type=float, default=12.0, help="skip repeating a point whose one step takes longer")
    ap.add_argument("--opt", action="append", default=[], help="engine option name=value (A/B runs)")
    args = ap.parse_args()
    import torch
    from oracle import gblup_oracle as O
    from tblup_b200 import GblupEngine, MODE_AUTO, synth
    n, m, h2 = 5000, 50000, 0.4
    x, y = synth.synth_dataset(n, m, h2=h2, seed=0)
    tr, va, te = synth.split_indices(n, seed=0)
    eng = GblupEngine(x, y, perm=np.concatenate([tr, va, te]))
    eng.set_rowset(0, tr, va)
    for kv in args.opt:
        name, val = kv.split("=")
        eng.set_option(name, int(val))
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    points = []
    for k in [int(t) for t in args.ks.split(",")]:
        flat8, off8 = synth.random_genomes(8, m, k, seed=7 + k)
        eng.set_precision("mixed")
        f_def = eng.evaluate_packed(flat8, off8, [0], h2, MODE_AUTO)[:, 0]
        facts = {q: eng.info(q) for q in ("last_c16", "last_fp4", "last_fused_scale", "last_mixed")}
        eng.set_precision("fp64")
        f_64 = eng.evaluate_packed(flat8, off8, [0], h2, MODE_AUTO)[:, 0]
        eng.set_precision("mixed")
        exact = np.array([O.exact_blup(flat8[off8[i]:off8[i + 1]], tr, va, x, y, h2) for i in range(2)])
        parity = {"max_abs_default_vs_fp64": float(np.abs(f_def - f_64).max()),
                  "max_abs_default_vs_exact_oracle_2_genomes": float(np.abs(f_def[:2] - exact).max())}
        for P in [int(t) for t in args.pops.split(",")]:
            flat, off = synth.random_genomes(P, m, k, seed=100 + k + P)
            eng.stage(flat=flat, off=off)
            fit = torch.empty(P, dtype=torch.float64, device="cuda")
            t0 = time.perf_counter()
            eng.evaluate_staged([0], h2=h2, mode=MODE_AUTO, out_device_ptr=fit.data_ptr())      # warm-up
            torch.cuda.synchronize()
            one = time.perf_counter() - t0
            steps = args.steps if one < args.budget_s else 1
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps):
                eng.evaluate_staged([0], h2=h2, mode=MODE_AUTO, out_device_ptr=fit.data_ptr())
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            eng.set_option("profile", 1)
            eng.reset_counters()
            eng.evaluate_staged([0], h2=h2, mode=MODE_AUTO, out_device_ptr=fit.data_ptr())
            torch.cuda.synchronize()
            st = {s: v[0] for s, v in eng.stage_times().items()}
            eng.set_option("profile", 0)
            chol = st["chol_update"] + st["chol_panel"]
            pt = {"k": k, "pop": P, "evals_per_s": P / (ms * 1e-3), "ms_per_step": ms, "wave": eng.last_wave(),
                  "stage_ms": {s: round(v, 3) for s, v in st.items()}, "gram_over_cholesky": st["gram"] / chol if chol else None,
                  "gram_tops": 2.0 * k * (3200 * 3201 / 2 + 800 * 3200) * P / (st["gram"] * 1e-3) / 1e12 if st["gram"] else None,
                  "cross_products": "int16" if facts["last_c16"] else "int32", "parity_at_this_k": parity,
                  "fp64_fallbacks": eng.info("last_fallbacks")}
            points.append(pt)
            print(json.dumps(pt), flush=True)
    # crossover: interpolate gram/cholesky = 1 over k at the largest population
    pmax = max(p["pop"] for p in points)
    xs = [(p["k"], p["gram_over_cholesky"]) for p in points if p["pop"] == pmax and p["gram_over_cholesky"]]
    cross = None
    for (k0, r0), (k1, r1) in zip(xs, xs[1:]):
        if r0 < 1.0 <= r1:
            cross = k0 + (k1 - k0) * (1.0 - r0) / (r1 - r0)
    out = {"workload": "c5 sweep on 5000 x 50000 (n_t 3200, n_v 800), one GPU", "points": points,
           "gram_equals_cholesky_at_k": cross, "note": "crossover interpolated at pop %d; None = not crossed in range" % pmax}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(out, fh, indent=1)
    print("crossover k =", cross)
    eng.close()


if __name__ == "__main__":
    main()
