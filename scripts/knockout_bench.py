#!/usr/bin/env python
"""Knockout local search (tblup/local.py:50-76) at the headline shape: one genome of k = 5 001 markers on
5 000 x 50 000, GPU search (tb_knockout: speculative batches) against the reference's sequential loop, of which only
the first few steps are timed on the host (one step = one single-threaded blup(), ~5 s) and compared decision by decision.

    python scripts/knockout_bench.py --out gpurun_out/r02_knockout.json
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
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_knockout.json"))
    ap.add_argument("--k", type=int, default=5001)
    ap.add_argument("--cpu-steps", type=int, default=6)
    args = ap.parse_args()
    from oracle import gblup_oracle as O
    from oracle import knockout_oracle as K
    from tblup_b200 import GblupEngine, MODE_AUTO, synth
    n, m, h2 = 5000, 50000, 0.4
    x, y = synth.synth_dataset(n, m, h2=h2, seed=0)
    tr, va, te = synth.split_indices(n, seed=0)
    rng = np.random.default_rng(3)
    genome = np.sort(rng.choice(m, size=args.k, replace=False))
    with GblupEngine(x, y, perm=np.concatenate([tr, va, te])) as eng:
        eng.set_rowset(0, tr, va)
        start = float(eng.evaluate([genome], slots=[0], h2=h2, mode=MODE_AUTO)[0, 0])
        t0 = time.perf_counter()
        keep, best, evals, batches = eng.knockout(genome, start, slot=0, h2=h2)
        t_gpu = time.perf_counter() - t0
        launches = eng.launch_count()
    # the reference loop's first steps on the host (exact oracle = the same decisions, pinned in tests)
    xf = None
    mask = np.ones(len(genome), dtype=bool)
    best_cpu = start
    t0 = time.perf_counter()
    agree = True
    for i in range(args.cpu_steps):
        mask[i] = False
        f = O.exact_blup(genome[mask], tr, va, x, y, h2)
        if f > best_cpu:
            best_cpu = f
        else:
            mask[i] = True
        agree = agree and (mask[i] == keep[i])
    t_cpu = (time.perf_counter() - t0) / args.cpu_steps
    out = {"shape": "5000 x 50000, k = %d" % args.k, "start_fitness": start, "final_fitness": best,
           "markers_kept": int(keep.sum()), "evaluations_consumed": evals, "batched_passes": batches,
           "gpu_seconds": t_gpu, "gpu_launches": launches,
           "host_seconds_per_step_exact_oracle_all_cores": t_cpu,
           "host_seconds_extrapolated_sequential": t_cpu * len(genome),
           "first_%d_decisions_agree_with_sequential_oracle" % args.cpu_steps: bool(agree)}
    print(json.dumps(out))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
