#!/usr/bin/env python
"""Start-up helpers at the headline shape (5 000 x 50 000): full-marker GRM + PCA split, top-SNPs marker scan -- GPU path
against the reference's host calls on the same data (tblup.utils.make_grm / sklearn f_regression), timed once each.

    python scripts/startup_bench.py --out gpurun_out/r02_startup.json
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
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_startup.json"))
    ap.add_argument("--n", type=int, default=5000)
    ap.add_argument("--m", type=int, default=50000)
    args = ap.parse_args()
    from oracle import gblup_oracle as O
    from tblup_b200 import GblupEngine, seeder, splitter, synth
    x, y = synth.synth_dataset(args.n, args.m, h2=0.4, seed=0)
    out = {"animals": args.n, "markers": args.m, "host_cores": len(os.sched_getaffinity(0))}
    t0 = time.perf_counter()
    g_gpu = splitter.full_grm(x)
    out["full_grm_gpu_s"] = time.perf_counter() - t0          # includes the ingest of the matrix
    # sklearn's PCA picks its randomized solver at this size and draws from the GLOBAL numpy RNG (random_state=None): the
    # reference's split is reproducible only through main.py's numpy.random.seed(args.seed); same state for both calls
    np.random.seed(0)
    t0 = time.perf_counter()
    split_gpu = splitter.pca_splitter(x)
    out["pca_splitter_gpu_s"] = time.perf_counter() - t0
    xf = x.astype(np.float64)
    t0 = time.perf_counter()
    g_ref = O.ref_make_grm(xf)                               # tblup/utils.py:7-18, all BLAS threads
    out["make_grm_host_s"] = time.perf_counter() - t0
    out["grm_max_abs_diff"] = float(np.abs(g_gpu - g_ref).max())
    np.random.seed(0)
    split_ref = O.ref_pca_split(g_ref)
    out["pca_split_identical"] = bool(split_gpu[0] == split_ref[0] and split_gpu[1] == split_ref[1])
    n_tr = int(0.64 * args.n)
    with GblupEngine(x, y) as eng:
        t0 = time.perf_counter()
        order, scores = seeder.sorted_indices(eng, y, n_tr)
        out["seeder_scan_gpu_s"] = time.perf_counter() - t0   # five folds, context already resident
    t0 = time.perf_counter()
    ref_scores = O.ref_seed_scores(xf, y, n_tr)
    out["seeder_scan_host_s"] = time.perf_counter() - t0
    ref_order = np.flip(np.argsort(ref_scores, axis=0), 0)
    out["seeder_scores_max_rel_diff"] = float(np.max(np.abs(scores - ref_scores) / np.maximum(np.abs(ref_scores), 1e-300)))
    out["seeder_top100_identical"] = bool(np.array_equal(order[:100], ref_order[:100]))
    print(json.dumps(out))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
