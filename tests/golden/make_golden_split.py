#!/usr/bin/env python
"""Golden fixture for the start-up helpers built on the full-marker GRM / on per-marker statistics: the LIVE reference's
``pca_splitter`` (tblup/evaluator.py:641-663) and ``TopSNPsSeedStrategy.get_sorted_indices`` with the ``p_value`` metric
(tblup/seeder.py:144-160,202-210) on small seeded inputs.  Build container only:

    python tests/golden/make_golden_split.py
"""
import os
import random
import sys
import tempfile

import numpy as np

REF = os.environ.get("TBLUP_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, REF)

import tblup  # noqa: E402  (the live reference)
from tblup.seeder import TopSNPsSeedStrategy, p_value  # noqa: E402

from oracle.gblup_oracle import synth_genotypes  # noqa: E402


def main():
    out = {}
    for tag, (n, m, seed) in {"a": (120, 400, 21), "b": (203, 900, 22)}.items():
        x, y = synth_genotypes(n, m, h2=0.4, seed=seed)
        if tag == "b":
            x[:, 5] = 0                 # a monomorphic marker: f_regression gives NaN, which the seeder ranks FIRST
            x[:, 17] = x[:, 16]         # two identical markers: tied scores
        xf = x.astype(np.float64)
        out["x_" + tag], out["y_" + tag] = x, y
        for outl in (False, True):
            tr, te = tblup.pca_splitter(xf, outliers=outl)
            out["pca_train_%s_%d" % (tag, int(outl))] = np.array(tr)
            out["pca_test_%s_%d" % (tag, int(outl))] = np.array(te)
        out["grm_" + tag] = tblup.make_grm(xf)
        with tempfile.TemporaryDirectory() as tmp:
            g, p = os.path.join(tmp, "geno.npy"), os.path.join(tmp, "pheno.npy")
            np.save(g, xf)
            np.save(p, y)
            random.seed(seed)
            np.random.seed(seed)
            ev = tblup.BlupParallelEvaluator(g, p, 0.4, n_procs=1, snp_remover=None)
            strat = TopSNPsSeedStrategy(ev, p_value, g, p)
            out["seed_train_" + tag] = np.array(ev.training_indices)
            out["seed_order_" + tag] = np.asarray(strat.indices)
            # the summed metric itself, for a tolerance check independent of tie order
            from sklearn.model_selection import KFold
            scores = np.zeros(m)
            for train, _ in KFold(n_splits=5).split(ev.training_indices):
                scores += p_value(xf[train], y[train].ravel())
            out["seed_scores_" + tag] = scores
    np.savez_compressed(os.path.join(HERE, "split_seed.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    import warnings
    warnings.simplefilter("ignore")
    main()
