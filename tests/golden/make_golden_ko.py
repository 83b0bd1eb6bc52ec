#!/usr/bin/env python
"""Golden fixture for the knockout local search: the LIVE reference's KnockoutLocalSearch.search() (tblup/local.py:50-76)
on small seeded inputs.  Run in the build container only:

    python tests/golden/make_golden_ko.py

Stores inputs (dosages, phenotypes, splits, the best individual's genome and fitness) and the reference's outputs (the
returned genome, i.e. which markers were knocked out, the final fitness, and the fitness of every step's candidate)."""
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
from tblup.local import KnockoutLocalSearch  # noqa: E402

from oracle.gblup_oracle import synth_genotypes  # noqa: E402


class Indv:
    def __init__(self, genome, fitness):
        self.genome, self.fitness = genome, fitness

    def __len__(self):
        return len(self.genome)


class Pop(list):
    evaluator = None


def case(name, n, m, ks, seed, removed=()):
    x, y = synth_genotypes(n, m, h2=0.4, seed=seed)
    out = {"x": x, "y": y, "h2": 0.4, "seed": seed}
    with tempfile.TemporaryDirectory() as tmp:
        g, p = os.path.join(tmp, "geno.npy"), os.path.join(tmp, "pheno.npy")
        np.save(g, x.astype(np.float64))
        np.save(p, y)
        random.seed(seed)
        np.random.seed(seed)
        ev = tblup.BlupParallelEvaluator(g, p, 0.4, n_procs=1, snp_remover=tblup.SNPRemovalHandler(10, 0.0, 0.4, False))
        ev.snp_remover.removed = np.array(removed, dtype=float)
        out["train"], out["valid"], out["test"] = map(np.array, (ev.training_indices, ev.validation_indices, ev.testing_indices))
        out["removed"] = np.array(removed, dtype=np.int64)
        rng = np.random.default_rng(seed + 1)
        xf = x.astype(np.float64)
        for c, k in enumerate(ks):
            genomes = [rng.choice(m, size=k, replace=False) for _ in range(3)]
            fits = [tblup.BlupParallelEvaluator.blup(gg, ev.training_indices, ev.validation_indices, xf, y, 0.4) for gg in genomes]
            pop = Pop(Indv(gg, ff) for gg, ff in zip(genomes, fits))
            pop.evaluator = ev
            # record every candidate's fitness by wrapping the static method the search calls (no reference code changed)
            trace = []
            orig = ev.blup

            def spy(*a, **kw):
                f = orig(*a, **kw)
                trace.append(f)
                return f
            ev.blup = spy
            kept, best = KnockoutLocalSearch(pop).search()
            del ev.blup
            b = int(np.argmax(fits))
            out["genome%d" % c] = genomes[b]
            out["start_fitness%d" % c] = fits[b]
            out["kept%d" % c] = np.asarray(kept)
            out["best_fitness%d" % c] = best
            out["trace%d" % c] = np.array(trace)
    out["n_cases"] = len(ks)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: (v.shape if hasattr(v, "shape") and v.shape else v) for k, v in out.items() if k.startswith(("kept", "best", "start"))})


if __name__ == "__main__":
    # k straddles n (gblup at the start, snp_blup once enough markers are gone: tblup/evaluator.py:257), k << n, and a
    # run with previously removed markers united back in (combine_with_removed, local.py:55)
    case("ko_small", n=120, m=600, ks=[124, 60], seed=11)
    case("ko_removed", n=100, m=400, ks=[90], seed=12, removed=[3, 17, 250, 251])
