#!/usr/bin/env python
"""Golden fixtures for the DE step, generated from the LIVE reference (run in the build container only):

    python tests/golden/make_golden_de.py

Runs the reference's own DERandOneEvolver / RandomKeyIndividual / DifferentialEvolutionSelector on a seeded
population and records the key matrices before and after, the random draws (re-derived with the oracle's
``draw_like_reference`` from the same seed and asserted to reproduce the reference's offspring bit for bit),
the decoded genomes and the selection outcome for a synthetic fitness vector."""
import os
import random
import sys

import numpy as np

REF = os.environ.get("TBLUP_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, REF)
if not hasattr(np, "asscalar"):
    np.asscalar = lambda a: a.item()

import tblup  # noqa: E402
from oracle import de_oracle as D  # noqa: E402


class Pop(list):
    generation = 0


def case(name, P, dim, length, seed, F, CR, clip, generations):
    random.seed(seed)
    np.random.seed(seed)
    pop = Pop(tblup.RandomKeyIndividual(length, dim) for _ in range(P))
    evolver = tblup.DERandOneEvolver(dim, CR, F, clip)
    selector = tblup.DifferentialEvolutionSelector()
    rec = {k: [] for k in ("keys", "child", "abc", "fixed", "mask", "F_used", "genomes", "child_genomes", "pfit", "cfit",
                           "take")}
    frng = np.random.default_rng(seed + 99)
    for ind in pop:
        ind.set_fitness(float(frng.random()))
    for gen in range(1, generations + 1):
        pop.generation = gen
        keys = np.stack([ind.get_internal_genome() for ind in pop])
        st_py, st_np = random.getstate(), np.random.get_state()
        children = evolver.evolve(pop)                                   # the reference draws here
        after_py, after_np = random.getstate(), np.random.get_state()
        random.setstate(st_py)
        np.random.set_state(st_np)                                       # replay the same draws with the oracle
        mi = D.mutation_intensity(gen, F)
        abc, fixed, masks, mine = [], [], [], []
        for i in range(P):
            a, b, c, f, u = D.draw_like_reference(P, dim, i)
            abc.append((a, b, c)); fixed.append(f); masks.append(u)
            mine.append(D.de_rand_one(keys, i, a, b, c, f, u, mi, CR, clip, dim))
        assert random.getstate() == after_py and all(np.array_equal(x, y) for x, y in zip(np.random.get_state()[1:2], after_np[1:2]))
        child_keys = np.stack([ch.get_internal_genome() for ch in children])
        assert np.array_equal(child_keys, np.stack(mine)), "oracle DE step differs from the reference"
        cfit = frng.random(P)
        cfit[frng.integers(0, P)] = np.nan                               # NaN never wins (selector.py:28)
        for ch, f in zip(children, cfit):
            ch.set_fitness(float(f))
        pfit = np.array([ind.fitness for ind in pop])
        new = selector.select(pop, children)
        take = np.array([n is ch for n, ch in zip(new, children)])
        assert np.array_equal(take, D.select(pfit, cfit))
        rec["keys"].append(keys); rec["child"].append(child_keys); rec["abc"].append(np.array(abc))
        rec["fixed"].append(np.array(fixed)); rec["mask"].append(np.stack(masks) < CR); rec["F_used"].append(mi)
        rec["genomes"].append(np.stack([ind.genome for ind in pop]))
        rec["child_genomes"].append(np.stack([ch.genome for ch in children]))
        for g_ref, kv in zip(rec["child_genomes"][-1], child_keys):
            assert np.array_equal(g_ref, D.decode(kv, length))
        rec["pfit"].append(pfit); rec["cfit"].append(cfit); rec["take"].append(take)
        pop = Pop(new)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), P=P, dim=dim, length=length, seed=seed, F=F, CR=CR,
                        clip=clip, **{k: np.stack(v) for k, v in rec.items()},
                        final_keys=np.stack([ind.get_internal_genome() for ind in pop]))
    print(name, "generations", generations, "replaced per generation", [int(t.sum()) for t in rec["take"]])


if __name__ == "__main__":
    case("de_small", P=12, dim=300, length=40, seed=5, F=0.5, CR=0.8, clip=False, generations=6)
    case("de_clip", P=9, dim=200, length=25, seed=8, F=0.5, CR=0.8, clip=True, generations=5)
