"""Device DE step against the reference's own DE (fixtures recorded from the live reference by
tests/golden/make_golden_de.py): fed the reference's random draws, the device reproduces the offspring keys bit for
bit, decodes the same SNP sets, and -- with the fitness it computes itself -- makes the selection the oracle makes."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import de_oracle as D
from oracle import gblup_oracle as O

pytestmark = pytest.mark.gpu


def _engine(dim, n=120, seed=3):
    from tblup_b200 import GblupEngine
    x, y = O.synth_genotypes(n, dim, h2=0.4, seed=seed)
    import random
    random.seed(seed)
    np.random.seed(seed)
    tr, va, te = O.ref_splits(n)
    eng = GblupEngine(x, y, perm=np.concatenate([tr, va, te]))
    eng.set_rowset(0, tr, va)
    return eng, x, y, tr, va


@pytest.mark.parametrize("name", ["de_small", "de_clip"])
def test_device_de_replays_reference_generations(name):
    from tblup_b200.de import DeviceDE
    g = load_golden(name)
    P, dim, length = int(g["P"]), int(g["dim"]), int(g["length"])
    CR, clip = float(g["CR"]), bool(g["clip"])
    eng, x, y, tr, va = _engine(dim)
    try:
        keys = g["keys"][0].copy()
        de = DeviceDE(eng, P, length, keys=keys)
        fit = de.evaluate(slots=[0], h2=0.4)
        want0 = np.array([O.exact_blup(D.decode(kv, length), tr, va, x, y, 0.4) for kv in keys])
        assert np.abs(fit - want0).max() < 1e-7
        cur_keys, cur_fit = keys, want0
        for gen in range(g["keys"].shape[0]):
            # the population the reference had is only the same as ours while selections agree; drive both from ours
            abc, fixed, mask = g["abc"][gen], g["fixed"][gen], g["mask"][gen]
            F = float(g["F_used"][gen])
            take = de.step(F, CR, slots=[0], h2=0.4, clip=clip, abc=abc, fixed=fixed, mask=mask)
            child = np.stack([D.de_rand_one(cur_keys, i, *abc[i], fixed[i], np.where(mask[i], 0.0, 1.0), F, CR, clip, dim)
                              for i in range(P)])
            assert np.array_equal(de.child_keys(), child)                       # bit for bit, no fma contraction
            got_sets = np.sort(de.last_genomes(), axis=1)
            want_sets = np.stack([np.sort(D.decode(kv, length)) for kv in child])
            if not clip:                                                         # clipping creates ties at the bounds
                assert np.array_equal(got_sets, want_sets)
            cfit = np.array([O.exact_blup(gs, tr, va, x, y, 0.4) for gs in got_sets])
            assert np.abs(de.child_fitness() - cfit).max() < 1e-7
            want_take = D.select(cur_fit, cfit)
            near_tie = np.abs(cfit - cur_fit) < 1e-6
            assert np.array_equal(take[~near_tie], want_take[~near_tie])
            cur_keys = np.where(take[:, None], child, cur_keys)
            cur_fit = np.where(take, cfit, cur_fit)
            assert np.array_equal(de.keys(), cur_keys)
            assert np.abs(de.fitness() - cur_fit).max() < 1e-7
            if gen == 0:                                                         # first generation: same start as the reference
                assert np.array_equal(g["keys"][0], keys) and np.array_equal(g["child"][0], child)
    finally:
        eng.close()


def test_decode_matches_argsort_top_k_with_ties_and_negatives():
    from tblup_b200.de import DeviceDE
    eng, *_ = _engine(257)
    try:
        rng = np.random.default_rng(0)
        keys = rng.normal(size=(6, 257))
        keys[1, :50] = 0.25                      # a block of ties straddling the threshold
        keys[2] = np.round(keys[2], 1)           # many ties
        keys[3] = -np.abs(keys[3])               # all negative
        keys[4, ::3] = -0.0
        for length in (1, 40, 256, 257):
            de = DeviceDE(eng, 6, length, keys=keys)
            for i in range(6):
                got = de.genome(i)
                assert len(np.unique(got)) == length and np.all(np.diff(got) > 0)
                thr = np.sort(keys[i])[257 - length]
                assert np.all(keys[i][got] >= thr)                               # nothing below the k-th largest key
                assert np.all(keys[i][np.setdiff1d(np.arange(257), got)] <= thr)
                if len(np.unique(keys[i])) == 257:
                    assert np.array_equal(got, np.sort(D.decode(keys[i], length)))
    finally:
        eng.close()


def test_device_driven_generations_improve_fitness():
    """Production mode: draws made on the device.  Greedy selection can only raise each individual's fitness."""
    from tblup_b200.de import DeviceDE
    eng, x, y, tr, va = _engine(600, n=150, seed=9)
    try:
        de = DeviceDE(eng, 16, 150, seed=11)
        k0 = de.keys()
        assert k0.min() >= 0.0 and k0.max() < 1.0 and abs(k0.mean() - 0.5) < 0.02
        f0 = de.evaluate()
        prev = f0.copy()
        for gen in range(1, 7):
            take = de.step(5.0 if gen % 5 == 0 else 0.5, 0.8, seed=gen)
            f = de.fitness()
            assert np.all(f >= prev - 1e-15) and np.array_equal(f > prev, take & (f > prev))
            prev = f
        assert prev.max() > f0.max() or np.any(prev > f0)
        best = int(np.argmax(prev))
        assert abs(O.exact_blup(de.genome(best), tr, va, x, y, 0.4) - prev[best]) < 1e-7
    finally:
        eng.close()
