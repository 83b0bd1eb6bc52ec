"""Knockout local search on the device against the live reference's KnockoutLocalSearch.search()
recorded in tests/golden/ko_*.npz (tblup/local.py:50-76) and against the exact oracle on every leave-one-out list."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import gblup_oracle as O
from oracle import knockout_oracle as K

pytestmark = pytest.mark.gpu


def _engine(g):
    from tblup_b200 import GblupEngine
    eng = GblupEngine(g["x"], g["y"], perm=np.concatenate([g["train"], g["valid"], g["test"]]))
    eng.set_rowset(0, g["train"], g["valid"])
    return eng


@pytest.mark.parametrize("name", ["ko_small", "ko_removed"])
@pytest.mark.parametrize("precision", ["mixed", "fp64"])
def test_knockout_reproduces_reference_decisions(name, precision):
    g = load_golden(name)
    with _engine(g) as eng:
        eng.set_precision(precision)
        for c in range(int(g["n_cases"])):
            genome = np.union1d(g["genome%d" % c], g["removed"]).astype(int)      # combine_with_removed (local.py:55)
            keep, best, evals, batches = eng.knockout(genome, float(g["start_fitness%d" % c]), slot=0, h2=float(g["h2"]))
            assert np.array_equal(genome[keep], g["kept%d" % c])                   # the same markers knocked out
            assert abs(best - float(g["best_fitness%d" % c])) < 1e-6
            assert evals == len(genome) and 1 <= batches <= len(genome)


def test_leave_one_out_scan_against_exact_oracle():
    g = load_golden("ko_small")
    genome = np.asarray(g["genome0"]).astype(int)          # k = 124 > n = 120: candidates take the gblup branch (123 > 120)
    dup = np.concatenate([genome[:40], genome[:3]])        # a list with duplicates, snp_blup branch
    with _engine(g) as eng:
        for lst in (genome, dup):
            got = eng.knockout_scan(lst, slot=0, h2=float(g["h2"]))
            want = K.leave_one_out(lst, g["train"], g["valid"], g["x"], g["y"], float(g["h2"]))
            assert got.shape == want.shape and np.abs(got - want).max() < 1e-6
        with pytest.raises(RuntimeError):
            eng.knockout_scan(genome[:1])


def test_knockout_crossing_the_branch_boundary_matches_sequential_oracle():
    """k just above n: every accepted drop moves the list towards len <= n, where blup() switches from gblup to
    snp_blup (tblup/evaluator.py:257); the batched search has to make the sequential loop's decisions across it."""
    g = load_golden("ko_small")
    x, y, h2 = g["x"], g["y"], float(g["h2"])
    rng = np.random.default_rng(5)
    genome = rng.choice(x.shape[1], size=x.shape[0] + 3, replace=False)
    start = O.exact_blup(genome, g["train"], g["valid"], x, y, h2)
    want_keep, want_best, _ = K.ref_knockout(genome, start, g["train"], g["valid"], x, y, h2)
    with _engine(g) as eng:
        keep, best, evals, batches = eng.knockout(genome, start, slot=0, h2=h2)
    assert np.array_equal(keep, want_keep) and abs(best - want_best) < 1e-6
    assert batches < evals          # speculation paid at least once


def test_local_search_class_through_the_installed_seam(tmp_path):
    """tblup_b200.local.KnockoutLocalSearch with the evaluator's contexts already closed (main.py:42-45 runs the search
    after the ``with evaluator`` block): same result as the reference's search."""
    import random
    from tblup_b200 import evaluator as ev
    from tblup_b200.local import KnockoutLocalSearch
    g = load_golden("ko_removed")
    np.save(tmp_path / "geno.npy", g["x"].astype(np.float64))
    np.save(tmp_path / "pheno.npy", g["y"])
    random.seed(int(g["seed"]))
    np.random.seed(int(g["seed"]))
    e = ev.BlupParallelEvaluator(str(tmp_path / "geno.npy"), str(tmp_path / "pheno.npy"), float(g["h2"]),
                                 snp_remover=ev.SNPRemovalHandler(10, 0.0, float(g["h2"]), False))
    assert list(e.training_indices) == list(g["train"])
    e.snp_remover.removed = g["removed"].astype(float)

    class Indv:
        def __init__(self, genome, fitness):
            self.genome, self.fitness = genome, fitness

    class Pop(list):
        evaluator = None

    pop = Pop([Indv(g["genome0"], float(g["start_fitness0"])), Indv(g["genome0"][:50], 0.01)])
    pop.evaluator = e
    search = KnockoutLocalSearch(pop)
    kept, best = search.search()
    assert np.array_equal(np.asarray(kept), g["kept0"]) and abs(best - float(g["best_fitness0"])) < 1e-6
    assert search.evaluations == len(np.union1d(g["genome0"], g["removed"]))
