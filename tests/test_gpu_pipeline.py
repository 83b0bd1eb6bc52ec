"""GPU parity tests, stage by stage and end to end, through the C-ABI (tblup_b200.engine -> ctypes).

Bars (BASELINE.json north_star): uncentred Gram entries and the integer centring terms bit-exact against the
oracle; fitness within 1e-6 absolute of the reference values stored in tests/golden/*.npz.
"""
import numpy as np
import pytest

from conftest import load_golden, unpack
from oracle import gblup_oracle as O

pytestmark = pytest.mark.gpu

FIT_TOL = 1e-6   # north_star: "per-individual fitness agrees with the reference numpy path within 1e-6 absolute"


def _engine(x, y, train, valid, test=None, extra_sets=(), precision="mixed"):
    from tblup_b200 import GblupEngine
    rest = [] if test is None else list(test)
    perm = np.concatenate([np.asarray(train), np.asarray(valid), np.asarray(rest, dtype=np.int64)]).astype(np.int64)
    if perm.size != x.shape[0]:
        missing = np.setdiff1d(np.arange(x.shape[0]), perm)
        perm = np.concatenate([perm, missing])
    eng = GblupEngine(x, y, perm=perm)
    eng.set_precision(precision)
    eng.set_rowset(0, train, valid)
    for slot, (t, v) in enumerate(extra_sets, start=1):
        eng.set_rowset(slot, t, v)
    return eng, perm


@pytest.fixture(scope="module")
def mid():
    g = load_golden("fit_mid")
    eng, perm = _engine(g["x"], g["y"], g["train"], g["valid"], g["test"], precision="fp64")
    yield g, eng, perm
    eng.close()


@pytest.mark.parametrize("impl", ["simt", "tc"])
@pytest.mark.parametrize("k", [1, 31, 128, 700, 1500, -900])
def test_gram_bit_exact(mid, impl, k):
    g, eng, perm = mid
    rng = np.random.default_rng(abs(k) + 7)
    m = g["x"].shape[1]
    idx = rng.integers(0, m, size=-k) if k < 0 else rng.choice(m, size=k, replace=False)
    rows = 400
    got = eng.gram_debug(idx, rows, impl=impl)
    want = np.tril(O.exact_gram(g["x"], idx, perm[:rows]))
    assert got.dtype == np.int32
    assert np.array_equal(got.astype(np.int64), want)


def test_gram_partial_rows(mid):
    """Row counts that are not tile multiples, including a single 128-row tile and an odd count."""
    g, eng, perm = mid
    idx = np.arange(0, 1500, 3)
    for rows in (1, 127, 129, 257, 385):
        got = eng.gram_debug(idx, rows, impl="tc")
        assert np.array_equal(got.astype(np.int64), np.tril(O.exact_gram(g["x"], idx, perm[:rows])))


@pytest.mark.parametrize("mode", [O.MODE_GBLUP, O.MODE_SNPBLUP])
def test_stage_by_stage_against_oracle(mid, mode):
    from tblup_b200 import engine as E
    g, eng, perm = mid
    x, y, h2 = g["x"], g["y"], float(g["h2"])
    tr, va = g["train"], g["valid"]
    genomes = [gen for gen in unpack(g["genomes_flat"], g["genomes_off"])]
    job = 3
    fit_ref, d = O.exact_fitness(genomes[job], tr, va, x, y, h2, mode, detail=True)
    nt, nv = len(tr), len(va)
    api_mode = E.MODE_GBLUP if mode == O.MODE_GBLUP else E.MODE_SNPBLUP

    eng.set_option("stop_after", 4)          # after the scale stage: M = [A ; G_vt] before factorisation
    eng.evaluate(genomes, slots=[0], h2=h2, mode=api_mode)
    dims = eng.debug_dims(job)
    C = eng.debug_fetch(E.DBG_C, job)
    assert np.array_equal(np.tril(C[:nt + nv, :nt + nv]).astype(np.int64)[:, :nt], np.tril(d["C"])[:, :nt])
    s = eng.debug_fetch(E.DBG_S, job)
    assert np.array_equal(s[:nt + nv], d["s"])
    SQ = eng.debug_fetch(E.DBG_SQ, job)
    assert int(SQ[0]) == d["S"] and int(SQ[1]) == d["Q"]
    M = eng.debug_fetch(E.DBG_M, job)
    ntp = dims["ntp"]
    assert np.array_equal(np.tril(M[:nt, :nt]), np.tril(d["A"]))          # same integer numerators, one division
    g_vt = O.exact_grm_block(d["C"][nt:, :nt], d["s"][nt:], d["s"][:nt], d["S"], d["Q"], d["N"])
    assert np.array_equal(M[ntp:ntp + nv, :nt], g_vt)
    if ntp > nt:
        assert np.array_equal(np.tril(M[nt:ntp, :ntp]), np.tril(np.eye(ntp)[nt:ntp]))

    eng.set_option("stop_after", -1)
    fit = eng.evaluate(genomes, slots=[0], h2=h2, mode=api_mode)[:, 0]
    L = np.tril(eng.debug_fetch(E.DBG_M, job)[:nt, :nt])
    assert np.allclose(L @ L.T, d["A"], rtol=0, atol=1e-10 * np.abs(d["A"]).max())
    alpha = eng.debug_fetch(E.DBG_ALPHA, job)[:nt]
    assert np.allclose(alpha, d["alpha"], rtol=1e-9, atol=1e-12)
    pred = eng.debug_fetch(E.DBG_PRED, job)
    assert np.allclose(pred, d["pred"], rtol=1e-9, atol=1e-12)
    assert abs(fit[job] - fit_ref) < 1e-12


@pytest.mark.parametrize("precision", ["fp64", "mixed"])
@pytest.mark.parametrize("name", ["fit_small", "fit_offset", "fit_mid"])
def test_fitness_matches_reference(name, precision):
    """Every branch of blup() on the reference's own numbers: forced gblup, forced snp_blup, the k > n
    dispatch, the testing split (train+valid -> test) and each cross-validation fold."""
    from tblup_b200 import engine as E
    g = load_golden(name)
    x, y, h2 = g["x"], g["y"], float(g["h2"])
    tr, va, te = g["train"], g["valid"], g["test"]
    f_tr = unpack(g["fold_train_flat"], g["fold_train_off"])
    f_va = unpack(g["fold_valid_flat"], g["fold_valid_off"])
    sets = [(np.concatenate((tr, va)), te)] + list(zip(f_tr, f_va))
    eng, _ = _engine(x, y, tr, va, te, extra_sets=sets, precision=precision)
    try:
        genomes = unpack(g["genomes_flat"], g["genomes_off"])
        got_g = eng.evaluate(genomes, slots=[0], h2=h2, mode=E.MODE_GBLUP)[:, 0]
        assert eng.last_precision() == precision
        got_s = eng.evaluate(genomes, slots=[0], h2=h2, mode=E.MODE_SNPBLUP)[:, 0]
        got_all = eng.evaluate(genomes, slots=list(range(len(sets) + 1)), h2=h2, mode=E.MODE_AUTO)
        assert np.abs(got_g - g["ref_gblup"]).max() < FIT_TOL
        assert np.abs(got_s - g["ref_snp_blup"]).max() < FIT_TOL
        assert np.abs(got_all[:, 0] - g["ref_blup"]).max() < FIT_TOL
        assert np.abs(got_all[:, 1] - g["ref_blup_testing"]).max() < FIT_TOL
        assert np.abs(got_all[:, 2:] - g["ref_blup_folds"]).max() < FIT_TOL
    finally:
        eng.close()


@pytest.mark.parametrize("precision", ["fp64", "mixed"])
@pytest.mark.parametrize("name", ["traj_gblup", "traj_snpblup", "traj_intercv", "traj_intracv"])
def test_trajectory_replay(name, precision):
    """Fixed-seed DE runs of the reference (its own Population/evolver/selector): every batch it evaluated
    gets the same fitness (1e-6) and therefore the same parent-vs-child decisions and selected panel."""
    from tblup_b200 import engine as E
    g = load_golden(name)
    x, y, h2 = g["x"], g["y"], float(g["h2"])
    regressor = str(g["regressor"])
    tr, va, te = g["train"], g["valid"], g["test"]
    sets = []
    if "fold_train_flat" in g:
        sets = list(zip(unpack(g["fold_train_flat"], g["fold_train_off"]),
                        unpack(g["fold_valid_flat"], g["fold_valid_off"])))
    eng, _ = _engine(x, y, tr, va, te, extra_sets=sets, precision=precision)
    try:
        genomes = unpack(g["genomes_flat"], g["genomes_off"])
        pos = 0
        fpos = 0
        pop = int(g["pop"])
        current = np.full(pop, -np.inf)
        for b, size in enumerate(g["batch_sizes"]):
            batch = genomes[pos:pos + size]
            pos += size
            if regressor == "intracv_blup":
                got = eng.evaluate(batch, slots=list(range(1, len(sets) + 1)), h2=h2, mode=E.MODE_AUTO).mean(axis=1)
            elif regressor == "intercv_blup":
                got = eng.evaluate(batch, slots=[1 + int(g["split_of_batch"][b])], h2=h2, mode=E.MODE_AUTO)[:, 0]
            else:
                got = eng.evaluate(batch, slots=[0], h2=h2, mode=E.MODE_AUTO)[:, 0]
            ref = g["fitness"][fpos:fpos + size]
            where = g["positions"][fpos:fpos + size]
            fpos += size
            assert np.abs(got - ref).max() < FIT_TOL
            # the selector's decision (tblup/selector.py:28: child replaces parent iff strictly fitter)
            decide_ref = ref > current[where]
            decide_got = got > current[where]
            ties = np.abs(ref - current[where]) < FIT_TOL
            assert np.array_equal(decide_ref[~ties], decide_got[~ties])
            current[where] = np.where(decide_ref, ref, current[where])
            if b < len(g["pop_fitness"]):          # the replayed selection is the reference's own population
                assert np.allclose(current, g["pop_fitness"][b], rtol=0, atol=1e-12)
    finally:
        eng.close()


def test_mixed_precision_factor_and_refinement():
    """The TF32 tensor-core factor is only a preconditioner: L L^T matches A to TF32 accuracy, and the fp64
    refinement against the exact integer operator lands on the oracle's alpha / predictions / fitness."""
    from tblup_b200 import engine as E
    g = load_golden("fit_mid")
    x, y, h2 = g["x"], g["y"], float(g["h2"])
    tr, va = g["train"], g["valid"]
    eng, _ = _engine(x, y, tr, va, g["test"], precision="mixed")
    try:
        genomes = unpack(g["genomes_flat"], g["genomes_off"])
        fit = eng.evaluate(genomes, slots=[0], h2=h2, mode=E.MODE_GBLUP)[:, 0]
        assert eng.last_precision() == "mixed"
        nt = len(tr)
        for job in (0, 3, 5):
            ref, d = O.exact_fitness(genomes[job], tr, va, x, y, h2, O.MODE_GBLUP, detail=True)
            L = np.tril(eng.debug_fetch(E.DBG_L32, job)[:nt, :nt]).astype(np.float64)
            rel = np.abs(np.tril(L @ L.T - d["A"])).max() / np.abs(d["A"]).max()
            assert rel < 5e-3, rel                       # TF32: 10 mantissa bits
            assert rel > 1e-7                            # ... and it really is the low-precision factor
            sweeps = int(eng.debug_fetch(E.DBG_SWEEPS, job)[0]) & 255
            assert 1 <= sweeps <= 6, sweeps
            # refinement stops once the predicted remaining error is below 1e-8 of max|alpha|
            amax = np.abs(d["alpha"]).max()
            assert np.abs(eng.debug_fetch(E.DBG_ALPHA, job)[:nt] - d["alpha"]).max() < 2e-8 * amax
            assert np.abs(eng.debug_fetch(E.DBG_PRED, job) - d["pred"]).max() < 2e-8 * np.abs(d["pred"]).max() + 1e-9
            assert abs(fit[job] - ref) < 1e-7
    finally:
        eng.close()


def test_config1_shape_against_oracle():
    """BASELINE config 1 shape (1000 animals x 10000 markers): random k-subsets on both sides of k = n."""
    from tblup_b200 import engine as E
    import random
    x, y = O.synth_genotypes(1000, 10000, h2=0.4, seed=0)
    random.seed(0)
    np.random.seed(0)
    tr, va, te = O.ref_splits(1000)
    eng, _ = _engine(x, y, tr, va, te)
    try:
        rng = np.random.default_rng(1)
        genomes = [rng.choice(10000, size=k, replace=False) for k in (100, 999, 1000, 1001, 1500, 1500, 2500, 4000)]
        want = np.array([O.exact_blup(gm, tr, va, x, y, 0.4) for gm in genomes])
        for precision in ("fp64", "mixed"):
            eng.set_precision(precision)
            got = eng.evaluate(genomes, slots=[0], h2=0.4, mode=E.MODE_AUTO)[:, 0]
            assert eng.last_precision() == precision
            assert np.abs(got - want).max() < (1e-9 if precision == "fp64" else 1e-7), precision
        xf = x.astype(np.float64)
        for i in (0, 4):
            assert abs(got[i] - O.ref_blup(genomes[i].astype(int), tr, va, xf, y, 0.4)) < FIT_TOL
    finally:
        eng.close()


def test_waves_and_ragged_batches():
    """The wave scheduler must not change results: same batch with 1, 3 and all individuals per wave."""
    from tblup_b200 import engine as E
    g = load_golden("fit_small")
    eng, _ = _engine(g["x"], g["y"], g["train"], g["valid"], g["test"])
    try:
        genomes = unpack(g["genomes_flat"], g["genomes_off"])
        base = eng.evaluate(genomes, slots=[0], h2=float(g["h2"]), mode=E.MODE_AUTO)
        for mw in (1, 3, 5):
            eng.set_option("max_wave", mw)
            again = eng.evaluate(genomes, slots=[0], h2=float(g["h2"]), mode=E.MODE_AUTO)
            assert eng.last_wave() == mw
            assert np.array_equal(base, again)
    finally:
        eng.close()


def test_error_paths():
    from tblup_b200 import GblupEngine
    g = load_golden("fit_small")
    x = g["x"].copy()
    with pytest.raises(ValueError):
        GblupEngine(np.where(x == 2, 3, x), g["y"])
    eng, _ = _engine(g["x"], g["y"], g["train"], g["valid"], g["test"])
    try:
        with pytest.raises(IndexError):
            eng.evaluate([np.array([0, x.shape[1]])])
        with pytest.raises(RuntimeError):
            eng.evaluate([np.array([1, 2, 3])], slots=[5])          # undefined row set
        with pytest.raises(RuntimeError):
            eng.evaluate([np.array([], dtype=np.int64)])
        # negative indices wrap like numpy fancy indexing
        a = eng.evaluate([np.array([-1, 5, 9])])[0, 0]
        b = eng.evaluate([np.array([x.shape[1] - 1, 5, 9])])[0, 0]
        assert a == b
    finally:
        eng.close()


def test_int16_and_int32_cross_product_storage_agree():
    """Mixed precision stores the cross-products as int16 when 4 k <= 32 767 for every genome of the batch (exact:
    C_ab <= 4 k).  Same fitness and the same integers as the int32 layout, for contiguous and scattered row sets,
    and a batch with a genome beyond the bound falls back to int32."""
    from tblup_b200 import engine as E
    g = load_golden("fit_mid")
    x, y, h2 = g["x"], g["y"], float(g["h2"])
    tr, va, te = g["train"], g["valid"], g["test"]
    eng, perm = _engine(x, y, tr, va, te, extra_sets=[(np.concatenate([tr[40:], va[:10]]), np.concatenate([tr[:40], va[10:]]))])
    try:
        rng = np.random.default_rng(11)
        m = x.shape[1]
        genomes = [rng.choice(m, size=k, replace=False) for k in (3, 200, 1499)] + [rng.integers(0, m, size=8191)]
        for slots in ([0], [1], [0, 1]):
            eng.set_option("narrow_c", 1)
            a = eng.evaluate(genomes, slots=slots, h2=h2, mode=E.MODE_GBLUP)
            c16 = eng.debug_fetch(E.DBG_C, 0)
            eng.set_option("narrow_c", 0)
            b = eng.evaluate(genomes, slots=slots, h2=h2, mode=E.MODE_GBLUP)
            c32 = eng.debug_fetch(E.DBG_C, 0)
            if slots == [0]:     # training animals are universe positions 0 .. n_t-1: rows x training columns are formed
                nt = len(tr) + len(va)
                assert np.array_equal(np.tril(c16[:nt, :nt])[:, :len(tr)], np.tril(c32[:nt, :nt])[:, :len(tr)])
            assert np.abs(a - b).max() < 1e-9
        eng.set_option("narrow_c", 1)
        big = genomes + [rng.integers(0, m, size=8192)]          # 4 * 8192 > 32 767: the whole batch uses int32
        got = eng.evaluate(big, slots=[0], h2=h2, mode=E.MODE_GBLUP)[:, 0]
        want = np.array([O.exact_fitness(gen, tr, va, x, y, h2, O.MODE_GBLUP) for gen in big])
        assert np.abs(got - want).max() < FIT_TOL
    finally:
        eng.close()


def test_mixed_factor_across_several_panels():
    """n_t = 704 training animals: three 256-wide block columns, so the wide panel path (inverse of the 256 x 256
    diagonal block + one K = 256 GEMM for the rows below) and the fp16-operand outer updates both run.  The factor
    of record (fp16 copy) must reproduce A to TF32 accuracy and the refined solution the oracle's."""
    import random
    from tblup_b200 import engine as E
    n, m = 1100, 3000
    x, y = O.synth_genotypes(n, m, h2=0.4, seed=21)
    random.seed(21)
    np.random.seed(21)
    tr, va, te = O.ref_splits(n)
    assert len(tr) == 704
    eng, perm = _engine(x, y, tr, va, te)
    try:
        rng = np.random.default_rng(22)
        genomes = [rng.choice(m, size=k, replace=False) for k in (150, 1101, 2500)]
        h2 = 0.4
        for wide in (1, 0):
            eng.set_option("wide_panel", wide)
            fit = eng.evaluate(genomes, slots=[0], h2=h2, mode=E.MODE_GBLUP)[:, 0]
            for job, gen in enumerate(genomes):
                ref, d = O.exact_fitness(gen, tr, va, x, y, h2, O.MODE_GBLUP, detail=True)
                nt = len(tr)
                L = np.tril(eng.debug_fetch(E.DBG_L32, job)[:nt, :nt]).astype(np.float64)
                rel = np.abs(np.tril(L @ L.T - d["A"])).max() / np.abs(d["A"]).max()
                assert 1e-7 < rel < 5e-3, (wide, job, rel)
                assert (int(eng.debug_fetch(E.DBG_SWEEPS, job)[0]) & 255) <= 4
                assert abs(fit[job] - ref) < 1e-7, (wide, job, fit[job], ref)
    finally:
        eng.close()


def test_cross_validation_folds_with_aligned_holes():
    """k-fold row sets whose held-out fold is an 8-aligned contiguous run of the training animals take the
    'contiguous with one hole' kernels (vectorised mat-vec, held-out predictions from row + column parts).
    Against the oracle, and against the generic position-lookup kernels (int32 layout), for both branches."""
    import random
    from tblup_b200 import engine as E
    n, m = 1100, 3000
    x, y = O.synth_genotypes(n, m, h2=0.4, seed=31)
    random.seed(31)
    np.random.seed(31)
    tr, va, te = O.ref_splits(n)
    assert len(tr) == 704
    folds = O.ref_make_fold_indices(list(tr), 8)              # 8 folds of 88 animals
    eng, perm = _engine(x, y, tr, va, te, extra_sets=[(f[0], f[1]) for f in folds])
    try:
        rng = np.random.default_rng(32)
        genomes = [rng.choice(m, size=k, replace=False) for k in (90, 1101, 2000)]
        slots = list(range(1, 9))
        for mode, omode in ((E.MODE_GBLUP, O.MODE_GBLUP), (E.MODE_SNPBLUP, O.MODE_SNPBLUP)):
            eng.set_option("narrow_c", 1)
            got = eng.evaluate(genomes, slots=slots, h2=0.4, mode=mode)
            assert eng.info("last_c16") == 1 and eng.info("last_split") == 0      # one Gram per genome, hole kernels
            eng.set_option("narrow_c", 0)
            generic = eng.evaluate(genomes, slots=slots, h2=0.4, mode=mode)
            assert eng.info("last_split") == 1       # int32 layout has no hole kernels: one permuted panel per fold
            eng.set_option("perm_rows", 0)
            lookup = eng.evaluate(genomes, slots=slots, h2=0.4, mode=mode)            # position-lookup kernels
            assert eng.info("last_split") == 0 and np.abs(lookup - generic).max() < 1e-9
            eng.set_option("perm_rows", 1)
            eng.set_option("narrow_c", 1)
            want = np.array([[O.exact_fitness(g, f[0], f[1], x, y, 0.4, omode) for f in folds] for g in genomes])
            assert np.abs(got - want).max() < 1e-7
            assert np.abs(got - generic).max() < 1e-9
        # base split and folds in one call (mixed contiguous / hole row sets)
        both = eng.evaluate(genomes, slots=[0, 3], h2=0.4, mode=E.MODE_GBLUP)
        assert abs(both[1, 0] - O.exact_fitness(genomes[1], tr, va, x, y, 0.4, O.MODE_GBLUP)) < 1e-7
        assert abs(both[1, 1] - O.exact_fitness(genomes[1], folds[2][0], folds[2][1], x, y, 0.4, O.MODE_GBLUP)) < 1e-7
    finally:
        eng.close()


@pytest.mark.parametrize("h2", [0.02, 0.9, 0.99, 0.999, 0.9999])
def test_heritability_extremes_fall_back_to_fp64(h2):
    """lambda = (1 - h2) / h2 -> 0 makes G_tt + lambda I ill-conditioned (singular G when k < n_t): the low-precision
    factor breaks down or stops preconditioning.  Those jobs are evaluated again in fp64, so the fitness stays within
    the bar for every heritability the reference accepts."""
    from tblup_b200 import engine as E
    g = load_golden("fit_mid")
    x, y = g["x"], g["y"]
    tr, va, te = g["train"], g["valid"], g["test"]
    eng, perm = _engine(x, y, tr, va, te)
    try:
        rng = np.random.default_rng(1)
        m = x.shape[1]
        genomes = [rng.choice(m, size=k, replace=False) for k in (40, 255, 300, 401, 1500)]
        total_fallbacks = 0
        for mode, om in ((E.MODE_GBLUP, O.MODE_GBLUP), (E.MODE_SNPBLUP, O.MODE_SNPBLUP)):
            got = eng.evaluate(genomes, slots=[0], h2=h2, mode=mode)[:, 0]
            total_fallbacks += eng.info("last_fallbacks")
            want = np.array([O.exact_fitness(gen, tr, va, x, y, h2, om) for gen in genomes])
            # the fp64 solves of a matrix with cond ~ 1e6+ agree with the oracle's to ~cond * eps
            assert np.all(np.isfinite(got)) and np.abs(got - want).max() < FIT_TOL, (h2, mode, got, want)
        if h2 <= 0.9:
            assert total_fallbacks == 0
        if h2 >= 0.999:
            assert total_fallbacks > 0
    finally:
        eng.close()


def test_scattered_row_set_runs_as_a_prefix_after_row_permutation():
    """A Monte-Carlo style split (evaluator.py:555-561: random 80/20 of training + validation) has scattered universe
    positions.  The gather permutes the panel rows (training first, validation next), after which the contiguous
    kernels -- fused scaling included -- apply.  Same fitness as the position-lookup kernels and as the oracle."""
    from tblup_b200 import engine as E
    g = load_golden("fit_mid")
    x, y, h2 = g["x"], g["y"], float(g["h2"])
    tr, va, te = g["train"], g["valid"], g["test"]
    rng = np.random.default_rng(77)
    pool = rng.permutation(np.concatenate([tr, va]))
    mc_train, mc_valid = pool[:256], pool[256:]                 # 256 / 64, scattered over the universe prefix
    eng, perm = _engine(x, y, tr, va, te, extra_sets=[(mc_train, mc_valid)])
    try:
        m = x.shape[1]
        genomes = [rng.choice(m, size=k, replace=False) for k in (17, 320, 401, 1500)] + [rng.integers(0, m, size=600)]
        for mode, omode in ((E.MODE_GBLUP, O.MODE_GBLUP), (E.MODE_SNPBLUP, O.MODE_SNPBLUP)):
            eng.set_option("perm_rows", 1)
            a = eng.evaluate(genomes, slots=[1], h2=h2, mode=mode)[:, 0]
            assert eng.info("last_perm") == 1 and eng.info("last_fused_scale") == 1
            eng.set_option("perm_rows", 0)
            b = eng.evaluate(genomes, slots=[1], h2=h2, mode=mode)[:, 0]
            assert eng.info("last_perm") == 0
            eng.set_option("perm_rows", 1)
            want = np.array([O.exact_fitness(gen, mc_train, mc_valid, x, y, h2, omode) for gen in genomes])
            assert np.abs(a - want).max() < 1e-7 and np.abs(b - want).max() < 1e-7
        # the contiguous base split is untouched; together with the scattered set the call is evaluated one row set at
        # a time (a second Gram is far cheaper than position lookups in the refinement)
        base = eng.evaluate(genomes, slots=[0], h2=h2, mode=E.MODE_GBLUP)
        assert eng.info("last_perm") == 0
        both = eng.evaluate(genomes, slots=[0, 1], h2=h2, mode=E.MODE_GBLUP)
        assert eng.info("last_split") == 1 and np.abs(both[:, 0] - base[:, 0]).max() < 1e-12
        assert np.abs(both[:, 1] - np.array([O.exact_fitness(gen, mc_train, mc_valid, x, y, h2, O.MODE_GBLUP)
                                             for gen in genomes])).max() < 1e-7
        # fp64 precision goes through the same permuted panel
        eng.set_precision("fp64")
        c = eng.evaluate(genomes, slots=[1], h2=h2, mode=E.MODE_GBLUP)[:, 0]
        assert eng.info("last_perm") == 1
        want = np.array([O.exact_fitness(gen, mc_train, mc_valid, x, y, h2, O.MODE_GBLUP) for gen in genomes])
        assert np.abs(c - want).max() < 1e-9
    finally:
        eng.close()


@pytest.mark.parametrize("name", ["fit_small", "fit_mid"])
def test_two_ctas_per_matrix_solve_matches_one(name):
    """Small batches run the solve with a cluster of two CTAs per matrix (half of every block step, of the mat-vec work
    units and of the validation rows each; partial sums exchanged through distributed shared memory).  Forced on and
    off: same fitness (to the refinement tolerance) for the base split, the testing split and k-fold row sets, and
    against the reference fixtures."""
    from tblup_b200 import engine as E
    g = load_golden(name)
    x, y, h2 = g["x"], g["y"], float(g["h2"])
    tr, va, te = g["train"], g["valid"], g["test"]
    f_tr = unpack(g["fold_train_flat"], g["fold_train_off"])
    f_va = unpack(g["fold_valid_flat"], g["fold_valid_off"])
    extra = [(np.concatenate([tr, va]), te)] + [(t, v) for t, v in zip(f_tr, f_va)]
    eng, perm = _engine(x, y, tr, va, te, extra_sets=extra)
    genomes = [gen for gen in unpack(g["genomes_flat"], g["genomes_off"])]
    try:
        out = {}
        for mode in (0, 2):
            eng.set_option("solve_pair", mode)
            out[mode] = [eng.evaluate(genomes, slots=[0], h2=h2, mode=E.MODE_AUTO),
                         eng.evaluate(genomes, slots=[1], h2=h2, mode=E.MODE_AUTO),
                         eng.evaluate(genomes, slots=list(range(2, 2 + len(f_tr))), h2=h2, mode=E.MODE_AUTO)]
            assert eng.last_precision() == "mixed"
        for a, b in zip(out[0], out[2]):
            assert np.array_equal(np.isnan(a), np.isnan(b)) and np.nanmax(np.abs(a - b)) < 1e-9
        assert np.abs(out[2][0][:, 0] - g["ref_blup"]).max() < FIT_TOL
    finally:
        eng.close()


@pytest.mark.parametrize("name", ["fit_small", "fit_offset", "fit_mid"])
@pytest.mark.parametrize("pair", [0, 2])
def test_small_matrices_converge_without_the_fp64_fallback(name, pair):
    """Regression (r02): a warp left split by the CAS-loop shared atomics of the single-pass mat-vec reached the block
    barriers in pieces; with one work unit (n_t <= 128) the refinement diverged on EVERY small matrix and the fp64
    fallback silently repaired the result.  The default path has to converge by itself here: no fallback, two or three
    sweeps, with one and with two CTAs per matrix."""
    from tblup_b200 import engine as E
    g = load_golden(name)
    eng, _ = _engine(g["x"], g["y"], g["train"], g["valid"], g["test"])
    try:
        eng.set_option("solve_pair", pair)
        genomes = unpack(g["genomes_flat"], g["genomes_off"])
        for mode, ref in ((E.MODE_GBLUP, "ref_gblup"), (E.MODE_SNPBLUP, "ref_snp_blup")):
            got = eng.evaluate(genomes, slots=[0], h2=float(g["h2"]), mode=mode)[:, 0]
            assert eng.last_precision() == "mixed" and eng.info("last_fallbacks") == 0
            codes = [int(eng.debug_fetch(E.DBG_SWEEPS, j)[0]) for j in range(len(genomes))]
            assert all(c in (1, 2, 3) for c in codes), codes          # flags 256 (not converged) / 512 (pivot) clear
            assert np.abs(got - g[ref]).max() < FIT_TOL
    finally:
        eng.close()
