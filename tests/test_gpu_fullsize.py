"""BASELINE.json's headline shape (config 2: 5 000 animals x 50 000 markers, k = 5 001) on the GPU, checked through
size-independent properties (the CPU oracle needs ~5 s per genome at this size, so only two genomes are compared
against it directly):

* the tcgen05 Gram equals the plain dp4a Gram bit for bit on a full-size genome;
* the integer Gram is a sum over markers: permuting the genome changes nothing (bit-exact fitness in fp64 mode), and a
  marker listed twice is not the same as listed once;
* mixed precision (TF32 factor + fp64 refinement) agrees with the fp64 factorisation far inside the 1e-6 bar;
* a batch evaluated in several waves equals the batch evaluated at once;
* two genomes against the reference algorithm on the CPU (oracle.ref_blup).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N, M, K = 5000, 50000, 5001


@pytest.fixture(scope="module")
def big():
    from tblup_b200 import GblupEngine, synth
    x, y = synth.synth_dataset(N, M, h2=0.4, seed=0)
    tr, va, te = synth.split_indices(N, seed=0)
    eng = GblupEngine(x, y, perm=np.concatenate([tr, va, te]))
    eng.set_rowset(0, tr, va)
    eng.set_rowset(1, np.concatenate([tr, va]), te)
    yield eng, x, y, tr, va, te
    eng.close()


def test_gram_tc_equals_simt_full_size(big):
    eng = big[0]
    rng = np.random.default_rng(1)
    idx = rng.choice(M, size=K, replace=False)
    a = eng.gram_debug(idx, 4000, impl="tc")
    b = eng.gram_debug(idx, 4000, impl="simt")
    assert np.array_equal(a, b)
    assert a.max() <= 4 * K and a.min() >= 0 and np.all(np.diag(a) >= np.abs(a).max(axis=1) // 2)


def test_permutation_invariance_and_multiset(big):
    from tblup_b200 import engine as E
    eng = big[0]
    rng = np.random.default_rng(2)
    g = rng.choice(M, size=K, replace=False)
    dup = np.concatenate([g[:-1], g[:1]])               # same length, first marker twice instead of the last marker
    eng.set_precision("fp64")
    f = eng.evaluate([g, rng.permutation(g), np.sort(g), dup], mode=E.MODE_GBLUP)[:, 0]
    assert f[0] == f[1] == f[2]
    assert f[3] != f[0] and abs(f[3] - f[0]) < 0.05


def test_mixed_agrees_with_fp64_and_waves_do_not_matter(big):
    from tblup_b200 import engine as E, synth
    eng = big[0]
    flat, off = synth.random_genomes(24, M, K, seed=7)
    eng.set_precision("fp64")
    ref = eng.evaluate_packed(flat, off, slots=[0], mode=E.MODE_AUTO)[:, 0]
    eng.set_precision("mixed")
    mix = eng.evaluate_packed(flat, off, slots=[0], mode=E.MODE_AUTO)[:, 0]
    assert eng.last_precision() == "mixed"
    assert np.abs(mix - ref).max() < 1e-8
    eng.set_option("max_wave", 7)
    again = eng.evaluate_packed(flat, off, slots=[0], mode=E.MODE_AUTO)[:, 0]
    eng.set_option("max_wave", 0)
    assert eng.last_wave() == 7 or eng.last_wave() == 24
    assert np.array_equal(again, mix)
    test_fit = eng.evaluate_packed(flat[:off[4]], off[:5], slots=[1], mode=E.MODE_AUTO)[:, 0]     # 4 000 train -> 1 000 test
    assert np.all(np.isfinite(test_fit)) and np.all((test_fit >= 0) & (test_fit <= 1))


def test_fused_chain_equals_launch_per_stage_chain(big):
    """Small waves run the diagonal-block chain of a block column as ONE kernel on shared-memory operands
    (chol_chain256_kernel), large ones as a launch per stage (chol_diag32 / chol_narrow): same device code, same order of
    operations, so the factor and with it every fitness is bit-identical; neither path may lean on the fp64 fallback.
    Slot 1 (4 000 training animals) ends in a partial block column."""
    from tblup_b200 import engine as E, synth
    eng = big[0]
    flat, off = synth.random_genomes(12, M, K, seed=11)
    eng.set_precision("mixed")
    out = {}
    for fused, inverse in ((0, 1), (1 << 20, 0), (1 << 20, 1)):
        eng.set_option("chain_fused", fused)
        eng.set_option("chain_inverse", inverse)       # 0: the fused kernel leaves the 256-block inverse to trinv256_kernel
        out[fused, inverse] = [eng.evaluate_packed(flat, off, slots=[s], mode=E.MODE_AUTO)[:, 0] for s in (0, 1)]
        assert eng.last_precision() == "mixed" and eng.info("last_fallbacks") == 0
    eng.set_option("chain_fused", -1)
    for key in ((1 << 20, 0), (1 << 20, 1)):
        for a, b in zip(out[0, 1], out[key]):
            assert np.all(np.isfinite(a)) and np.array_equal(a, b), key


def test_half_precision_block_columns_agree_with_fp32_block_columns(big):
    """t16 (default): the entries of a block column below its diagonal block live as halves between the update that forms
    them and the panel GEMM that turns them into factor entries (fp16 operands); t16 = 0 keeps them in fp32 and multiplies
    in TF32.  Both are 10-bit preconditioners of the same exact operator: the refined fitness agrees far inside the bar."""
    from tblup_b200 import engine as E, synth
    eng = big[0]
    flat, off = synth.random_genomes(6, M, K, seed=13)
    eng.set_precision("mixed")
    out = {}
    for t16 in (1, 0):
        eng.set_option("t16", t16)
        out[t16] = eng.evaluate_packed(flat, off, slots=[0], mode=E.MODE_AUTO)[:, 0]
        assert eng.last_precision() == "mixed" and eng.info("last_fallbacks") == 0
    eng.set_option("t16", 1)
    assert np.abs(out[1] - out[0]).max() < 1e-8


def test_two_genomes_against_the_reference_algorithm(big):
    from oracle import gblup_oracle as O
    from tblup_b200 import engine as E
    eng, x, y, tr, va, te = big
    rng = np.random.default_rng(3)
    genomes = [rng.choice(M, size=K, replace=False), rng.choice(M, size=3000, replace=False)]   # gblup and snp_blup branches
    eng.set_precision("mixed")
    got = eng.evaluate(genomes, slots=[0], mode=E.MODE_AUTO)[:, 0]
    xf = x.astype(np.float64)
    for gnm, f in zip(genomes, got):
        assert abs(O.ref_blup(gnm.astype(int), list(tr), list(va), xf, y, 0.4) - f) < 1e-6


@pytest.mark.parametrize("k", [K, 50000])
def test_fp4_gram_bit_exact_full_size(big, k):
    """The DEFAULT Gram (tcgen05 kind::mxf4 on E2M1 nibbles) against the exact integer oracle at the headline k and at
    k = 50 000 (config 4's subset size: sums up to 200 000, beyond int16, every k-block of the panel in use)."""
    from oracle import gblup_oracle as O
    eng, x, y, tr, va, te = big
    rng = np.random.default_rng(11 + k)
    idx = rng.choice(M, size=k, replace=False)
    rows = 4000
    got = eng.gram_debug(idx, rows, impl="fp4")
    want = np.tril(O.exact_gram(x, idx, np.concatenate([tr, va])[:rows]))
    assert np.array_equal(got.astype(np.int64), want)
    assert got.max() > (32767 if k == 50000 else 0)


@pytest.mark.parametrize("impl", ["tc_pair", "fp4_pair", "tc_cg2", "fp4_cg2"])
def test_paired_gram_equals_single_cta_gram_full_size(big, impl):
    """The paired Gram schedules (clusters of two CTAs: B tile shared by TMA multicast, or one tcgen05 CTA pair with
    cta_group::2 MMAs) against the single-CTA kernel, bit for bit, at the headline shape and on row counts that leave an odd number of row blocks / a ragged last tile."""
    eng = big[0]
    rng = np.random.default_rng(17)
    idx = rng.choice(M, size=K, replace=False)
    for rows in (4000, 3200, 129, 385, 1):
        a = eng.gram_debug(idx[:K if rows > 400 else 700], rows, impl=impl)
        b = eng.gram_debug(idx[:K if rows > 400 else 700], rows, impl=impl.split("_")[0])
        assert np.array_equal(a, b), rows


def test_k50000_fitness_uses_int32_cross_products(big):
    """k = 50 000: 4 k > 32 767, so the wave keeps int32 cross-products (fused scaling, mixed precision still on);
    fitness against the exact oracle."""
    from oracle import gblup_oracle as O
    from tblup_b200 import engine as E
    eng, x, y, tr, va, te = big
    rng = np.random.default_rng(5)
    genomes = [rng.permutation(M), rng.choice(M, size=40000, replace=False)]
    eng.set_precision("mixed")
    got = eng.evaluate(genomes, slots=[0], mode=E.MODE_AUTO)[:, 0]
    assert eng.info("last_c16") == 0 and eng.info("last_fp4") == 1 and eng.last_precision() == "mixed"
    want = [O.exact_blup(g, tr, va, x, y, 0.4) for g in genomes]
    assert np.abs(got - np.array(want)).max() < 1e-6


def test_config3_shape_ten_aligned_folds(big):
    """BASELINE config 3: 10-fold intra-generation CV over the 3 200 training animals (2 880 train / 320 held out per
    fold, tblup/evaluator.py:455-483, :509-537): one Gram per genome, ten factorisations; four genomes (both branches of
    blup()) against the exact oracle, fold by fold."""
    from oracle import gblup_oracle as O
    from tblup_b200 import GblupEngine, engine as E
    _, x, y, tr, va, te = big
    folds = [(np.asarray(t), np.asarray(v)) for t, v in O.ref_make_fold_indices(list(tr), 10)]
    assert [len(v) for _, v in folds] == [320] * 10
    rng = np.random.default_rng(9)
    genomes = [rng.choice(M, size=k, replace=False) for k in (K, K, 4800, 5000)]
    with GblupEngine(x, y, perm=np.concatenate([tr, va, te])) as eng:
        for f, (t, v) in enumerate(folds):
            eng.set_rowset(f, t, v)
        got = eng.evaluate(genomes, slots=list(range(10)), mode=E.MODE_AUTO)
        assert eng.last_precision() == "mixed" and eng.info("last_split") == 0
    want = np.array([O.exact_fitness_rowsets(g, folds, x, y, 0.4) for g in genomes])
    assert got.shape == want.shape == (4, 10)
    assert np.abs(got - want).max() < 1e-6


def test_nan_semantics_on_device():
    """scipy.stats.pearsonr conventions of tblup/evaluator.py:286,:314 on the GPU: a constant validation phenotype or a
    genome of monomorphic markers (G = 0/0) gives NaN -- and NaN never wins the reference's selection
    (tblup/selector.py:28: ``child.fitness > parent.fitness``)."""
    from tblup_b200 import GblupEngine, engine as E, synth
    n, m = 600, 3000
    x, y = synth.synth_dataset(n, m, h2=0.4, seed=21)
    x[:, :40] = 0                                   # monomorphic markers
    tr, va, te = synth.split_indices(n, seed=21)
    y_const = y.copy()
    y_const[va] = 3.25
    rng = np.random.default_rng(1)
    ok_small, ok_big = rng.choice(np.arange(40, m), size=300, replace=False), rng.choice(np.arange(40, m), size=700, replace=False)
    mono = np.arange(40)
    mono_big = np.concatenate([np.arange(40)] * 16)          # 640 > n: the gblup branch on monomorphic columns
    for precision in ("mixed", "fp64"):
        with GblupEngine(x, y_const, perm=np.concatenate([tr, va, te])) as eng:
            eng.set_precision(precision)
            eng.set_rowset(0, tr, va)
            f = eng.evaluate([ok_small, ok_big], slots=[0], mode=E.MODE_AUTO)[:, 0]
            assert np.all(np.isnan(f)), f
        with GblupEngine(x, y, perm=np.concatenate([tr, va, te])) as eng:
            eng.set_precision(precision)
            eng.set_rowset(0, tr, va)
            f = eng.evaluate([ok_small, mono, ok_big, mono_big], slots=[0], mode=E.MODE_AUTO)[:, 0]
            assert np.isfinite(f[0]) and np.isfinite(f[2]) and np.isnan(f[1]) and np.isnan(f[3]), f
            # greedy selection exactly as tblup/selector.py:28 writes it
            parent, child = 0.1, float(f[1])
            assert not (child > parent)
