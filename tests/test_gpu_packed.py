"""2-bit packed resident genotypes (north_star (a)) against the int8 path and the oracle, through the
C-ABI (tb_create_ex).  Integer stages bit-exact; fitness identical between the two storages."""
import numpy as np
import pytest

from conftest import load_golden, unpack
from oracle import gblup_oracle as O

pytestmark = pytest.mark.gpu


def _perm(g):
    return np.concatenate([g["train"], g["valid"], g["test"]]).astype(np.int64)


@pytest.fixture(scope="module")
def engines():
    from tblup_b200 import GblupEngine
    from tblup_b200.genoio import pack_dosages
    g = load_golden("fit_mid")
    perm = _perm(g)
    made = {
        "int8": GblupEngine(g["x"], g["y"], perm=perm, storage="int8"),
        "packed_from_dense": GblupEngine(g["x"], g["y"], perm=perm, storage="packed2"),
        "packed_from_packed": GblupEngine(pack_dosages(g["x"]), g["y"], perm=perm, storage="packed2"),
        "int8_from_packed": GblupEngine(pack_dosages(g["x"]), g["y"], perm=perm, storage="int8"),
    }
    for e in made.values():
        e.set_rowset(0, g["train"], g["valid"])
        e.set_rowset(1, np.concatenate([g["train"], g["valid"]]), g["test"])
    yield g, perm, made
    for e in made.values():
        e.close()


def test_resident_bytes(engines):
    g, perm, made = engines
    n, m = g["x"].shape
    ldn = (n + 127) // 128 * 128
    assert made["int8"].resident_genotype_bytes() == m * ldn
    assert made["packed_from_dense"].resident_genotype_bytes() == m * ldn // 4
    assert made["packed_from_packed"].resident_genotype_bytes() == m * ldn // 4


@pytest.mark.parametrize("which", ["packed_from_dense", "packed_from_packed", "int8_from_packed"])
@pytest.mark.parametrize("k", [1, 130, 1500, -700])
def test_gram_bit_exact(engines, which, k):
    g, perm, made = engines
    rng = np.random.default_rng(abs(k) + 3)
    m = g["x"].shape[1]
    idx = rng.integers(0, m, size=-k) if k < 0 else rng.choice(m, size=k, replace=False)     # k < 0: with duplicates
    rows = 389
    got = made[which].gram_debug(idx, rows, impl="tc")
    assert np.array_equal(got.astype(np.int64), np.tril(O.exact_gram(g["x"], idx, perm[:rows])))


@pytest.mark.parametrize("which", ["packed_from_dense", "packed_from_packed", "int8_from_packed"])
def test_fitness_and_centring_identical_to_int8(engines, which):
    from tblup_b200 import engine as E
    g, perm, made = engines
    genomes = unpack(g["genomes_flat"], g["genomes_off"])
    h2 = float(g["h2"])
    for mode in (E.MODE_GBLUP, E.MODE_SNPBLUP, E.MODE_AUTO):
        a = made["int8"].evaluate(genomes, slots=[0, 1], h2=h2, mode=mode)
        sa = made["int8"].debug_fetch(E.DBG_S, 0)
        qa = made["int8"].debug_fetch(E.DBG_SQ, 0)
        b = made[which].evaluate(genomes, slots=[0, 1], h2=h2, mode=mode)
        assert np.array_equal(sa, made[which].debug_fetch(E.DBG_S, 0))       # training-row column sums went in
        assert np.array_equal(qa, made[which].debug_fetch(E.DBG_SQ, 0))
        assert np.array_equal(a, b)                                          # same integers in, same kernels after
    want = np.array([O.exact_blup(gen, g["train"], g["valid"], g["x"], g["y"], h2) for gen in genomes])
    got = made[which].evaluate(genomes, slots=[0], h2=h2, mode=E.MODE_AUTO)[:, 0]
    assert np.abs(got - want).max() < 1e-6


def test_packed_input_with_code_3_is_rejected():
    from tblup_b200 import GblupEngine
    from tblup_b200.genoio import PackedGenotypes
    rng = np.random.default_rng(0)
    x = rng.integers(0, 3, size=(64, 200)).astype(np.int8)
    from tblup_b200.genoio import pack_dosages
    p = pack_dosages(x)
    bad = p.data.copy()
    bad[17, 3] |= 0b1100          # animal 13 of marker 17 -> code 3
    with pytest.raises(RuntimeError, match="dosages"):
        GblupEngine(PackedGenotypes(bad, 64), np.zeros(64))


def test_evaluator_reads_bed(tmp_path, monkeypatch):
    """The drop-in evaluator on a PLINK .bed file + packed residency gives the fitness of the dense .npy run."""
    import random
    from tblup_b200 import evaluator as EV
    from tblup_b200.genoio import pack_dosages, write_bed
    g = load_golden("fit_mid")
    x, y = g["x"][:300, :1200], g["y"][:300]
    np.save(tmp_path / "g.npy", x.astype(np.float64))
    np.save(tmp_path / "y.npy", y)
    write_bed(str(tmp_path / "g.bed"), pack_dosages(x))
    rng = np.random.default_rng(5)
    genomes = [rng.choice(1200, size=k, replace=False) for k in (40, 300, 301, 700)]

    def run(path, storage):
        monkeypatch.setenv("TBLUP_B200_STORAGE", storage)
        random.seed(3)
        np.random.seed(3)
        ev = EV.BlupParallelEvaluator(str(path), str(tmp_path / "y.npy"), 0.4, snp_remover=EV.SNPRemovalHandler(10, 0.1, 0.4, False))
        with ev:
            return ev._fitness_matrix(genomes, [0, EV.TESTING_SLOT]), ev.training_indices

    a, tr_a = run(tmp_path / "g.npy", "int8")
    b, tr_b = run(tmp_path / "g.bed", "packed2")
    assert tr_a == tr_b
    assert np.array_equal(a, b)


@pytest.mark.parametrize("k", [1, 63, 255, 256, 257, 1500, -900, -4000])
def test_fp4_gram_bit_exact(engines, k):
    """E2M1 operands on tcgen05 kind::mxf4 (all block scales 2^0): the fp32 accumulators must hold exactly the
    integers of the int8 path / the oracle, for every k-block remainder and with duplicated markers."""
    g, perm, made = engines
    rng = np.random.default_rng(abs(k) + 17)
    m = g["x"].shape[1]
    idx = rng.integers(0, m, size=-k) if k < 0 else rng.choice(m, size=k, replace=False)
    for rows in (389, 128, 31):
        got = made["packed_from_dense"].gram_debug(idx, rows, impl="fp4")
        assert np.array_equal(got.astype(np.int64), np.tril(O.exact_gram(g["x"], idx, perm[:rows])))


def test_fp4_and_int8_gram_give_identical_fitness(engines):
    from tblup_b200 import engine as E
    g, perm, made = engines
    genomes = unpack(g["genomes_flat"], g["genomes_off"])
    h2 = float(g["h2"])
    eng = made["packed_from_dense"]
    eng.set_option("gram_fp4", 1)
    a = eng.evaluate(genomes, slots=[0, 1], h2=h2, mode=E.MODE_AUTO)
    assert eng.info("last_fp4") == 1
    ca = eng.debug_fetch(E.DBG_C, 2)
    eng.set_option("gram_fp4", 0)
    b = eng.evaluate(genomes, slots=[0, 1], h2=h2, mode=E.MODE_AUTO)
    assert eng.info("last_fp4") == 0
    cb = eng.debug_fetch(E.DBG_C, 2)
    eng.set_option("gram_fp4", 1)
    nt = len(g["train"])
    assert np.array_equal(np.tril(ca)[:, :nt], np.tril(cb)[:, :nt])
    assert np.array_equal(a, b)


def test_clone_holds_the_same_resident_data():
    """tb_clone: a second context filled by a device-to-device copy (same GPU here; tests/test_gpu_large.py covers two
    GPUs) gives bit-identical cross-products and fitness."""
    from tblup_b200 import GblupEngine, engine as E, synth
    n, m = 500, 3000
    x, y = synth.synth_dataset(n, m, h2=0.4, seed=51)
    tr, va, te = synth.split_indices(n, seed=51)
    rng = np.random.default_rng(51)
    genomes = [rng.choice(m, size=kk, replace=False) for kk in (100, 501, 900)]
    for storage in ("packed2", "int8"):
        with GblupEngine(x, y, perm=np.concatenate([tr, va, te]), storage=storage) as a:
            a.set_rowset(0, tr, va)
            with a.clone(0) as b:
                b.set_rowset(0, tr, va)
                assert b.resident_genotype_bytes() == a.resident_genotype_bytes()
                impl = "fp4" if storage == "packed2" else "tc"
                assert np.array_equal(a.gram_debug(genomes[1], 400, impl=impl), b.gram_debug(genomes[1], 400, impl=impl))
                fa = a.evaluate(genomes, slots=[0], mode=E.MODE_AUTO)
                fb = b.evaluate(genomes, slots=[0], mode=E.MODE_AUTO)
                assert np.array_equal(fa, fb)
