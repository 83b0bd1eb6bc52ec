"""Host-side genotype containers: 2-bit SNP-major packing and PLINK .bed translation, against the
oracle's loop-form restatement and a hand-written known-answer file."""
import numpy as np
import pytest

from oracle import gblup_oracle as O
from tblup_b200 import genoio as G


@pytest.mark.parametrize("n", [1, 3, 4, 5, 64, 131])
def test_pack_matches_loop_definition_and_round_trips(n):
    rng = np.random.default_rng(n)
    x = rng.integers(0, 3, size=(n, 23)).astype(np.int8)
    p = G.pack_dosages(x)
    assert p.shape == x.shape and p.data.shape == (23, (n + 3) // 4)
    assert np.array_equal(p.data, O.pack2_loops(x))
    assert np.array_equal(p.unpack(), x)
    pick = [5, 5, 0, 22]
    assert np.array_equal(p.unpack(pick), x[:, pick])


def test_pack_accepts_the_reference_float_matrix_and_rejects_non_dosages():
    x = np.array([[0.0, 1.0, 2.0], [2.0, 2.0, 0.0]])
    assert np.array_equal(G.pack_dosages(x).unpack(), x.astype(np.int8))
    with pytest.raises(ValueError):
        G.pack_dosages(np.array([[0.5, 1.0]]))
    with pytest.raises(ValueError):
        G.pack_dosages(np.array([[3, 1]], dtype=np.int8))


def test_bed_known_answer(tmp_path):
    """Five samples, two markers, written by hand from the PLINK 1 format description: sample i in bits 2i..2i+1,
    00 = two copies of the first allele, 10 = one, 11 = none."""
    #            s0  s1  s2  s3 | s4
    # marker 0:   2   1   0   2 |  1      -> 00 10 11 00 | 10  -> byte0 = 0b00_11_10_00 = 0x38, byte1 = 0b10 = 0x02
    # marker 1:   0   0   2   1 |  2      -> 11 11 00 10 | 00  -> byte0 = 0b10_00_11_11 = 0x8f, byte1 = 0x00
    raw = bytes([0x6C, 0x1B, 0x01, 0x38, 0x02, 0x8F, 0x00])
    path = tmp_path / "tiny.bed"
    path.write_bytes(raw)
    p = G.read_bed(str(path), 5)
    want = np.array([[2, 0], [1, 0], [0, 2], [2, 1], [1, 2]], dtype=np.int8)
    assert np.array_equal(p.unpack(), want)
    assert raw == O.bed_bytes_loops(want)
    out = tmp_path / "again.bed"
    G.write_bed(str(out), p)
    assert out.read_bytes() == raw


@pytest.mark.parametrize("n", [4, 7, 130])
def test_bed_round_trip_matches_oracle_bytes(tmp_path, n):
    rng = np.random.default_rng(100 + n)
    x = rng.integers(0, 3, size=(n, 41)).astype(np.int8)
    path = tmp_path / "g.bed"
    G.write_bed(str(path), G.pack_dosages(x))
    assert path.read_bytes() == O.bed_bytes_loops(x)
    assert np.array_equal(G.read_bed(str(path), n).unpack(), x)


def test_bed_errors(tmp_path):
    bad_magic = tmp_path / "a.bed"
    bad_magic.write_bytes(bytes([0x6C, 0x1B, 0x00, 0x00]))
    with pytest.raises(ValueError, match="magic"):
        G.read_bed(str(bad_magic), 4)
    truncated = tmp_path / "b.bed"
    truncated.write_bytes(bytes([0x6C, 0x1B, 0x01, 0x00, 0x00, 0x00]))
    with pytest.raises(ValueError, match="whole number"):
        G.read_bed(str(truncated), 5)          # 2 bytes per marker, 3 bytes of body
    missing = tmp_path / "c.bed"
    missing.write_bytes(bytes([0x6C, 0x1B, 0x01, 0b00_01_00_00]))
    with pytest.raises(ValueError, match="missing genotype"):
        G.read_bed(str(missing), 4)
    with pytest.raises(ValueError):
        G.load_genotypes(str(missing))          # .bed needs the animal count


def test_load_genotypes_npy(tmp_path):
    x = np.random.default_rng(1).integers(0, 3, size=(6, 9)).astype(np.float64)
    path = tmp_path / "g.npy"
    np.save(path, x)
    got = G.load_genotypes(str(path))
    assert got.dtype == np.int8 and np.array_equal(got, x)
