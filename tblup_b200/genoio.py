"""Genotype containers either side of the hot path: dense ``.npy`` dosages (the only format the reference knows,
tblup/utils.py:95, tblup/evaluator.py:188,215) and the 2-bit SNP-major packing the device can keep resident.

Packed layout (``PackedGenotypes.data``): uint8 ``[m][ceil(n / 4)]``, one row per marker, animal ``4q + i`` in bits
``2i .. 2i+1`` of byte ``q`` -- the row layout and bit order of a PLINK ``.bed`` body.  Each 2-bit code is the dosage
itself (0, 1, 2); code 3 never occurs in a valid container (padding bits of the last byte are 0).

``read_bed`` / ``write_bed`` translate between this container and PLINK 1 binary files (magic ``6c 1b 01``,
SNP-major): PLINK codes 00 / 10 / 11 are 2 / 1 / 0 copies of the first allele and 01 is a missing call, which has no
counterpart in the reference (its ``make_grm`` takes ``np.mean`` over whatever numbers it is given) and is rejected.

The byte shuffling here is format conversion on the host, done once per data set; the per-generation arithmetic
reads the packed rows on the GPU (gather_kernel<PACKED>, tblup_b200/csrc/gather.cu).
"""
import os

import numpy as np

BED_MAGIC = bytes((0x6C, 0x1B, 0x01))
# PLINK code (2 bits) -> dosage of the first allele; 01 = missing
_PLINK_TO_DOSAGE = np.array([2, 3, 1, 0], dtype=np.uint8)
_DOSAGE_TO_PLINK = np.array([3, 2, 0, 1], dtype=np.uint8)       # dosage 0 -> 11, 1 -> 10, 2 -> 00


def _byte_table(code_map):
    """256-entry table applying a 2-bit -> 2-bit map to the four fields of a byte."""
    b = np.arange(256, dtype=np.uint16)
    out = np.zeros(256, dtype=np.uint16)
    for i in range(4):
        out |= code_map[(b >> (2 * i)) & 3].astype(np.uint16) << (2 * i)
    return out.astype(np.uint8)


class PackedGenotypes:
    """2-bit SNP-major dosages: ``data`` uint8 [m][ceil(n/4)], ``n`` animals, ``m`` markers."""

    def __init__(self, data, n):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        if data.ndim != 2 or data.shape[1] != (n + 3) // 4:
            raise ValueError("packed genotypes must be [markers][ceil(n/4)] bytes, got %r for n = %d" % (data.shape, n))
        self.data = data
        self.n = int(n)
        self.m = int(data.shape[0])

    @property
    def shape(self):                      # (animals, markers), like the dense matrix it stands for
        return (self.n, self.m)

    @property
    def nbytes(self):
        return self.data.nbytes

    def unpack(self, markers=None):
        """Dense int8 [n][len(markers)] (all markers when None)."""
        rows = self.data if markers is None else self.data[np.asarray(markers)]
        fields = np.stack([(rows >> (2 * i)) & 3 for i in range(4)], axis=-1)      # [.., q, i]
        return np.ascontiguousarray(fields.reshape(rows.shape[0], -1)[:, :self.n].T.astype(np.int8))


def pack_dosages(geno):
    """Dense dosages (animals x markers, any real dtype holding 0/1/2) -> PackedGenotypes."""
    from .engine import as_dosage_int8
    g = as_dosage_int8(geno)
    n, m = g.shape
    q = (n + 3) // 4
    t = np.zeros((m, 4 * q), dtype=np.uint8)
    t[:, :n] = g.T
    t = t.reshape(m, q, 4)
    data = t[:, :, 0] | (t[:, :, 1] << 2) | (t[:, :, 2] << 4) | (t[:, :, 3] << 6)
    return PackedGenotypes(data, n)


def read_bed(path, n):
    """PLINK 1 .bed (SNP-major) with ``n`` samples -> PackedGenotypes of first-allele dosages.  Raises on a bad
    magic number, a truncated body or a missing call."""
    stride = (n + 3) // 4
    size = os.path.getsize(path)
    with open(path, "rb") as f:
        if f.read(3) != BED_MAGIC:
            raise ValueError("%s: not a SNP-major PLINK 1 .bed file (magic bytes 6c 1b 01 expected)" % path)
        body = np.fromfile(f, dtype=np.uint8)
    if size < 3 + stride or (size - 3) % stride != 0:
        raise ValueError("%s: body of %d bytes is not a whole number of %d-byte marker rows (n = %d)"
                         % (path, size - 3, stride, n))
    rows = body.reshape(-1, stride)
    data = _byte_table(_PLINK_TO_DOSAGE)[rows]
    # padding fields of the last byte are 00 in the file -> dosage code 2 after translation: clear them
    pad = 4 * stride - n
    if pad:
        data[:, -1] &= np.uint8(0xFF >> (2 * pad))
    # a missing call (01) became code 3: look for a field with both bits set
    both = data & (data >> 1) & np.uint8(0x55)
    if both.any():
        j, qb = np.argwhere(both)[0]
        raise ValueError("%s: missing genotype call at marker %d (animals %d..%d); the GBLUP path has no "
                         "missing-value handling -- impute first" % (path, j, 4 * qb, 4 * qb + 3))
    return PackedGenotypes(data, n)


def write_bed(path, packed):
    """PackedGenotypes -> PLINK 1 .bed (SNP-major); padding fields written as 00 like PLINK does."""
    data = _byte_table(_DOSAGE_TO_PLINK)[packed.data]
    pad = 4 * packed.data.shape[1] - packed.n
    if pad:
        data[:, -1] &= np.uint8(0xFF >> (2 * pad))
    with open(path, "wb") as f:
        f.write(BED_MAGIC)
        data.tofile(f)


def load_genotypes(path, n=None):
    """``.npy`` -> validated int8 dosages (animals x markers); ``.bed`` -> PackedGenotypes (needs ``n``)."""
    if str(path).endswith(".bed"):
        if n is None:
            raise ValueError("reading a .bed file needs the number of animals (length of the phenotype vector)")
        return read_bed(path, n)
    from .engine import as_dosage_int8
    return as_dosage_int8(np.load(path))
