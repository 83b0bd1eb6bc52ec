"""Start-up helpers of the reference that are built on the GRM of ALL markers, on the GPU.

``pca_splitter`` (tblup/evaluator.py:641-663) projects the full genomic relationship matrix ``make_grm(data)``
(tblup/utils.py:7-18: n x n x m float64 dgemm, minutes on the host at 5 000 x 50 000 and beyond reach at 20 000 x 500 000)
onto two principal components and splits the animals by their distance from the centroid.  Here the GRM comes from the
same kernels a fitness evaluation uses -- one "genome" listing every marker: E2M1 tcgen05 Gram (exact integers), exact
integer centring terms, one fp64 division per entry on the host -- and the PCA / sort are the reference's own library
calls on that matrix, so the split is the reference's.
"""
import numpy as np

from .engine import GblupEngine, MODE_GBLUP, DBG_C, DBG_S, DBG_SQ

_STAGE_GRAM = 3


def full_grm(data, device=0, storage="packed2"):
    """VanRaden GRM of all markers over all animals (what ``tblup.utils.make_grm(data)`` returns), n x n float64.

    ``data``: dense dosages (animals x markers) or a ``genoio.PackedGenotypes``."""
    n = data.shape[0]
    m = data.shape[1]
    if n < 4:
        raise ValueError("full_grm needs at least 4 animals")
    with GblupEngine(data, np.zeros(n), device=device, storage=storage) as eng:
        # a row set that covers every animal (the split itself is irrelevant: the pipeline stops after the Gram)
        eng.set_rowset(0, np.arange(n - 2), np.arange(n - 2, n))
        eng.set_option("stop_after", _STAGE_GRAM)
        eng.evaluate([np.arange(m)], slots=[0], h2=0.5, mode=MODE_GBLUP)
        rpad = eng.debug_dims(0)["rpad"]
        c = eng.debug_fetch(DBG_C, 0).astype(np.int64)[:n, :n]
        s = eng.debug_fetch(DBG_S, 0)[:n].astype(np.int64)
        S, Q = (int(v) for v in eng.debug_fetch(DBG_SQ, 0))
        eng.set_option("stop_after", -1)
    del rpad
    c = np.tril(c) + np.tril(c, -1).T
    N = n
    num = (N * N) * c - N * (s[:, None] + s[None, :]) + Q
    den = 2 * N * S - Q
    return 2.0 * num.astype(np.float64) / float(den)


def pca_splitter(data, split=0.8, outliers=False, device=0):
    """tblup/evaluator.py:641-663 with the GRM formed on the GPU; same return value (two index lists)."""
    from sklearn.decomposition import PCA
    proj = PCA(n_components=2)
    x = proj.fit_transform(full_grm(data, device=device))
    mu = np.mean(x, axis=0)
    dists = (x - mu) ** 2
    dists = dists[:, 0] + dists[:, 1]
    idx_dist = [(i, dists[i]) for i in range(len(dists))]
    idx_dist.sort(key=lambda tup: tup[1], reverse=outliers)
    idxs = [t[0] for t in idx_dist]
    k = int(len(idxs) * split)
    return idxs[:k], idxs[k:]
