"""Synthetic genotypes / phenotypes / genome batches of the shapes BASELINE.json names:
p_j ~ U(0.05, 0.5), X_ij ~ Binomial(2, p_j) as int8, 1 % of markers are QTL with N(0,1) effects,
y = (X - 2p) beta + e with var(e) set from h2.  Used by bench.py and the full-size tests."""
import numpy as np


def synth_dataset(n, m, h2=0.4, seed=0, offset=0.0):
    rng = np.random.default_rng(seed)
    p = rng.uniform(0.05, 0.5, size=m)
    x = np.empty((n, m), dtype=np.int8)
    step = max(1, (1 << 24) // max(n, 1))
    for j0 in range(0, m, step):
        j1 = min(m, j0 + step)
        x[:, j0:j1] = rng.binomial(2, p[j0:j1], size=(n, j1 - j0)).astype(np.int8)
    n_qtl = max(1, m // 100)
    qtl = rng.choice(m, size=n_qtl, replace=False)
    beta = rng.standard_normal(n_qtl)
    g = (x[:, qtl].astype(np.float64) - 2 * p[qtl]) @ beta
    var_g = float(np.var(g)) or 1.0
    e = rng.standard_normal(n) * np.sqrt(var_g * (1 - h2) / h2)
    return x, g + e + offset


def synth_dataset_fast(n, m, h2=0.4, seed=0):
    """Same model as ``synth_dataset`` for very large shapes (config 4: 20 000 x 500 000 = 10 GB of dosages):
    one uniform byte per genotype compared against two per-marker thresholds instead of a binomial draw
    (dosage probabilities quantised to 1/256)."""
    rng = np.random.default_rng(seed)
    p = rng.uniform(0.05, 0.5, size=m)
    ge1 = np.clip(np.rint(256 * (1 - (1 - p) ** 2)), 1, 255).astype(np.uint8)    # P(x >= 1)
    eq2 = np.clip(np.rint(256 * p ** 2), 0, 254).astype(np.uint8)                 # P(x == 2)
    x = np.empty((n, m), dtype=np.int8)
    step = max(1, (1 << 28) // max(m, 1))
    for r0 in range(0, n, step):
        r1 = min(n, r0 + step)
        u = rng.integers(0, 256, size=(r1 - r0, m), dtype=np.uint8)
        x[r0:r1] = (u < ge1).astype(np.int8) + (u < eq2).astype(np.int8)
    n_qtl = max(1, m // 100)
    qtl = rng.choice(m, size=n_qtl, replace=False)
    beta = rng.standard_normal(n_qtl)
    g = (x[:, qtl].astype(np.float64) - 2 * p[qtl]) @ beta
    var_g = float(np.var(g)) or 1.0
    e = rng.standard_normal(n) * np.sqrt(var_g * (1 - h2) / h2)
    return x, g + e


def split_indices(n, seed=0, train_test=0.8, train_valid=0.8):
    """Shuffled train/validation/test split with the reference's proportions and rounding
    (tblup/evaluator.py:165-166,196-203: sklearn rounds the held-out part up)."""
    rng = np.random.default_rng(seed)
    order = rng.permutation(n)
    n_test = int(np.ceil((1 - train_test) * n - 1e-9))
    rest, test = order[:n - n_test], order[n - n_test:]
    n_valid = int(np.ceil((1 - train_valid) * len(rest) - 1e-9))
    train, valid = rest[:len(rest) - n_valid], rest[len(rest) - n_valid:]
    return train, valid, test


def random_genomes(P, m, k, seed=0):
    """P uniform random k-subsets without replacement (the initial random-key population,
    tblup/individual.py:152-156) packed as (flat int32, offsets int64)."""
    rng = np.random.default_rng(seed)
    flat = np.empty(P * k, dtype=np.int32)
    for i in range(P):
        flat[i * k:(i + 1) * k] = rng.permutation(m)[:k]
    off = np.arange(P + 1, dtype=np.int64) * k
    return flat, off
