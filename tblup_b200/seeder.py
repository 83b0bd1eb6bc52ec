"""Top-SNPs seeding (tblup/seeder.py) with the GWAS-style marker scan on the GPU.

The reference's ``SeedStrategy.get_sorted_indices`` (seeder.py:144-160) loads the dense float64 matrix, and for each of
five KFold splits calls the metric on ``X[train]`` -- for ``p_value`` (seeder.py:202-210) that is
``sklearn.feature_selection.f_regression``: a univariate regression of the phenotype on every marker.  Everything
f_regression computes follows from three per-marker sums over the fold's animals (sum x, sum x^2, sum x (y - ybar)),
which ``tb_marker_stats`` forms from the resident 2-bit matrix in one HBM pass; the F statistic, its p-value (scipy) and the
ranking are the reference's arithmetic on m numbers.  Reference quirk kept: the folds are taken over POSITIONS
0 .. len(training_indices)-1 and used as row numbers of the raw matrix (seeder.py:157-158).
"""
import numpy as np


def f_regression_from_sums(n, sx, sxx, sxw, w_norm):
    """(F, p) of sklearn.feature_selection.f_regression(X, y) (center=True, force_finite=True) from per-marker sums over
    the n samples: sx = sum x, sxx = sum x^2, sxw = sum x (y - ybar), w_norm = ||y - ybar||."""
    from scipy import stats
    with np.errstate(divide="ignore", invalid="ignore"):
        x_norms = np.sqrt(sxx - n * (sx / n) ** 2)
        corr = sxw / x_norms
        corr = corr / w_norm
    corr[np.isnan(corr)] = 0.0                       # constant marker (or constant phenotype)
    deg = n - 2
    with np.errstate(divide="ignore", invalid="ignore"):
        f = corr ** 2 / (1 - corr ** 2) * deg
        p = stats.f.sf(f, 1, deg)
    inf = np.isinf(f)
    f[inf] = np.finfo(f.dtype).max
    nan = np.isnan(f)
    f[nan] = 0.0
    p[nan] = 1.0
    return f, p


def p_value_scores(engine, rows, y):
    """``p_value(X[rows], y[rows])`` of the reference (negated p-values, seeder.py:202-210) from the resident matrix."""
    rows = np.asarray(rows)
    yy = np.asarray(y, dtype=np.float64).ravel()[rows]
    w = yy - yy.mean()
    sx, sxx, sxw = engine.marker_stats(rows, w)
    _, p = f_regression_from_sums(len(rows), sx, sxx, sxw, float(np.linalg.norm(w)))
    return -1 * p


def sorted_indices(engine, y, n_training, n_splits=5):
    """The reference's ``get_sorted_indices``: summed metric over KFold(n_splits) splits of the positions
    0 .. n_training-1, markers in descending order of the sum."""
    from sklearn.model_selection import KFold
    scores = np.zeros(engine.m)
    for train, _ in KFold(n_splits=n_splits).split(np.arange(n_training)):
        scores += p_value_scores(engine, train, y)
    return np.flip(np.argsort(scores, axis=0), 0), scores


class TopSNPsSeedStrategy:
    """Drop-in for tblup.seeder.TopSNPsSeedStrategy (same attributes and methods; seeder.py:112-199) whose marker
    ranking comes from the GPU scan.  ``tblup_b200.install`` derives it from the reference class when that is importable."""

    N_SPLITS = 5

    def __init__(self, evaluator, metric, geno_path, pheno_path):
        try:
            self.training_indices = evaluator.training_indices
        except AttributeError:
            raise AttributeError("The provided evaluator {} does not calculate training indices, which are needed "
                                 "for a seeder to filter the data.".format(evaluator.__class__.__name__))
        self.metric = metric
        self.indices = self.get_sorted_indices(geno_path, pheno_path)
        self.current_index = 0

    def get_sorted_indices(self, geno_path, pheno_path):
        from .engine import GblupEngine
        from .evaluator import _devices_from_env
        from .genoio import load_genotypes
        y = np.load(pheno_path)
        geno = load_genotypes(geno_path, int(np.asarray(y).size))
        with GblupEngine(geno, y, device=_devices_from_env()[0]) as engine:
            order, self.scores = sorted_indices(engine, y, len(self.training_indices), self.N_SPLITS)
        return order

    def get_next_indices(self, length):
        n = self.current_index
        self.current_index += length
        if self.current_index > len(self.indices):
            return np.random.choice(self.indices, length, replace=False)
        return self.indices[n:n + length]

    def reset(self):
        self.current_index = 0
