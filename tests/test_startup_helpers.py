"""Start-up helpers of SURVEY 8f row F4 -- pca_splitter (tblup/evaluator.py:641-663) and the top-SNPs seeder
(tblup/seeder.py:144-160,202-210) -- against outputs of the LIVE reference recorded by tests/golden/make_golden_split.py.
CPU part: the restatements and the host arithmetic; GPU part (marked): the GRM / marker scan from the resident matrix."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import gblup_oracle as O


@pytest.fixture(scope="module")
def g():
    return load_golden("split_seed")


@pytest.mark.parametrize("tag", ["a", "b"])
def test_oracle_restatements_match_the_reference(g, tag):
    x, y = g["x_" + tag], g["y_" + tag]
    assert np.abs(O.ref_make_grm(x.astype(np.float64)) - g["grm_" + tag]).max() < 1e-12
    for outl in (0, 1):
        tr, te = O.ref_pca_split(g["grm_" + tag], outliers=bool(outl))
        assert tr == list(g["pca_train_%s_%d" % (tag, outl)]) and te == list(g["pca_test_%s_%d" % (tag, outl)])
    scores = O.ref_seed_scores(x, y, len(g["seed_train_" + tag]))
    assert np.array_equal(scores, g["seed_scores_" + tag])
    assert np.array_equal(np.flip(np.argsort(scores, axis=0), 0), g["seed_order_" + tag])


@pytest.mark.parametrize("tag", ["a", "b"])
def test_f_regression_from_sums_is_sklearn_f_regression(g, tag):
    """The three per-marker sums the device forms are all sklearn's f_regression needs -- constant markers (NaN -> F 0,
    p 1 with force_finite) and duplicated markers included."""
    from sklearn.feature_selection import f_regression
    from tblup_b200.seeder import f_regression_from_sums
    x, y = g["x_" + tag].astype(np.float64), g["y_" + tag]
    rows = np.arange(0, x.shape[0], 2)
    xs, ys = x[rows], y[rows]
    w = ys - ys.mean()
    f, p = f_regression_from_sums(len(rows), xs.sum(0), (xs * xs).sum(0), xs.T @ w, float(np.linalg.norm(w)))
    f_ref, p_ref = f_regression(xs, ys)
    assert np.allclose(f, f_ref, rtol=1e-10, atol=1e-12) and np.allclose(p, p_ref, rtol=1e-10, atol=1e-300)


def test_pca_splitter_logic_with_a_host_grm(g, monkeypatch):
    from tblup_b200 import splitter
    monkeypatch.setattr(splitter, "full_grm", lambda data, device=0, storage="packed2": O.exact_grm_block(
        O.exact_gram(data, np.arange(data.shape[1]), np.arange(data.shape[0])),
        *(lambda t: (t[0], t[0], t[1], t[2], t[3]))(O.exact_centring_terms(data, np.arange(data.shape[1]), np.arange(data.shape[0]),
                                                                            np.arange(data.shape[0])))))
    for tag in ("a", "b"):
        for outl in (0, 1):
            tr, te = splitter.pca_splitter(g["x_" + tag], outliers=bool(outl))
            assert tr == list(g["pca_train_%s_%d" % (tag, outl)]) and te == list(g["pca_test_%s_%d" % (tag, outl)])


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b"])
def test_full_grm_and_pca_split_on_the_gpu(g, tag):
    from tblup_b200 import splitter
    x = g["x_" + tag]
    grm = splitter.full_grm(x)
    assert np.abs(grm - g["grm_" + tag]).max() < 1e-11
    for outl in (0, 1):
        tr, te = splitter.pca_splitter(x.astype(np.float64), outliers=bool(outl))
        assert tr == list(g["pca_train_%s_%d" % (tag, outl)]) and te == list(g["pca_test_%s_%d" % (tag, outl)])


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b"])
@pytest.mark.parametrize("storage", ["packed2", "int8"])
def test_top_snps_ranking_on_the_gpu(g, tag, storage, tmp_path):
    from tblup_b200 import GblupEngine, seeder
    x, y = g["x_" + tag], g["y_" + tag]
    n_tr = len(g["seed_train_" + tag])
    with GblupEngine(x, y, storage=storage) as eng:
        order, scores = seeder.sorted_indices(eng, y, n_tr)
        # the sums themselves, exact against numpy
        rows = np.array([5, 3, 3, 40, 7])
        w = np.array([0.5, -1.25, 2.0, 3.0, -0.75])
        sx, sxx, sxw = eng.marker_stats(rows, w)
        xs = x[rows].astype(np.float64)
        assert np.array_equal(sx, xs.sum(0)) and np.array_equal(sxx, (xs * xs).sum(0)) and np.allclose(sxw, xs.T @ w, rtol=1e-14, atol=1e-14)
    ref_scores, ref_order = g["seed_scores_" + tag], g["seed_order_" + tag]
    assert np.allclose(scores, ref_scores, rtol=1e-9, atol=1e-300)
    # same ranking; markers with (numerically) tied scores may swap among themselves
    assert np.allclose(ref_scores[order], ref_scores[ref_order], rtol=1e-9, atol=1e-300)
    assert sorted(order.tolist()) == list(range(x.shape[1]))
    # the drop-in strategy class through files
    np.save(tmp_path / "geno.npy", x.astype(np.float64))
    np.save(tmp_path / "pheno.npy", y)

    class Ev:
        training_indices = list(g["seed_train_" + tag])

    strat = seeder.TopSNPsSeedStrategy(Ev(), None, str(tmp_path / "geno.npy"), str(tmp_path / "pheno.npy"))
    assert np.allclose(ref_scores[strat.indices], ref_scores[ref_order], rtol=1e-9, atol=1e-300)
    first = strat.get_next_indices(10)
    assert np.array_equal(first, strat.indices[:10]) and strat.current_index == 10
    strat.reset()
    assert strat.current_index == 0
