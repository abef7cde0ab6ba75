"""Parity of the i-vector back-end (PldaDev statistics / EFR / LDA / WCCN, cosine / Mahalanobis /
two-covariance scoring -- SURVEY §8f rank 3) against the fp64 restatement of PldaTools.cpp.
The reference ships no fixture for IvTest / IvNorm -> parity unpinned at the reference boundary."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from lia_ral_b200 import capi
    capi.init(0)
    return capi


def _dev_set(d, n_spk, seed, lo=2, hi=7):
    rng = np.random.default_rng(seed)
    cls = np.repeat(np.arange(n_spk), rng.integers(lo, hi, n_spk))
    spk = rng.standard_normal((d, n_spk)) * 1.5
    A = rng.standard_normal((d, d)) / np.sqrt(d) + np.eye(d)
    data = spk[:, cls] + A @ rng.standard_normal((d, len(cls))) + 0.3
    return np.ascontiguousarray(data), cls.astype(np.int32)


def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("d,n_spk", [(16, 40), (100, 150)], ids=["d16", "d100"])
def test_dev_statistics_and_normalisation(capi, oracle, d, n_spk):
    data, cls = _dev_set(d, n_spk, seed=3)
    mean, sm, S, W, B = capi.iv_cov_mat(data, cls, n_spk)
    mean_r, sm_r, S_r, W_r, B_r = oracle.iv_cov_mat(data, cls, n_spk)
    assert _rel(mean, mean_r) < 1e-13 and _rel(sm, sm_r) < 1e-13
    for got, ref in ((S, S_r), (W, W_r), (B, B_r)):
        assert _rel(got, ref) < 1e-11 and np.array_equal(got, got.T)
    assert _rel(capi.iv_wccn_chol(data, cls, n_spk), oracle.iv_wccn_chol(data, cls, n_spk)) < 1e-8
    assert _rel(capi.iv_mahalanobis_matrix(data, cls, n_spk), oracle.invert(W_r)) < 1e-9
    # EFR / sphNorm matrices: whitening property + agreement with the restatement (same sign rule)
    for cov in (S_r, W_r):
        E = capi.iv_efr_matrix(cov)
        assert np.allclose(E @ cov @ E.T, np.eye(d), atol=1e-9)
        assert _rel(E, oracle.iv_efr_matrix(cov)) < 1e-7
    # one full sphericalNuisanceNormalization iteration (:1822-1929): center, rotate, length-norm
    E = oracle.iv_efr_matrix(S_r)
    got = capi.iv_normalize(data, mu=mean_r, M=E, length_norm=True)
    ref = oracle.iv_length_norm(oracle.iv_rotate_left(E, oracle.iv_center(data, mean_r)))
    assert _rel(got, ref) < 1e-12 and np.allclose(np.linalg.norm(got, axis=0), 1.0)
    assert _rel(capi.iv_normalize(data, mu=mean_r), oracle.iv_center(data, mean_r)) < 1e-15
    assert _rel(capi.iv_normalize(data, length_norm=True), oracle.iv_length_norm(data)) < 1e-14
    # LDA: leading eigenvectors of W^-1 B
    rank = min(10, n_spk - 1, d)
    L, L_r = capi.iv_lda(W_r, B_r, rank), oracle.iv_lda(W_r, B_r, rank)
    assert L.shape == (rank, d) and np.allclose(np.linalg.norm(L, axis=1), 1.0)
    lam = np.sort(np.linalg.eigvals(np.linalg.solve(W_r, B_r)).real)[::-1]
    gap = np.abs(np.diff(lam[:rank + 1])).min() / lam[0]
    resid = np.linalg.solve(W_r, B_r) @ L.T - L.T * lam[:rank]
    assert np.abs(resid).max() < 1e-8 * lam[0]
    if gap > 1e-4:
        assert _rel(L, L_r) < 1e-6


@pytest.mark.parametrize("d,nm,nt", [(16, 7, 11), (100, 300, 257)], ids=["tiny", "d100"])
def test_scorings(capi, oracle, d, nm, nt):
    data, cls = _dev_set(d, 3 * d, seed=5)
    _, _, _, W, B = oracle.iv_cov_mat(data, cls, 3 * d)
    rng = np.random.default_rng(6)
    models, segments = rng.standard_normal((d, nm)) + 0.2, rng.standard_normal((d, nt)) - 0.1
    trials = rng.random((nm, nt)) < 0.7
    for tr in (None, trials):
        assert _rel(capi.iv_cosine_scoring(models, segments, tr), oracle.iv_cosine(models, segments, tr)) < 1e-12
        Mah = oracle.invert(W)
        got, ref = capi.iv_mahalanobis_scoring(models, segments, Mah, tr), oracle.iv_mahalanobis(models, segments, Mah, tr)
        assert _rel(got, ref) < 1e-11
        if tr is not None:
            assert np.all(got[~tr] == 0.0)
    # a non-symmetric "Mahalanobis" matrix loaded from a file is legal in the reference
    Mns = Mah + 0.05 * rng.standard_normal((d, d))
    assert _rel(capi.iv_mahalanobis_scoring(models, segments, Mns), oracle.iv_mahalanobis(models, segments, Mns)) < 1e-11
    got, ref = capi.iv_two_cov_scoring(models, segments, W, B), oracle.iv_two_cov(models, segments, W, B)
    assert _rel(got, ref) < 1e-8
    with pytest.raises(capi.LrError):
        capi.iv_two_cov_scoring(models, segments, W - 10 * np.eye(d), B)   # not positive definite


@pytest.mark.parametrize("d,rF,rG,n_spk", [(12, 5, 3, 40), (60, 20, 0, 200), (100, 30, 10, 300)],
                         ids=["small", "no-channel", "d100"])
def test_plda_em_iterations(capi, oracle, d, rF, rG, n_spk):
    """Three PldaModel::em_iteration steps (PldaTools.cpp:2329-2343, 2359-2485, 2790-2813) chained on
    the device against the restated loops: F, G, Sigma, the minimum-divergence mean and the centred
    development data after every iteration."""
    data, cls = _dev_set(d, n_spk, seed=11, lo=1, hi=6)
    rng = np.random.default_rng(12)
    F, G = rng.standard_normal((d, rF)), (rng.standard_normal((d, rG)) if rG else None)
    _, _, Sigma, _, _ = oracle.iv_cov_mat(data, cls, n_spk)
    data = data - data.mean(1, keepdims=True)          # PLDA.cpp: updateMean + centerData
    Delta = np.zeros(d)
    dev = (data, F, G, Sigma, Delta)
    ref = (data, F, G, Sigma, Delta)
    for it in range(3):
        dev = capi.plda_em_iteration(dev[0], cls, n_spk, *dev[1:])
        ref = oracle.plda_em_iteration(ref[0], cls, n_spk, *ref[1:])
        names = ("data", "F", "G", "Sigma", "Delta")
        for name, a, b in zip(names, dev, ref):
            if name == "G" and rG == 0:
                continue
            tol = 1e-12 if name == "data" else 1e-7
            assert np.abs(a - b).max() <= tol * max(np.abs(b).max(), 1e-3), (it, name)
    # the speaker subspace explains between-speaker variance: F F^T is no longer the random start
    assert np.linalg.norm(dev[1]) > 0 and np.all(np.isfinite(dev[3]))
    assert np.all(np.linalg.eigvalsh((dev[3] + dev[3].T) / 2) > 0)
