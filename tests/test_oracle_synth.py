"""lr_oracle.c vs the independent numpy formulation on seeded synthetic data, plus the
domain properties the reference relies on (EM monotone, threads == no threads)."""
import numpy as np
import pytest

from lia_ral_b200 import synth
from oracle import np_oracle
from oracle.ffi import Oracle


@pytest.fixture(scope="module")
def small():
    w, mean, cov = synth.make_ubm(C=32, D=12, seed=11)
    X = synth.make_frames(w, mean, cov, 700, seed=12)
    return w, mean, cov, X


def test_bwstats_matches_numpy(oracle, small):
    w, mean, cov, X = small
    g = oracle.gmm(w, mean, cov)
    U = 5
    f2r = (np.arange(len(X)) * U // len(X)).astype(np.int32)
    f2r[::17] = -1  # unselected frames
    N, F = oracle.bwstats(g, X, f2r, U)
    N2, F2 = np_oracle.bwstats(w, mean, cov, X, f2r, U)
    assert np.allclose(N, N2, rtol=1e-10, atol=1e-12)
    assert np.allclose(F, F2, rtol=1e-10, atol=1e-10)
    assert np.isclose(N.sum(), (f2r >= 0).sum())
    N3, F3 = oracle.bwstats(g, X, f2r, U, threads=3)
    assert np.array_equal(N, N3) and np.array_equal(F, F3)


def test_jfa_normalize_features_matches_numpy(oracle, small):
    """orc_jfa_normalize_features vs a log-domain numpy restatement of JFAAcc::normalizeFeatures
    (AccumulateJFAStat.cpp:4623-4680), with a frame range covered twice."""
    w, mean, cov, X = small
    cov = cov * 3.0
    ux = 0.3 * np.sqrt(cov) * np.random.default_rng(5).standard_normal(mean.shape)
    segs = [(10, 300), (250, 100), (600, 0)]
    Y = oracle.jfa_normalize_features(oracle.gmm(w, mean + ux, cov), ux, X, segs)
    Z = np.array(X, dtype=np.float32)
    for b, n in segs:
        for t in range(b, b + n):
            x = Z[t].astype(np.float64)
            ll = np.log(w) - 0.5 * np.log(2 * np.pi * cov).sum(1) - 0.5 * ((x - mean - ux) ** 2 / cov).sum(1)
            p = np.exp(ll - ll.max())
            Z[t] = (x - (p / p.sum()) @ ux).astype(np.float32)
    assert np.abs(Y - Z).max() <= 2e-7 * np.abs(X).max()
    assert np.array_equal(Y[350:], np.asarray(X, dtype=np.float32)[350:])
    assert np.abs(Y - X)[10:350].max() > 0.1


def test_em_matches_numpy_and_is_monotone(oracle, small):
    w, mean, cov, X = small
    w0, m0, c0 = synth.perturb_ubm(w, mean, cov, seed=13, frac=1.0, scale=0.5)
    g = oracle.gmm(w0, m0, c0)
    llk, n, occ, m1, m2 = oracle.em_accumulate(g, X)
    llk2, occ2, m12, m22 = np_oracle.em_stats(w0, m0, c0, X)
    assert n == len(X)
    assert np.isclose(llk, llk2, rtol=1e-12)
    assert np.allclose(occ, occ2, rtol=1e-10) and np.allclose(m1, m12, rtol=1e-9, atol=1e-10)
    assert np.allclose(m2, m22, rtol=1e-9, atol=1e-10)
    llk_t, _, occ_t, m1_t, m2_t = oracle.em_accumulate(g, X, threads=4)
    assert np.isclose(llk, llk_t, rtol=1e-13) and np.allclose(m2, m2_t, rtol=1e-12)
    gm, gc = oracle.mean_cov(X)
    prev = llk
    for it in range(4):
        wn, mn, cn = oracle.em_get(g, occ, m1, m2)
        cn, _, _ = oracle.variance_control(cn, 0.01, 100.0, gc)
        g = oracle.gmm(wn, mn, cn)
        cur, _, occ, m1, m2 = oracle.em_accumulate(g, X)
        assert cur >= prev - 1e-9
        prev = cur


def test_fast_build_agrees(small):
    w, mean, cov, X = small
    o, f = Oracle(), Oracle(fast=True)
    g = o.gmm(w, mean, cov)
    a = o.em_accumulate(g, X)
    b = f.em_accumulate(f.gmm(w, mean, cov), X, threads=2)
    assert np.isclose(a[0], b[0], rtol=1e-10) and np.allclose(a[3], b[3], rtol=1e-8, atol=1e-9)


def test_set_it_parameter_and_variance_control(oracle):
    assert oracle.set_it_parameter(0.5, 0.1, 1, 0) == 0.5
    assert np.isclose(oracle.set_it_parameter(0.5, 0.1, 5, 4), 0.1)
    assert np.isclose(oracle.set_it_parameter(0.5, 0.1, 5, 2), 0.3)
    cov = np.array([[0.1, 1.0, 50.0]])
    out, nf, nc = oracle.variance_control(cov, 0.5, 10.0, np.ones(3))
    assert np.allclose(out, [[0.5, 1.0, 10.0]]) and (nf, nc) == (1, 1)


def test_topk_semantics(oracle, small):
    w, mean, cov, X = small
    g = oracle.gmm(w, mean, cov)
    K = 5
    llk_c, idx, top_lk, rest, rest_w = oracle.llk_determine_top(g, X, K, complete=True)
    llk_p, idx2, _, _, _ = oracle.llk_determine_top(g, X, K, complete=False)
    full = oracle.llk_all(g, X)
    assert np.array_equal(idx, idx2)
    assert np.allclose(llk_c, full, rtol=1e-13)            # COMPLETE == all components
    assert (llk_p <= llk_c + 1e-12).all()
    assert (np.diff(top_lk, axis=1) <= 0).all()             # descending
    lj = np_oracle.log_joint(w, mean, cov, X)
    assert np.array_equal(np.argsort(-lj, axis=1, kind="stable")[:, :K], idx)
    assert np.allclose(rest_w, 1 - w[idx].sum(1))
    # clamp
    far = (X[:3] + 1e4).astype(np.float32)
    assert (oracle.llk_all(g, far) == -200.0).all()
    # client model through the stored top-K
    wc, mc, cc = synth.perturb_ubm(w, mean, cov, seed=5, frac=0.5)
    gcl = oracle.gmm(wc, mc, cc)
    llk_cl = oracle.llk_use_top(gcl, X, idx, rest, complete=True)
    ljc = np_oracle.log_joint(wc, mc, cc, X)
    expect = np.log(np.exp(np.take_along_axis(ljc, idx.astype(np.int64), 1)).sum(1) + rest)
    assert np.allclose(llk_cl, expect, rtol=1e-12)


@pytest.fixture(scope="module")
def tv_case():
    C, D, R, U = 16, 6, 10, 12
    w, mean, cov = synth.make_ubm(C=C, D=D, seed=21)
    N, F = synth.make_bw_stats(U, w, mean, cov, frames_per_utt=300, active=8, seed=22)
    invvar = (1.0 / cov).reshape(-1)
    T = synth.make_T(R, C, D, invvar, seed=23) * 20
    return C, D, R, U, w, mean, cov, N, F, invvar, T


def test_tv_ivector_and_estep(oracle, tv_case):
    C, D, R, U, w, mean, cov, N, F, invvar, T = tv_case
    Fc = oracle.tv_subtract_m(N, F, mean.reshape(-1))
    assert np.allclose(Fc, F - np.repeat(N, D, axis=1) * mean.reshape(-1)[None, :])
    tett = oracle.tv_tett(T, invvar, C, D)
    ref = np.stack([(T[:, c * D:(c + 1) * D] * invvar[c * D:(c + 1) * D]) @ T[:, c * D:(c + 1) * D].T
                    for c in range(C)])
    assert np.allclose(tett, ref, rtol=1e-12, atol=1e-14)
    assert np.array_equal(tett, oracle.tv_tett(T, invvar, C, D, threads=3))
    W = oracle.tv_ivectors(N, Fc, T, invvar, tett)
    W2, _ = np_oracle.ivectors(N, Fc, T, invvar)
    assert np.allclose(W, W2, rtol=1e-9, atol=1e-12)
    Wt = oracle.tv_ivectors(N, Fc, T, invvar, tett, threads=4)
    assert np.array_equal(W, Wt)
    We, A, Cmx, Rm, r, meanW = oracle.tv_estep(N, Fc, T, invvar, tett)
    W3, A3, C3, R3, r3, m3 = np_oracle.tv_estep(N, Fc, T, invvar)
    assert np.allclose(We, W3, rtol=1e-9, atol=1e-12)
    assert np.allclose(A, A3, rtol=1e-9, atol=1e-10) and np.allclose(Cmx, C3, rtol=1e-9, atol=1e-10)
    assert np.allclose(Rm, R3, rtol=1e-9) and np.allclose(r, r3) and np.allclose(meanW, m3)
    _, A4, C4, _, _, _ = oracle.tv_estep(N, Fc, T, invvar, tett, threads=3)
    assert np.allclose(A, A4, rtol=1e-12) and np.allclose(Cmx, C4, rtol=1e-12)
    Tn = oracle.tv_mstep(A, Cmx, C, D)
    assert np.allclose(Tn, np_oracle.tv_mstep(A, Cmx, C, D), rtol=1e-8, atol=1e-12)
    # Cmx is accumulated into, not reset (AccumulateTVStat.cpp:1784-1788 relies on resetTmpAcc)
    _, _, Cacc, _, _, _ = oracle.tv_estep(N, Fc, T, invvar, tett, Cmx=Cmx.copy())
    assert np.allclose(Cacc, 2 * Cmx)


def test_tv_mindiv_and_orthonormalize(oracle, tv_case):
    C, D, R, U, w, mean, cov, N, F, invvar, T = tv_case
    Fc = oracle.tv_subtract_m(N, F, mean.reshape(-1))
    tett = oracle.tv_tett(T, invvar, C, D)
    _, A, Cmx, Rm, r, meanW = oracle.tv_estep(N, Fc, T, invvar, tett)
    Tn = oracle.tv_mstep(A, Cmx, C, D)
    mean2, T2 = oracle.tv_mindiv(Rm, r, meanW, mean.reshape(-1), Tn, n_sessions=U, C=C, D=D)
    cov_y = Rm / U - np.outer(r / U, r / U)
    Ch = oracle.upper_cholesky(cov_y)
    assert np.allclose(Ch.T @ Ch, cov_y) and np.allclose(Ch, np.triu(Ch))
    assert np.allclose(T2, Ch @ Tn) and np.allclose(mean2, mean.reshape(-1) + meanW @ Tn)
    Q = oracle.tv_orthonormalize(T)
    assert np.allclose(Q @ Q.T, np.eye(R), atol=1e-8)
    a = np.random.default_rng(0).standard_normal((7, 7))
    a = a @ a.T + np.eye(7)
    assert np.allclose(oracle.invert(a) @ a, np.eye(7), atol=1e-10)


@pytest.mark.parametrize("rG", [0, 3])
def test_plda_scoring(oracle, rG):
    F, G, Sigma, models, model_of, segments = synth.make_plda(
        d=20, rF=6, rG=rG, n_test=9, sessions=[1, 1, 2, 3, 1, 2], seed=31)
    s = oracle.plda_native_scoring(F, G, Sigma, models, model_of, segments)
    s2 = np_oracle.plda_scores(F, G, Sigma, models, model_of, segments)
    assert s.shape == (6, 9)
    assert np.allclose(s, s2, rtol=1e-8, atol=1e-9)


def test_approximate_ivector_modes_against_numpy(oracle):
    """The restated approximate-extraction loops (AccumulateTVStat.cpp:1225-1242, 1600-1609,
    2348-2396, 2566-2609, 2837-2855, 2999-3052, 3116-3136) against independent numpy algebra."""
    rng = np.random.default_rng(0)
    C, D, R, U = 6, 4, 5, 7
    T = rng.standard_normal((R, C * D))
    invvar = rng.uniform(0.5, 2, C * D)
    w = rng.dirichlet(np.ones(C))
    N = rng.uniform(0, 5, (U, C))
    F = rng.standard_normal((U, C * D))
    mean = rng.standard_normal(C * D)
    Tn = oracle.tv_norm_t(T, invvar)
    assert np.allclose(Tn, T * np.sqrt(invvar))
    Fn = oracle.tv_norm_statistics(N, F, mean, invvar)
    assert np.allclose(Fn, (F - np.repeat(N, D, axis=1) * mean) * np.sqrt(invvar))
    Wc = oracle.tv_weighted_cov(Tn, w, C, D)
    assert np.allclose(Wc, (Tn * np.repeat(w, D)) @ Tn.T)
    Q, lam = oracle.eigen_sym(Wc)
    l2, Q2 = np.linalg.eigh(Wc)
    assert np.allclose(lam, l2[::-1]) and np.allclose(np.abs(Q), np.abs(Q2[:, ::-1]), atol=1e-10)
    assert all(Q[np.abs(Q[:, j]).argmax(), j] > 0 for j in range(R))   # documented sign convention
    Dm = oracle.tv_approximate_tctc(Tn, Q, C, D)
    ref = np.stack([np.diag(Q.T @ Tn[:, c * D:(c + 1) * D] @ Tn[:, c * D:(c + 1) * D].T @ Q) for c in range(C)])
    assert np.allclose(Dm, ref)
    W1 = oracle.tv_ivectors_ubm_weight(N, Fn, Tn, Wc)
    ref = np.stack([np.linalg.solve(np.eye(R) + N[s].sum() * Wc, Tn @ Fn[s]) for s in range(U)])
    assert np.allclose(W1, ref)
    W2 = oracle.tv_ivectors_eigen(N, Fn, Tn, Dm, Q)
    ref = np.stack([Q @ np.diag(1 / (1 + N[s] @ Dm)) @ Q.T @ (Tn @ Fn[s]) for s in range(U)])
    assert np.allclose(W2, ref)
    # the reference accumulates into _W in the eigenDecomposition estimator
    assert np.allclose(oracle.tv_ivectors_eigen(N, Fn, Tn, Dm, Q, W0=W1), W1 + W2)


def test_ivector_backend_against_numpy(oracle):
    """PldaDev statistics / normalisation matrices and the cosine / Mahalanobis / 2cov scorings
    (PldaTools.cpp:353-385, 527-571, 1124-1175, 1381-1415, 1853-1900, 3842-3910, 4083-4173)
    against independent numpy algebra."""
    rng = np.random.default_rng(1)
    d, nspk = 8, 12
    cls = np.repeat(np.arange(nspk), rng.integers(2, 6, nspk))
    n = len(cls)
    data = (rng.standard_normal((d, nspk)) * 1.5)[:, cls] + rng.standard_normal((d, n))
    mean, sm, S, W, B = oracle.iv_cov_mat(data, cls, nspk)
    assert np.allclose(mean, data.mean(1)) and np.allclose(S, np.cov(data, bias=True)) and np.allclose(S, W + B)
    wc = oracle.iv_wccn_chol(data, cls, nspk)
    Ww = np.mean([np.cov(data[:, cls == c], bias=True) for c in range(nspk)], axis=0)
    assert np.allclose(wc.T @ wc, np.linalg.inv(Ww)) and np.allclose(wc, np.triu(wc))
    E = oracle.iv_efr_matrix(S)
    assert np.allclose(E @ S @ E.T, np.eye(d))
    L = oracle.iv_lda(W, B, 3)
    ev = np.linalg.eig(np.linalg.inv(W) @ B)
    order = np.argsort(-ev[0].real)
    for j in range(3):
        v = ev[1][:, order[j]].real
        v /= np.linalg.norm(v)
        assert min(np.abs(L[j] - v).max(), np.abs(L[j] + v).max()) < 1e-8
    M, Sg = data[:, :5], data[:, 5:12]
    assert np.allclose(oracle.iv_cosine(M, Sg), (M.T @ Sg) / np.outer(np.linalg.norm(M, axis=0), np.linalg.norm(Sg, axis=0)))
    Mah = np.linalg.inv(W)
    ref = np.array([[-0.5 * (M[:, i] - Sg[:, j]) @ Mah @ (M[:, i] - Sg[:, j]) for j in range(7)] for i in range(5)])
    assert np.allclose(oracle.iv_mahalanobis(M, Sg, Mah), ref)
    iW, iB = np.linalg.inv(W), np.linalg.inv(B)
    G, H = iW @ np.linalg.inv(iB + 2 * iW) @ iW, iW @ np.linalg.inv(iB + iW) @ iW
    ref = np.array([[(M[:, i] + Sg[:, j]) @ G @ (M[:, i] + Sg[:, j]) - M[:, i] @ H @ M[:, i] - Sg[:, j] @ H @ Sg[:, j]
                     for j in range(7)] for i in range(5)])
    assert np.allclose(oracle.iv_two_cov(M, Sg, W, B), ref)
    tr = rng.random((5, 7)) < 0.5
    assert np.all(oracle.iv_cosine(M, Sg, tr)[~tr] == 0)


def test_plda_em_iteration_against_numpy(oracle):
    """The restated PldaModel::em_iteration (PldaTools.cpp:2329-2343, 2359-2485, 2790-2813) against an
    independent numpy formulation (per-speaker dense inverses instead of the eigenbasis trick)."""
    def em_np(data, cls, F, G, Sigma, Delta):
        X = data - Delta[:, None]
        d, n = X.shape
        rF, rG = F.shape[1], G.shape[1]
        obs = X @ X.T
        iS = np.linalg.inv(Sigma)
        Ftw, Gtw = F.T @ iS, G.T @ iS
        iGG = np.linalg.inv(Gtw @ G + np.eye(rG))
        FtwG = Ftw @ G
        S = iGG @ FtwG.T
        A = Ftw @ F - FtwG @ iGG @ FtwG.T
        Ehh, xh, U = np.zeros((rF + rG, rF + rG)), np.zeros((d, rF + rG)), np.zeros(rF + rG)
        for c in np.unique(cls):
            Xs = X[:, cls == c]
            ns = Xs.shape[1]
            M = np.linalg.inv(ns * A + np.eye(rF))
            Sx = iS @ Xs
            fi, gi = F.T @ Sx, G.T @ Sx
            eh = M @ (fi.sum(1) - S.T @ gi.sum(1))
            Eh = np.vstack([np.repeat(eh[:, None], ns, 1), iGG @ gi - (S @ eh)[:, None]])
            MsT = M @ S.T
            Ehh += ns * np.block([[M, -MsT], [-MsT.T, iGG + S @ MsT]]) + Eh @ Eh.T
            xh += Xs @ Eh.T
            U += Eh.sum(1)
        FG = xh @ np.linalg.inv(Ehh)
        Sig = (obs - FG @ xh.T) / n
        U /= n
        c = Ehh / n - np.outer(U, U)
        Fn = FG[:, :rF] @ np.linalg.cholesky(c[:rF, :rF])
        Gn = FG[:, rF:] @ np.linalg.cholesky(c[rF:, rF:]) if rG else np.zeros((d, 0))
        return X, Fn, Gn, Sig, Delta + FG @ U

    rng = np.random.default_rng(3)
    for d, rF, rG in ((10, 4, 3), (12, 5, 0)):
        nspk = 25
        cls = np.repeat(np.arange(nspk), rng.integers(1, 5, nspk)).astype(np.int32)
        n = len(cls)
        data = (rng.standard_normal((d, nspk)) * 1.2)[:, cls] + rng.standard_normal((d, n))
        a = b = (data, rng.standard_normal((d, rF)), rng.standard_normal((d, rG)),
                 np.cov(data, bias=True) + 0.1 * np.eye(d), 0.05 * rng.standard_normal(d))
        for it in range(3):
            a = oracle.plda_em_iteration(a[0], cls, nspk, a[1], a[2] if rG else None, a[3], a[4])
            b = em_np(b[0], cls, *b[1:])
            for x, y in zip(a, b):
                assert np.allclose(x, y, rtol=1e-8, atol=1e-10), it
