"""Parity of the Total-Variability / i-vector / PLDA device path against the fp64 oracle
(a literal restatement of AccumulateTVStat.cpp / PldaTools.cpp loops; the reference ships no
fixture for these programs -> parity unpinned at the reference boundary)."""
import numpy as np
import pytest

from lia_ral_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from lia_ral_b200 import capi
    capi.init(0)
    return capi


def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.fixture(scope="module", params=[(64, 12, 20, 37), (128, 20, 48, 150)], ids=["small", "medium"])
def tvcase(request, oracle):
    C, D, R, U = request.param
    w, mean, cov = synth.make_ubm(C, D, seed=21)
    invvar = (1.0 / cov).reshape(-1)
    N, F = synth.make_bw_stats(U, w, mean, cov, frames_per_utt=400, active=min(24, C), seed=22)
    T = synth.make_T(R, C, D, invvar, seed=23, scale=0.05)
    return dict(C=C, D=D, R=R, U=U, mean=mean.reshape(-1), invvar=invvar, N=N, F=F, T=T)


def _device(capi, c):
    tv = capi.TV(c["C"], c["D"], c["R"], c["U"], c["mean"], c["invvar"])
    tv.set_stats(c["N"], c["F"])
    tv.set_T(c["T"])
    return tv


def test_subtract_tett_ivectors(capi, oracle, tvcase):
    c = tvcase
    tv = _device(capi, c)
    tv.subtract_m()
    Fc_ref = oracle.tv_subtract_m(c["N"], c["F"], c["mean"])
    N, Fc = tv.get_stats()
    assert np.array_equal(N, c["N"]) and np.allclose(Fc, Fc_ref, rtol=1e-13, atol=1e-12)
    tv.estimate_tett()
    tv.estimate_w()
    tett = oracle.tv_tett(c["T"], c["invvar"], c["C"], c["D"])
    W_ref = oracle.tv_ivectors(c["N"], Fc_ref, c["T"], c["invvar"], tett)
    W = tv.get_W()
    assert _rel(W, W_ref) < 1e-9           # contract: 1e-4 relative
    # independent check of the oracle itself: dense solve of (I + sum_c N_c TETt_c) w = T S^-1 Fc
    L0 = np.eye(c["R"]) + np.tensordot(c["N"][0], tett, axes=1)
    w0 = np.linalg.solve(L0, c["T"] @ (c["invvar"] * Fc_ref[0]))
    assert np.allclose(W[0], w0, rtol=1e-8, atol=1e-10)


def test_em_iteration(capi, oracle, tvcase):
    """One full TotalVariability iteration (TotalVariability.cpp:118-153): substractM, TETt,
    estimateAandC, updateTestimate, minDivergence, orthonormalizeT."""
    c = tvcase
    C, D, R, U = c["C"], c["D"], c["R"], c["U"]
    tv = _device(capi, c)
    tv.reset_tmp_acc()
    tv.subtract_m()
    tv.estimate_tett()
    tv.estimate_a_and_c()
    Fc = oracle.tv_subtract_m(c["N"], c["F"], c["mean"])
    tett = oracle.tv_tett(c["T"], c["invvar"], C, D)
    W_r, A_r, Cmx_r, Rm_r, r_r, mw_r = oracle.tv_estep(c["N"], Fc, c["T"], c["invvar"], tett)
    A, Cmx, Rm, r, mw = tv.get_acc()
    assert _rel(tv.get_W(), W_r) < 1e-9
    assert _rel(A, A_r) < 1e-9 and _rel(Cmx, Cmx_r) < 1e-9
    assert _rel(Rm, Rm_r) < 1e-9 and _rel(r, r_r) < 1e-9 and _rel(mw, mw_r) < 1e-9
    # Cmx is NOT zeroed by estimateAandC (AccumulateTVStat.cpp:1719-1721): a second call doubles it
    tv.estimate_a_and_c()
    _, Cmx2, _, _, _ = tv.get_acc(want_A=False)
    assert _rel(Cmx2, 2.0 * Cmx_r) < 1e-9
    tv.reset_tmp_acc()
    tv.estimate_a_and_c()
    # M-step
    tv.update_t()
    T_r = oracle.tv_mstep(A_r, Cmx_r, C, D)
    T1 = tv.get_T()
    assert _rel(T1, T_r) < 1e-7
    # minimum divergence
    mean_r, T2_r = oracle.tv_mindiv(Rm_r, r_r, mw_r, c["mean"], T_r, float(U), C, D)
    tv.min_divergence(float(U))
    assert _rel(tv.get_T(), T2_r) < 1e-7 and _rel(tv.get_mean(), mean_r) < 1e-9
    # Gram-Schmidt
    tv.orthonormalize_t()
    T3_r = oracle.tv_orthonormalize(T2_r)
    T3 = tv.get_T()
    assert _rel(T3, T3_r) < 1e-6
    assert np.allclose(T3 @ T3.T, np.eye(R), atol=1e-6)


def test_em_increases_likelihood_proxy(capi, oracle, tvcase):
    """Three device EM iterations from the same statistics: the i-vectors explain the centred
    statistics better every iteration (the auxiliary function the M-step maximises)."""
    c = tvcase
    tv = _device(capi, c)
    Fc = oracle.tv_subtract_m(c["N"], c["F"], c["mean"])

    def fit(T, W):
        pred = (W @ T).reshape(c["U"], c["C"], c["D"]) * c["N"][:, :, None]
        res = Fc.reshape(c["U"], c["C"], c["D"]) - pred
        return float((res ** 2 * c["invvar"].reshape(c["C"], c["D"])[None]).sum())

    prev = None
    for it in range(3):
        tv.set_stats(c["N"], c["F"])  # the reference reloads N / F_X every iteration (:149-153)
        tv.reset_tmp_acc()
        tv.subtract_m()
        tv.estimate_tett()
        tv.estimate_a_and_c()
        tv.update_t()
        tv.estimate_tett()
        tv.estimate_w()
        cur = fit(tv.get_T(), tv.get_W())
        if prev is not None:
            assert cur <= prev * (1 + 1e-6)
        prev = cur


def test_not_positive_definite_is_reported(capi, tvcase):
    c = tvcase
    tv = _device(capi, c)
    N = c["N"].copy()
    N[:, 0] = 0.0  # component 0 never observed -> A_0 = 0 -> the M-step must fail loudly
    tv.set_stats(N, c["F"])
    tv.reset_tmp_acc()
    tv.subtract_m()
    tv.estimate_tett()
    tv.estimate_a_and_c()
    with pytest.raises(capi.LrError) as ei:
        tv.update_t()
    assert ei.value.code == 3


@pytest.mark.parametrize("rG,sessions", [(0, None), (5, None), (0, [1, 1, 3, 3, 2, 1]), (7, [2, 2, 2, 4])])
def test_plda_native_scoring(capi, oracle, rG, sessions):
    n_models = 23 if sessions is None else len(sessions)
    F, G, Sigma, models, model_of, segments = synth.make_plda(d=40, rF=12, rG=rG, n_models=n_models,
                                                              n_test=57, sessions=sessions, seed=31)
    ref = oracle.plda_native_scoring(F, G, Sigma, models, model_of, segments)
    got = capi.plda_native_scoring(F, G, Sigma, models, model_of, segments)
    assert got.shape == ref.shape
    # the trial matrix is a split-precision (fp16 hi / lo, 22 bits) tcgen05 GEMM: 1e-6 of the largest score
    assert np.abs(got - ref).max() < 1e-6 * max(1.0, np.abs(ref).max())


def test_plda_larger_tile(capi, oracle):
    F, G, Sigma, models, model_of, segments = synth.make_plda(d=100, rF=40, rG=0, n_models=300,
                                                              n_test=500, seed=32)
    ref = oracle.plda_native_scoring(F, G, Sigma, models, model_of, segments)
    got = capi.plda_native_scoring(F, G, Sigma, models, model_of, segments)
    assert np.abs(got - ref).max() < 1e-6 * np.abs(ref).max()


def test_plda_device_resident_fp32_and_sharded(capi, oracle):
    """lr_plda_native_scoring_dev: device operands, fp32 device scores, and two model shards scored
    independently (the e5 partitioning: models split across ranks, segments replicated) -- ragged sizes
    (not multiples of the 128-wide tiles), rank 200 like configs[4]."""
    import torch
    F, G, Sigma, models, model_of, segments = synth.make_plda(d=240, rF=200, rG=0, n_models=333, n_test=777,
                                                              seed=33)
    ref = oracle.plda_native_scoring(F, G, Sigma, models, model_of, segments)
    dm = torch.from_numpy(np.ascontiguousarray(models)).cuda()
    ds = torch.from_numpy(np.ascontiguousarray(segments)).cuda()
    out = torch.full((333, 780), float("nan"), dtype=torch.float32, device="cuda")   # ld 780 > n_test
    torch.cuda.synchronize()
    capi.plda_native_scoring_dev(F, G, Sigma, dm.data_ptr(), 333, model_of, ds.data_ptr(), 777, out.data_ptr(), 780)
    capi.synchronize()
    got = out.cpu().numpy()
    assert np.isnan(got[:, 777:]).all()                       # nothing written beyond n_test
    scale = np.abs(ref).max()
    assert np.abs(got[:, :777] - ref).max() < 2e-6 * scale
    # shards: rows [0, 200) and [200, 333) as two independent calls on column blocks of the models matrix
    for lo, hi in ((0, 200), (200, 333)):
        dsh = torch.from_numpy(np.ascontiguousarray(models[:, lo:hi])).cuda()
        o2 = torch.empty((hi - lo, 777), dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        capi.plda_native_scoring_dev(F, G, Sigma, dsh.data_ptr(), hi - lo, np.arange(hi - lo, dtype=np.int32),
                                     ds.data_ptr(), 777, o2.data_ptr())
        capi.synchronize()
        # (not bit-identical: each call scales its model operand by its own power of two before the split)
        assert np.abs(o2.cpu().numpy() - ref[lo:hi]).max() < 2e-6 * scale


def test_approximate_ivector_modes(capi, oracle, tvcase):
    """IvExtractor --mode ubmWeight / eigenDecomposition (IvExtractor.cpp:151-363): normTMatrix,
    normStatistics, getWeightedCov, computeEigenProblem, approximateTcTc and the two estimators
    against the literal restatement of AccumulateTVStat.cpp:1225-1242, 1600-1609, 2348-2396,
    2566-2609, 2837-2855, 2999-3052, 3116-3136."""
    c = tvcase
    C, D, R = c["C"], c["D"], c["R"]
    w, _, _ = synth.make_ubm(C, D, seed=21)
    tv = _device(capi, c)
    tv.norm_t()
    Tn_ref = oracle.tv_norm_t(c["T"], c["invvar"])
    assert _rel(tv.get_T(), Tn_ref) < 1e-14
    Wcov = tv.weighted_cov(w)
    Wcov_ref = oracle.tv_weighted_cov(Tn_ref, w, C, D)
    assert _rel(Wcov, Wcov_ref) < 1e-12 and np.array_equal(Wcov_ref, Wcov_ref.T)
    tv.norm_statistics()
    Fn_ref = oracle.tv_norm_statistics(c["N"], c["F"], c["mean"], c["invvar"])
    assert _rel(tv.get_stats()[1], Fn_ref) < 1e-13

    # ubmWeight: the device inverts L_s = I + n_s W in the eigenbasis of W, the oracle explicitly
    tv.estimate_w_ubm_weight(Wcov_ref)
    W_ref = oracle.tv_ivectors_ubm_weight(c["N"], Fn_ref, Tn_ref, Wcov_ref)
    assert _rel(tv.get_W(), W_ref) < 1e-9

    # eigen problem: eigenvalues descending, eigenvectors equal up to the documented sign convention
    Q, lam = capi.eigen_problem(Wcov_ref)
    Q_ref, lam_ref = oracle.eigen_sym(Wcov_ref)
    assert np.all(np.diff(lam) <= 0) and np.allclose(lam, lam_ref, rtol=1e-10, atol=1e-12 * lam_ref[0])
    assert np.allclose(Q.T @ Q, np.eye(R), atol=1e-10)
    assert np.allclose(Wcov_ref @ Q, Q * lam, atol=1e-9 * lam_ref[0])
    gap = np.abs(np.diff(lam_ref)).min() / lam_ref[0]
    if gap > 1e-6:   # well separated spectrum: the vectors themselves must agree
        assert np.allclose(Q, Q_ref, atol=1e-6 / gap * 1e-6)
    Dm = tv.approximate_tctc(Q)
    Dm_ref = oracle.tv_approximate_tctc(Tn_ref, Q, C, D)
    assert _rel(Dm, Dm_ref) < 1e-12

    # eigenDecomposition: accumulates into _W like the reference (W0 = the ubmWeight result above)
    W0 = tv.get_W()
    tv.estimate_w_eigen_decomposition(Dm_ref, Q)
    W2_ref = oracle.tv_ivectors_eigen(c["N"], Fn_ref, Tn_ref, Dm_ref, Q, W0=W0)
    assert _rel(tv.get_W(), W2_ref) < 1e-9
    # the approximation is exact when every T_c T_c^T shares the eigenvectors: sanity of the algebra
    i = 0
    dense = Q @ np.diag(1.0 / (1.0 + c["N"][i] @ Dm_ref)) @ Q.T @ (Tn_ref @ Fn_ref[i])
    assert np.allclose(W2_ref[i] - W0[i], dense, rtol=1e-9, atol=1e-12)


def test_packed_accumulator_roundtrip(capi, oracle, tvcase):
    """A_c is held as a packed lower triangle on the device (half the GEMM and half the all-reduce);
    lr_tv_get_acc must hand back the reference's full symmetric [C x R*R] layout."""
    c = tvcase
    tv = _device(capi, c)
    tv.reset_tmp_acc()
    tv.subtract_m()
    tv.estimate_tett()
    tv.estimate_a_and_c()
    A, Cmx, Rm, r, mw = tv.get_acc()
    R, C = c["R"], c["C"]
    A3 = A.reshape(C, R, R)
    assert np.array_equal(A3, A3.transpose(0, 2, 1))
    assert tv.acc_len() == C * R * (R + 1) // 2 + R * C * c["D"] + R * R + 2 * R
    Fc = oracle.tv_subtract_m(c["N"], c["F"], c["mean"])
    tett = oracle.tv_tett(c["T"], c["invvar"], C, c["D"])
    _, A_ref, _, _, _, _ = oracle.tv_estep(c["N"], Fc, c["T"], c["invvar"], tett)
    assert _rel(A, A_ref) < 1e-9
