"""Parity of the CUDA frames x components path (through the C ABI) against the fp64 oracle.

Tolerances (BASELINE.json north_star): top-Gaussian indices bit-exact; log-likelihoods within
1e-4 relative; statistics within 1e-4 relative (we hold them to much tighter bounds below so a
precision regression shows up long before the contract is at risk).
"""
import os

import numpy as np
import pytest

from lia_ral_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from lia_ral_b200 import capi
    capi.init(0)
    return capi


KERNELS = {"simt": 1, "tc": 2, "tc2p": 3}


@pytest.fixture(params=["simt", "tc", "tc2p"])
def kern(request, capi):
    """Run the test under the fp32 SIMT kernels, the tcgen05 one-pass statistics kernel and the
    tcgen05 two-pass kernels (forced, so a silent fall back to another path is impossible)."""
    capi.set_gmm_kernel(KERNELS[request.param])
    yield request.param
    capi.set_gmm_kernel(0)


# (rtol, atol relative to max|ref|) per kernel: the tcgen05 path rounds posteriors to fp16
# (11 bits) before the statistics GEMM, the same rounded posterior feeding N and F.
STAT_TOL = {"simt": (1e-4, 1e-7), "tc": (1e-4, 5e-5), "tc2p": (1e-4, 5e-5)}
LLK_ABS = {"simt": 2e-4, "tc": 1e-3, "tc2p": 1e-3}


def _relmax(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def _close(a, b, rtol, atol_rel):
    """elementwise |a-b| <= rtol |b| + atol_rel * max|b|"""
    return bool((np.abs(a - b) <= rtol * np.abs(b) + atol_rel * np.abs(b).max()).all())


@pytest.fixture(scope="module", params=[(2048, 60, 2500), (64, 60, 1500), (100, 13, 777)],
                ids=["2048c60d", "64c60d", "100c13d"])
def case(request):
    C, D, T = request.param
    w, mean, cov = synth.make_ubm(C=C, D=D, seed=1)
    X = synth.make_frames(w, mean, cov, T, seed=2)
    # a test model that is NOT the generator: posteriors spread over several components
    w2, m2, c2 = synth.perturb_ubm(w, mean, cov * 4.0, seed=3, frac=0.5, scale=1.0)
    return dict(C=C, D=D, T=T, w=w2, mean=m2, cov=c2, X=X)


def test_compute_all(capi, oracle, case):
    g = capi.GMM(case["w"], case["mean"], case["cov"])
    o = oracle.gmm(case["w"], case["mean"], case["cov"])
    got = g.get()
    assert np.allclose(got["covinv"], o.covinv, rtol=1e-15)
    assert np.allclose(got["det"], o.det, rtol=1e-12)
    assert np.allclose(got["cst"], o.cst, rtol=1e-12)


def test_llk_all(capi, oracle, case, kern):
    g = capi.GMM(case["w"], case["mean"], case["cov"])
    o = oracle.gmm(case["w"], case["mean"], case["cov"])
    ref = oracle.llk_all(o, case["X"], -1e9, 1e9)
    got = g.llk(case["X"], -1e9, 1e9)
    assert np.abs(got - ref).max() < 1e-4 * np.abs(ref).max()  # contract
    assert np.abs(got - ref).max() < LLK_ABS[kern]              # what the kernel delivers
    # clamp (minLLK / maxLLK of the shipped configs)
    lo, hi = np.percentile(ref, 30), np.percentile(ref, 70)
    got = g.llk(case["X"], lo, hi)
    assert np.allclose(got, np.clip(ref, lo, hi), atol=LLK_ABS[kern])


def _segments(T, U, rng):
    """ragged segments; some frames unselected, one file listed on two NDX lines"""
    cuts = np.sort(rng.choice(np.arange(1, T), size=3 * U, replace=False))
    segs, f2r = [], np.full(T, -1, dtype=np.int32)
    prev = 0
    for i, c in enumerate(cuts):
        if i % 5 != 4:  # every fifth run of frames is not selected by any label
            row = i % U
            segs.append((prev, c - prev, row))
            f2r[prev:c] = row
        prev = c
    return segs, f2r


def test_bwstats(capi, oracle, case, kern):
    rt, at = STAT_TOL[kern]
    C, D, T = case["C"], case["D"], case["T"]
    g = capi.GMM(case["w"], case["mean"], case["cov"])
    o = oracle.gmm(case["w"], case["mean"], case["cov"])
    U = 7
    segs, f2r = _segments(T, U, np.random.default_rng(5))
    N_ref, F_ref = oracle.bwstats(o, case["X"], f2r, U)
    N, F = g.bwstats(case["X"], segs, U)
    assert _close(N, N_ref, rt, at), _relmax(N, N_ref)
    assert _close(F, F_ref, rt, at), _relmax(F, F_ref)
    assert abs(N.sum() - (f2r >= 0).sum()) < (1e-3 if kern == "simt" else 1e-4 * T)
    # centred statistics F - mu N (substractM) are what the i-vector solve consumes
    Fc = F.reshape(U, C, D) - N[:, :, None] * case["mean"][None]
    Fc_ref = F_ref.reshape(U, C, D) - N_ref[:, :, None] * case["mean"][None]
    assert _relmax(Fc, Fc_ref) < 1e-4
    # accumulate-into semantics + a file listed on two NDX lines (AccumulateTVStat.cpp:339-346)
    extra = [(0, 200, 0), (0, 200, 3)]
    N2, F2 = g.bwstats(case["X"], extra, U, N=N.copy(), F=F.copy())
    f2 = np.full(T, -1, dtype=np.int32)
    f2[:200] = 0
    Na, Fa = oracle.bwstats(o, case["X"], f2, U)
    f2[:200] = 3
    Nb, Fb = oracle.bwstats(o, case["X"], f2, U)
    assert _close(N2, N_ref + Na + Nb, rt, at)
    assert _close(F2, F_ref + Fa + Fb, rt, at)


def test_bwstats_strided_and_empty(capi, oracle, case, kern):
    rt, at = STAT_TOL[kern]
    D, T = case["D"], case["T"]
    g = capi.GMM(case["w"], case["mean"], case["cov"])
    o = oracle.gmm(case["w"], case["mean"], case["cov"])
    wide = np.zeros((T, D + 3), dtype=np.float32)  # featureServerMask-style view: ldx > D
    wide[:, :D] = case["X"]
    Xv = wide[:, :D]
    f2r = np.zeros(T, dtype=np.int32)
    f2r[400:] = -1
    N_ref, F_ref = oracle.bwstats(o, Xv, f2r, 2)
    N, F = g.bwstats(Xv, [(0, 400, 0), (500, 0, 1)], 2)
    assert _close(N, N_ref, rt, at) and _close(F, F_ref, rt, at)
    assert not N[1].any() and not F[1].any()
    N, F = g.bwstats(Xv, [], 2)
    assert not N.any() and not F.any()


def test_em_accumulate_and_update(capi, oracle, case, kern):
    rt, at = STAT_TOL[kern]
    C, D = case["C"], case["D"]
    g = capi.GMM(case["w"], case["mean"], case["cov"])
    o = oracle.gmm(case["w"], case["mean"], case["cov"])
    X = case["X"]
    llk_ref, n_ref, occ_ref, m1_ref, m2_ref = oracle.em_accumulate(o, X, weight=0.5)
    llk, n, occ, m1, m2 = g.em_accumulate(X, weight=0.5)
    assert n == n_ref
    assert abs(llk - llk_ref) < 1e-5 * abs(llk_ref)
    assert _close(occ, occ_ref, rt, at)
    assert _close(m1, m1_ref, rt, at)
    assert _close(m2, m2_ref, rt, at)
    # getEM + varianceControl on the device vs the oracle on the ORACLE's statistics
    gm, gc = oracle.mean_cov(X)
    fl, ce = 0.3, 3.0
    w_r, mu_r, cv_r = oracle.em_get(o, occ_ref, m1_ref, m2_ref)
    cv_r, nf, nc = oracle.variance_control(cv_r, fl, ce, gc)
    g.em_update(occ_ref, m1_ref, m2_ref, fl, ce, gc)
    got = g.get()
    assert np.allclose(got["w"], w_r, rtol=1e-12)
    assert np.allclose(got["mean"], mu_r, rtol=1e-12, atol=1e-13)
    assert np.allclose(got["cov"], cv_r, rtol=1e-10, atol=1e-12)
    o2 = oracle.gmm(w_r, mu_r, cv_r)
    assert np.allclose(got["cst"], o2.cst, rtol=1e-10)
    # variances re-estimated from the DEVICE statistics (m2/occ - mu^2 cancels digits)
    w_d, mu_d, cv_d = oracle.em_get(o, occ, m1, m2)
    w_o, mu_o, cv_o = oracle.em_get(o, occ_ref, m1_ref, m2_ref)
    heavy = occ_ref > 1.0
    if heavy.any():
        assert _relmax(cv_d[heavy], cv_o[heavy]) < (1e-4 if kern == "simt" else 1e-3)
        assert _relmax(mu_d[heavy], mu_o[heavy]) < (1e-5 if kern == "simt" else 2e-4)
    # segments: only the listed frames, each once
    segs = [(10, 300), (700, 55)]
    sel = np.r_[10:310, 700:755]
    llk_ref, n_ref, occ_ref, m1_ref, m2_ref = oracle.em_accumulate(o, np.ascontiguousarray(X[sel]))
    g_fresh = capi.GMM(case["w"], case["mean"], case["cov"])  # g was re-estimated above
    llk, n, occ, m1, m2 = g_fresh.em_accumulate(X, segs=segs)
    assert n == len(sel) and abs(llk - llk_ref) < 1e-5 * abs(llk_ref)
    assert _close(m2, m2_ref, rt, at)


def test_mean_cov(capi, oracle, case):
    m_ref, c_ref = oracle.mean_cov(case["X"])
    m, c = capi.frames_mean_cov(case["X"])
    assert np.allclose(m, m_ref, rtol=1e-12, atol=1e-13) and np.allclose(c, c_ref, rtol=1e-11)


def test_topk_indices_bit_exact(capi, oracle, case):
    g = capi.GMM(case["w"], case["mean"], case["cov"])
    o = oracle.gmm(case["w"], case["mean"], case["cov"])
    X = case["X"]
    for K, complete in ((10, True), (1, False), (min(20, case["C"]), True)):
        llk_r, idx_r, top_r, rest_r, restw_r = oracle.llk_determine_top(o, X, K, complete)
        llk, idx, top, rest, restw = g.llk_topk(X, K, complete)
        assert np.array_equal(idx, idx_r)                     # bit-exact index selection
        assert np.allclose(top, top_r, rtol=1e-12, atol=0)    # fp64 re-evaluation
        assert np.allclose(restw, restw_r, rtol=0, atol=1e-14)
        assert np.allclose(llk, llk_r, rtol=0, atol=1e-4 * np.abs(llk_r).max())
        tot_r = top_r.sum(1) + rest_r
        assert np.allclose(top.sum(1) + rest, tot_r, rtol=1e-4)
        big = rest_r > 1e-6 * tot_r
        assert np.allclose(rest[big], rest_r[big], rtol=1e-4)


def test_use_topk_and_compute_test(capi, oracle, case):
    C, D, T = case["C"], case["D"], case["T"]
    X = case["X"]
    world = capi.GMM(case["w"], case["mean"], case["cov"])
    o_w = oracle.gmm(case["w"], case["mean"], case["cov"])
    K = 10
    cl_params = [synth.perturb_ubm(case["w"], case["mean"], case["cov"], seed=30 + i, frac=0.3,
                                   scale=0.3) for i in range(3)]
    clients = [capi.GMM(*p) for p in cl_params]
    o_cl = [oracle.gmm(*p) for p in cl_params]
    llk_w, idx, _, rest, _ = oracle.llk_determine_top(o_w, X, K, True)
    for complete in (True, False):
        ref = oracle.llk_use_top(o_cl[0], X, idx, rest, complete)
        got = clients[0].llk_use_topk(X, idx, rest, complete)
        assert np.allclose(got, ref, rtol=1e-12, atol=1e-12)
    a, b, c = T // 4, T // 2, T // 5
    segs = [(0, a), (b, c)]
    for per_segment in (False, True):
        mw, mc = capi.compute_test(world, clients, X, segs=segs, K=K, complete=True,
                                   per_segment=per_segment)
        groups = [np.r_[0:a], np.r_[b:b + c]] if per_segment else [np.r_[0:a, b:b + c]]
        for o_i, sel in enumerate(groups):
            assert abs(mw[o_i] - llk_w[sel].mean()) < 1e-4 * abs(llk_w[sel].mean())
            for i, oc in enumerate(o_cl):
                ref = oracle.llk_use_top(oc, X, idx, rest, True)[sel].mean()
                # LLR = client - world (ComputeTest.cpp:196-199)
                assert abs((mc[i, o_i] - mw[o_i]) - (ref - llk_w[sel].mean())) < 2e-4
    # reference fixture property (ComputeTest/test/test1.validate.res): client == world => LLR 0
    mw, mc = capi.compute_test(world, [world], X, K=K, complete=True)
    assert abs(mc[0, 0] - mw[0]) < 1e-12
    # worldDecime (ComputeTest.cpp:111-113, 162-165): inside every segment only frames idxFrame % 3 == 0 pick
    # the top list; the others reuse it (and its COMPLETE rest) for the world and for the clients
    dec = 3
    src = np.arange(T)
    for (b0, n) in segs:
        r = np.arange(n)
        src[b0:b0 + n] = b0 + r - r % dec
    idx_d, rest_d = idx[src], rest[src]
    on_grid = src == np.arange(T)
    llk_wd = np.where(on_grid, llk_w, oracle.llk_use_top(o_w, X, idx_d, rest_d, True))
    mw, mc = capi.compute_test(world, clients, X, segs=segs, K=K, complete=True, per_segment=True,
                               world_decime=dec)
    for o_i, sel in enumerate([np.r_[0:a], np.r_[b:b + c]]):
        assert abs(mw[o_i] - llk_wd[sel].mean()) < 1e-4 * abs(llk_wd[sel].mean())
        for i, oc in enumerate(o_cl):
            ref = oracle.llk_use_top(oc, X, idx_d, rest_d, True)[sel].mean()
            assert abs((mc[i, o_i] - mw[o_i]) - (ref - llk_wd[sel].mean())) < 2e-4
    assert abs(mw[0] - llk_w[0:a].mean()) > 1e-9   # decimation changes the world score


def test_gmmtokenizer_golden_on_gpu(capi, golden_dir):
    """The reference's own KAT (LIA_Utils/GmmTokenizer/test) through the CUDA path."""
    z = np.load(os.path.join(golden_dir, "gmmtokenizer.npz"))
    g = capi.GMM(z["w"], z["mean"], 1.0 / z["covinv"])
    g.set_cst(z["cst"])  # RAW files carry their own cst records
    X = np.ascontiguousarray(z["frames"][z["selected"]], dtype=np.float32)
    _, idx, _, _, _ = g.llk_topk(X, 6)
    best = idx[:, 0].astype(np.int64)
    collapsed = [best[0]] + [b for a, b in zip(best[:-1], best[1:]) if a != b]
    assert collapsed == list(z["sym_ref"])
    K = int(z["mce_topk"])
    g2 = capi.GMM(z["w"], z["mean"], 1.0 / z["covinv"])
    _, idx, _, _, _ = g2.llk_topk(X, K)
    mce = np.zeros_like(z["mce_ref"])
    for t in range(idx.shape[0]):
        for i in range(K):
            mce[idx[t, 0], idx[t, i]] += 1
    assert (mce == z["mce_ref"]).all()


def test_errors_are_loud(capi, case):
    g = capi.GMM(case["w"], case["mean"], case["cov"])
    with pytest.raises(capi.LrError):
        g.bwstats(case["X"], [(0, case["T"] + 1, 0)], 1)      # segment past the end
    with pytest.raises(capi.LrError):
        g.bwstats(case["X"], [(0, 10, 5)], 2)                 # row outside U
    with pytest.raises(capi.LrError):
        g.llk_topk(case["X"], case["C"] + 1)
    with pytest.raises(capi.LrError):
        capi.GMM(np.ones(4) / 4, np.zeros((4, 100)), np.ones((4, 100)))  # D > 63


def test_full_size_properties(capi, kern):
    """BASELINE config sizes (2048c/60d) where the oracle would take minutes: size-independent
    properties -- occupancies sum to the frame count, statistics are additive over a split of
    the frames, and the device-resident path equals the host path."""
    import torch
    w, mean, cov = synth.make_ubm(2048, 60, seed=1)
    T = 300_000
    X = synth.make_frames(w, mean, cov, T, seed=9)
    g = capi.GMM(w, mean, cov)
    llk, n, occ, m1, m2 = g.em_accumulate(X)
    assert n == T and abs(occ.sum() - T) < 1e-6 * T
    la, _, oa, m1a, m2a = g.em_accumulate(X[:123_457])
    lb, _, ob, m1b, m2b = g.em_accumulate(X[123_457:])
    assert abs((la + lb) - llk) < 1e-9 * abs(llk)
    # a different split moves the chunk boundaries of the fp32 partial sums: 1e-5, not 1e-9
    assert np.allclose(oa + ob, occ, rtol=2e-5, atol=1e-5)
    assert np.allclose(m2a + m2b, m2, rtol=2e-5, atol=1e-4)
    # device-resident frames + device statistics
    xd = torch.from_numpy(X).cuda()
    feats = capi.Feats(device_ptr=xd.data_ptr(), T=T, ldx=60, D=60)
    stats = torch.zeros(g.em_stats_len(), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    g.em_accumulate_dev(feats, 0, T, 1.0, stats.data_ptr())
    capi.synchronize()
    s = stats.cpu().numpy()
    assert np.allclose(s[:2048], occ, rtol=2e-5, atol=1e-5)
    assert np.allclose(s[2048:2048 + 2048 * 60], m1.reshape(-1), rtol=2e-5, atol=1e-4)
    assert abs(s[-2] - llk) < 1e-9 * abs(llk) and s[-1] == T
    # BW statistics: N row sums = frames per row, and F / N recovers the data mean per component
    segs = [(i * 3000, 3000, i) for i in range(100)]
    N, F = g.bwstats(X, segs, 100)
    assert np.allclose(N.sum(1), 3000.0, rtol=1e-6)


def test_tc_range_guard_and_outliers(capi, oracle):
    """Models whose normalised weights leave the fp16 range are refused by the tcgen05 path (and
    served by the SIMT path in auto mode); absurd outlier frames give finite results."""
    C, D, T = 256, 20, 640
    w, mean, cov = synth.make_ubm(C, D, seed=7)
    X = synth.make_frames(w, mean, cov, T, seed=8)
    g_ok = capi.GMM(w, mean, cov)
    bad_cov = cov.copy()
    bad_mean = mean.copy()
    bad_cov[3] = 1e-7          # a needle component far from the global mean: |K_c| >> 30000
    bad_mean[3] = 40.0
    g_bad = capi.GMM(w, bad_mean, bad_cov)
    o_bad = oracle.gmm(w, bad_mean, bad_cov)
    capi.set_gmm_kernel(2)
    try:
        g_ok.llk(X, -1e9, 1e9)                      # in range: served
        with pytest.raises(capi.LrError):
            g_bad.llk(X, -1e9, 1e9)                 # out of range: loud, never silently wrong
    finally:
        capi.set_gmm_kernel(0)
    got = g_bad.llk(X, -1e9, 1e9)                   # auto -> SIMT
    ref = oracle.llk_all(o_bad, X, -1e9, 1e9)
    assert np.abs(got - ref).max() < 1e-4 * np.abs(ref).max()
    # outliers: one frame 1e4 sigma away, one NaN-free huge negative; statistics stay finite
    Xo = X.copy()
    Xo[5] = 1e4
    Xo[6] = -3e3
    for k in (1, 2):
        capi.set_gmm_kernel(k)
        try:
            llk, n, occ, m1, m2 = g_ok.em_accumulate(Xo)
        finally:
            capi.set_gmm_kernel(0)
        assert np.isfinite(llk) and np.isfinite(occ).all() and np.isfinite(m1).all() and np.isfinite(m2).all()
        assert abs(occ.sum() - T) < 1e-3 * T


@pytest.mark.parametrize("C,D", [(1, 1), (3, 2), (257, 60), (128, 63)])
def test_shape_edges(capi, oracle, C, D):
    """degenerate / boundary shapes: single component, single dimension, C just above a slice,
    the largest supported vectSize; T = 1 and T not a multiple of any tile."""
    w, mean, cov = synth.make_ubm(C, D, seed=17)
    for T in (1, 131):
        X = synth.make_frames(w, mean, cov * 3.0, T, seed=18)
        g, o = capi.GMM(w, mean, cov), oracle.gmm(w, mean, cov)
        ref = oracle.llk_all(o, X, -1e9, 1e9)
        got = g.llk(X, -1e9, 1e9)
        assert np.abs(got - ref).max() < 1e-4 * max(1.0, np.abs(ref).max())
        llk_r, n_r, occ_r, m1_r, m2_r = oracle.em_accumulate(o, X)
        llk, n, occ, m1, m2 = g.em_accumulate(X)
        assert n == T and abs(llk - llk_r) < 1e-4 * max(1.0, abs(llk_r))
        assert np.allclose(occ, occ_r, rtol=1e-4, atol=1e-4) and np.allclose(m2, m2_r, rtol=1e-4, atol=1e-4 * np.abs(m2_r).max())
        K = min(3, C)
        _, idx_r, _, _, _ = oracle.llk_determine_top(o, X, K)
        _, idx, _, _, _ = g.llk_topk(X, K)
        assert np.array_equal(idx, idx_r)


def test_block_boundaries(capi):
    """frame counts straddling the staging block (2^18) and device block (2^21) sizes: the totals
    must not depend on where the blocks are cut."""
    import torch
    w, mean, cov = synth.make_ubm(256, 20, seed=19)
    g = capi.GMM(w, mean, cov)
    T = (1 << 21) + (1 << 18) + 777
    xd = torch.randn(T, 20, device="cuda") * 2.0
    X = xd.cpu().numpy()
    feats = capi.Feats(device_ptr=xd.data_ptr(), T=T, ldx=20, D=20)
    stats = torch.zeros(g.em_stats_len(), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    g.em_accumulate_dev(feats, 0, T, 1.0, stats.data_ptr())
    capi.synchronize()
    s = stats.cpu().numpy()
    llk, n, occ, m1, m2 = g.em_accumulate(X)          # host path, different block cuts
    assert n == T == s[-1] and abs(occ.sum() - T) < 1e-6 * T
    assert abs(s[-2] - llk) < 1e-9 * abs(llk)
    # fp32 TMEM partial sums run over up to 128 tiles (16 k frames) before the fp64 flush: different
    # block cuts regroup them, so agreement is ~3e-6, not 1e-9 (contract: 1e-4)
    assert np.allclose(s[:256], occ, rtol=2e-5, atol=1e-5)
    assert np.allclose(s[256:256 + 256 * 20], m1.ravel(), rtol=2e-5, atol=2e-5 * np.abs(m1).max())


def test_resident_em_operand_cache(capi, oracle):
    """EM iterations over resident frames reuse the frames' converted tensor-core operand (kept with the
    lr_feats handle; the normalised space is frozen while the mixture's global moments stay close).
    Three chained device iterations against the oracle; a second handle without history must give the
    same statistics; a changed buffer needs invalidate()."""
    import torch
    C, D, T = 256, 20, 40000
    w, mean, cov = synth.make_ubm(C, D, seed=23)
    X = synth.make_frames(w, mean, cov * 1.5, T, seed=24)
    start = synth.perturb_ubm(w, mean, cov, seed=25, frac=1.0, scale=0.3)
    xd = torch.tensor(X, device="cuda")
    feats = capi.Feats(device_ptr=xd.data_ptr(), T=T, ldx=D, D=D)
    g = capi.GMM(*start)
    go = oracle.gmm(*start)
    n = g.em_stats_len()
    stats = torch.zeros(n, dtype=torch.float64, device="cuda")
    for it in range(3):
        stats.zero_()
        torch.cuda.synchronize()
        g.em_accumulate_dev(feats, 0, T, 1.0, stats.data_ptr())
        capi.synchronize()
        s = stats.cpu().numpy()
        llk_r, n_r, occ_r, m1_r, m2_r = oracle.em_accumulate(go, X)
        assert abs(s[-2] - llk_r) < 1e-5 * abs(llk_r)
        assert np.abs(s[:C] - occ_r).max() < 1e-4 * occ_r.max()
        assert np.abs(s[C:C + C * D] - m1_r.ravel()).max() < 1e-4 * np.abs(m1_r).max()
        # a handle without history (converts afresh with the model's current normalisation)
        fresh = capi.Feats(device_ptr=xd.data_ptr(), T=T, ldx=D, D=D)
        stats2 = torch.zeros(n, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        g.em_accumulate_dev(fresh, 0, T, 1.0, stats2.data_ptr())
        capi.synchronize()
        # (fp64 RED adds land in a different order from run to run: equal to rounding, not bit for bit)
        assert torch.allclose(stats, stats2, rtol=1e-10, atol=1e-9)
        fresh.close()
        g.em_update_dev(stats.data_ptr())
        capi.synchronize()
        wn, mn, cn = oracle.em_get(go, occ_r, m1_r, m2_r)
        go = oracle.gmm(wn, mn, cn)
    # the buffer changes under the handle: stale until invalidate()
    xd.mul_(1.05)
    torch.cuda.synchronize()
    feats.invalidate()
    stats.zero_()
    torch.cuda.synchronize()
    g.em_accumulate_dev(feats, 0, T, 1.0, stats.data_ptr())
    capi.synchronize()
    llk_r, _, occ_r, _, _ = oracle.em_accumulate(go, np.ascontiguousarray(xd.cpu().numpy()))
    s = stats.cpu().numpy()
    assert abs(s[-2] - llk_r) < 1e-5 * abs(llk_r) and np.abs(s[:C] - occ_r).max() < 1e-4 * occ_r.max()


@pytest.mark.parametrize("kern", [1, 0], ids=["simt", "auto"])
def test_jfa_bwstats(capi, oracle, kern):
    """JFAAcc::computeAndAccumulateJFAStat (AccumulateJFAStat.cpp:520-576): per-session and per-speaker
    accumulators from one pass; several sessions per speaker, a session in two segments, += semantics."""
    C, D, T = 256, 20, 6000
    w, mean, cov = synth.make_ubm(C, D, seed=71)
    X = synth.make_frames(w, mean, cov * 2.0, T, seed=72)
    capi.set_gmm_kernel(kern)
    try:
        g, o = capi.GMM(w, mean, cov), oracle.gmm(w, mean, cov)
        # 5 sessions, 3 speakers; session 1 is made of two segments; frames 5500.. are in no segment
        segs = [(0, 1000, 0), (1000, 700, 1), (3000, 400, 1), (1700, 1300, 2), (3400, 1100, 3), (4500, 1000, 4)]
        spk = [0, 0, 1, 2, 2]
        N_h, F_h, N, F = g.jfa_bwstats(X, segs, spk, 3)
        f2r = np.full(T, -1, dtype=np.int32)
        for b, n, r in segs:
            f2r[b:b + n] = r
        sel = f2r >= 0
        Nh_r, Fh_r = oracle.bwstats(o, np.ascontiguousarray(X[sel]), f2r[sel], 5)
        tol = 1e-4 if kern == 0 else 1e-5
        assert np.abs(N_h - Nh_r).max() <= tol * np.abs(Nh_r).max()
        assert np.abs(F_h - Fh_r).max() <= tol * np.abs(Fh_r).max()
        for s_ in range(3):
            rows = [h for h in range(5) if spk[h] == s_]
            assert np.allclose(N[s_], N_h[rows].sum(0), rtol=1e-12, atol=1e-12)
            assert np.allclose(F[s_], F_h[rows].sum(0), rtol=1e-12, atol=1e-9)
        # (posteriors of the tensor-core path are fp16 x power of two: the total occupancy is exact to ~4e-6)
        assert abs(N.sum() - sel.sum()) < 2e-5 * sel.sum()
    finally:
        capi.set_gmm_kernel(0)


@pytest.mark.parametrize("shape,kern", [((256, 20, 5000), 1), ((256, 20, 5000), 0), ((2048, 60, 3000), 0)],
                         ids=["256c-simt", "256c-auto", "2048c-auto"])
def test_jfa_normalize_features(capi, oracle, shape, kern):
    """JFAAcc::normalizeFeatures (AccumulateJFAStat.cpp:4623-4680): x_t -= sum_k P(k | x_t) (U x)_k under the session
    model M + U x, for the selected frames only, in segment order.  One frame range is covered by two segments
    (compensated twice, the second time from the already compensated value), one segment is empty, and the frames
    outside every segment must come back bit-identical."""
    C, D, T = shape
    w, mean, cov = synth.make_ubm(C, D, seed=81)
    cov = cov * 3.0                      # soft posteriors: the offset is a real mixture over components
    X = synth.make_frames(w, mean, cov, T, seed=82)
    rng = np.random.default_rng(83)
    ux = 0.3 * np.sqrt(cov) * rng.standard_normal((C, D))
    capi.set_gmm_kernel(kern)
    try:
        g, o = capi.GMM(w, mean + ux, cov), oracle.gmm(w, mean + ux, cov)
        segs = [(100, 900), (1500, 0), (800, 600), (2000, T - 2300)]      # [800, 1000) twice
        Y = g.jfa_normalize_features(ux, X, [(b, n, 0) for b, n in segs])
        Y_ref = oracle.jfa_normalize_features(o, ux, X, segs)
    finally:
        capi.set_gmm_kernel(0)
    touched = np.zeros(T, bool)
    for b, n in segs:
        touched[b:b + n] = True
    assert np.array_equal(Y[~touched], X[~touched])
    moved = np.abs(Y_ref - X)[touched].max()
    assert moved > 0.05 * np.sqrt(cov).mean()                   # the compensation is not a no-op
    assert np.abs(Y - Y_ref).max() <= 1e-4 * moved + 4e-7 * np.abs(X).max(), np.abs(Y - Y_ref).max() / moved


def test_traintarget_validate_gmm_on_gpu(capi, golden_dir):
    """The reference's TrainTarget fixture (indicative pin, see tests/test_oracle_golden.py) through the CUDA
    EM-statistics path: occupancies -> MAPOccDep means vs the reference's adapted model."""
    from oracle import np_oracle
    z = np.load(os.path.join(golden_dir, "traintarget.npz"))
    cov = 1.0 / z["covinv"]
    g = capi.GMM(z["w"], z["mean"], cov)
    X = np.ascontiguousarray(z["frames"][z["selected"]], dtype=np.float32)
    _, n, occ, m1, m2 = g.em_accumulate(X)
    w_ml = occ / occ.sum()
    m_ml = np.where(occ[:, None] > 0, m1 / np.maximum(occ, 1e-300)[:, None], z["mean"])
    _, mean, _ = np_oracle.map_occ_dep(z["w"], z["mean"], cov, w_ml, m_ml, cov, n,
                                       r_mean=float(z["map_reg_factor_mean"]))
    ok = z["ok"]
    assert n == 32 and np.abs(mean - z["mean_ref"])[ok].max() < 1.2e-3


_FULL_REF = {}


@pytest.mark.parametrize("kname", ["tc", "tc2p", "tc_p1", "tc_p2"])
def test_full_size_oracle_parity(capi, oracle, kname):
    """BASELINE's own shape (2048c / 60d) against the fp64 oracle on 200 k frames in 10 utterances of 20 000
    frames (the oracle runs threaded: seconds).  The model is deliberately BLURRED (means shrunk towards the
    global mean, variances x 5): with the generator's own model every posterior is one-hot in 60 dimensions
    and fp16 posteriors are exact -- here a frame spreads over ~100 components (perplexity ~ 40), the case
    that exercises the fp16 posterior rounding.  Beyond the max-norm bound of the small cases:
      * per (utterance, component) RELATIVE checks wherever the occupation is >= 1: 1e-4 for occ >= 100 and
        the 3-sigma bound of the fp16 posterior rounding, 1.2e-3 / sqrt(occ), below that;
      * the contract itself: i-vectors (rank 40) computed from the GPU statistics vs from the oracle's agree
        to 1e-4 of the largest coefficient.
    "tc" runs the library default (lr_set_gmm_products level 0: five fp16 products per tile); the opt-in
    "tc_p1" (four products) meets the SAME bounds at these occupations -- the small cases of this file, where a
    component sees a handful of frames, are where it fails, which is why it is not the default; "tc_p2" (three
    products) is held to 3x the bounds: it sits at 1.1e-4 on the i-vectors, outside the contract
    (measured: scripts/products_probe.py)."""
    level = {"tc_p1": 1, "tc_p2": 2}.get(kname)
    slack = 3.0 if level == 2 else 1.0
    capi.set_gmm_kernel(KERNELS[kname.split("_")[0]])
    if level is not None:
        capi.set_gmm_products(level)
    try:
        C, D, U, per = 2048, 60, 10, 20000
        w, mean, cov = synth.make_ubm(C, D, seed=1)
        X = synth.make_frames(w, mean, cov, U * per, seed=21)
        w2, m2, c2 = w, mean * 0.15, cov * 5.0
        g, o = capi.GMM(w2, m2, c2), oracle.gmm(w2, m2, c2)
        f2r = (np.arange(U * per) // per).astype(np.int32)
        if "bw" not in _FULL_REF:      # the oracle pass (about 10 s) is shared by the two kernels
            _FULL_REF["bw"] = oracle.bwstats(o, X, f2r, U, threads=os.cpu_count() or 1)
        N_ref, F_ref = _FULL_REF["bw"]
        N, F = g.bwstats(X, [(u * per, per, u) for u in range(U)], U)
    finally:
        capi.set_gmm_kernel(0)
        capi.set_gmm_products(0)
    assert np.abs(N.sum(1) - per).max() < 1e-4 * per
    rel = np.abs(N - N_ref) / np.maximum(N_ref, 1e-300)
    big = N_ref >= 100.0
    mid = (N_ref >= 1.0) & ~big
    assert big.sum() > 50 and mid.sum() > 10000
    assert rel[big].max() < 1e-4 * slack, rel[big].max()
    assert (rel[mid] * np.sqrt(N_ref[mid])).max() < 1.2e-3 * slack, (rel[mid] * np.sqrt(N_ref[mid])).max()
    F3, F3r = F.reshape(U, C, D), F_ref.reshape(U, C, D)
    relF = np.linalg.norm(F3 - F3r, axis=2) / np.maximum(np.linalg.norm(F3r, axis=2), 1e-300)
    assert relF[big].max() < 1e-4 * slack, relF[big].max()
    assert (relF[mid] * np.sqrt(N_ref[mid])).max() < 1.2e-3 * slack
    # i-vectors from both sets of statistics
    R = 40
    invvar = (1.0 / c2).reshape(-1)
    Tm = synth.make_T(R, C, D, invvar, seed=23, scale=0.05)
    tett = oracle.tv_tett(Tm, invvar, C, D, threads=os.cpu_count() or 1)
    W_ref = oracle.tv_ivectors(N_ref, oracle.tv_subtract_m(N_ref, F_ref, m2.reshape(-1)), Tm, invvar, tett)
    W_gpu = oracle.tv_ivectors(N, oracle.tv_subtract_m(N, F, m2.reshape(-1)), Tm, invvar, tett)
    assert np.abs(W_gpu - W_ref).max() < 1e-4 * slack * np.abs(W_ref).max(), np.abs(W_gpu - W_ref).max() / np.abs(W_ref).max()
