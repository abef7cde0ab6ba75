"""The INT8 digit GEMM (csrc/gemm_i8.cu) behind the i-vector rows: fp64 operands cut into 7-bit digit
planes, exact int32 UMMA products, fp64 recombination.  Checked against numpy's fp64 product with a
bound relative to (row scale x column scale x K) -- the quantity the digit truncation is relative to --
and through the TV entry points against the cuBLAS fp64 cross-check and the oracle, including the
ill-conditioned posterior (a 60 000-frame utterance) VERDICT r1 asked for."""
import numpy as np
import pytest

from lia_ral_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from lia_ral_b200 import capi
    capi.init(0)
    yield capi
    capi.set_tv_gemm(0, 6)


def _bound(A, B, planes):
    """digit truncation: each operand keeps 7 planes bits below its row maximum; dropped classes add
    about as much again; products of K terms."""
    K = A.shape[1]
    return 32.0 * K * 2.0 ** (-7 * planes) * np.abs(A).max(axis=1)[:, None] * np.abs(B).max(axis=1)[None, :]


@pytest.mark.parametrize("M,N,K", [(128, 64, 128), (1, 1, 1), (37, 53, 100), (300, 200, 1000), (130, 700, 2048),
                                   (64, 40, 70000)],
                         ids=["one_tile", "scalar", "ragged", "multi_tile", "wide", "k_split"])
def test_gemm_digits_vs_numpy(capi, M, N, K):
    rng = np.random.default_rng(M * 1000 + N)
    A = rng.standard_normal((M, K)) * np.exp(rng.uniform(-6, 6, size=(M, 1)))   # rows of very different scale
    B = rng.standard_normal((N, K)) * np.exp(rng.uniform(-6, 6, size=(N, 1)))
    ref = A @ B.T
    for planes in (6, 7):
        C = capi.gemm_digits(A, B, planes=planes)
        err = np.abs(C - ref)
        assert (err <= _bound(A, B, planes) + 1e-15 * np.abs(ref)).all(), (planes, float((err / _bound(A, B, planes)).max()))
    # 6 planes are fp64-grade in the norm sense
    C6 = capi.gemm_digits(A, B, planes=6)
    assert np.abs(C6 - ref).max() <= 1e-10 * np.sqrt(K) * np.abs(A).max() * np.abs(B).max()


def test_gemm_digits_alpha_beta_and_zero_rows(capi):
    rng = np.random.default_rng(5)
    M, N, K = 200, 150, 640
    A = rng.standard_normal((M, K))
    B = rng.standard_normal((N, K))
    A[7] = 0.0           # an all-zero row has scale 0
    B[:, 600:] = 0.0
    C0 = rng.standard_normal((M, N))
    ref = C0 + 0.5 * (A @ B.T)
    C = capi.gemm_digits(A, B, C=C0, alpha=0.5, beta=1.0)
    assert np.abs(C - ref).max() <= 1e-11 * K
    assert np.array_equal(C[7], C0[7])
    # exactly representable operands give the exact integer product
    Ai = rng.integers(-64, 65, size=(M, K)).astype(np.float64)
    Bi = rng.integers(-64, 65, size=(N, K)).astype(np.float64)
    assert np.array_equal(capi.gemm_digits(Ai, Bi, planes=3), Ai @ Bi.T)


def test_gemm_digits_fewer_planes_lose_precision(capi):
    """The planes argument is live: 3 planes (21 bits) are visibly worse than 6."""
    rng = np.random.default_rng(9)
    A = rng.standard_normal((128, 512))
    B = rng.standard_normal((64, 512))
    ref = A @ B.T
    e3 = np.abs(capi.gemm_digits(A, B, planes=3) - ref).max()
    e6 = np.abs(capi.gemm_digits(A, B, planes=6) - ref).max()
    assert e6 < 1e-9 and e3 > 100 * e6 and e3 < 1e-3


def _tv_case(C, D, R, U, frames, seed, t_scale=0.05):
    w, mean, cov = synth.make_ubm(C, D, seed=seed)
    invvar = (1.0 / cov).reshape(-1)
    N, F = synth.make_bw_stats(U, w, mean, cov, frames_per_utt=frames, active=min(24, C), seed=seed + 1)
    T = synth.make_T(R, C, D, invvar, seed=seed + 2, scale=t_scale)
    return dict(C=C, D=D, R=R, U=U, mean=mean.reshape(-1), invvar=invvar, N=N, F=F, T=T)


def _estep(capi, c, which, planes=0):
    capi.set_tv_gemm(which, planes)
    tv = capi.TV(c["C"], c["D"], c["R"], c["U"], c["mean"], c["invvar"])
    tv.set_stats(c["N"], c["F"])
    tv.set_T(c["T"])
    tv.reset_tmp_acc()
    tv.subtract_m()
    tv.estimate_tett()
    tv.estimate_a_and_c()
    W = tv.get_W()
    A, Cmx, Rm, r, mw = tv.get_acc()
    tv.update_t()
    return W, A, Cmx, Rm, tv.get_T()


def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def test_tv_estep_digit_gemm_matches_cublas_fp64(capi):
    """Same E-step + M-step through the INT8 digit GEMM and through cuBLAS fp64."""
    c = _tv_case(128, 20, 48, 300, 400, seed=31)
    got = _estep(capi, c, 0, 6)
    ref = _estep(capi, c, 1)
    for g, r, name in zip(got, ref, ("W", "A", "Cmx", "Rm", "T")):
        assert _rel(g, r) < 1e-9, name


def test_ill_conditioned_posterior(capi, oracle):
    """A 60 000-frame utterance against a spread T: cond(L) ~ 1e4.  i-vectors within 1e-6 of the oracle
    (contract: 1e-4)."""
    c = _tv_case(128, 20, 48, 6, 60000, seed=41, t_scale=0.1)
    c["T"] = c["T"] * np.geomspace(1.0, 1e-3, c["R"])[:, None]   # spread spectrum: L's eigenvalues from ~1 to ~1e4
    Fc = oracle.tv_subtract_m(c["N"], c["F"], c["mean"])
    tett = oracle.tv_tett(c["T"], c["invvar"], c["C"], c["D"])
    L0 = np.eye(c["R"]) + np.tensordot(c["N"][0], tett, axes=1)
    cond = np.linalg.cond(L0)
    assert cond > 3e3, cond
    W_ref = oracle.tv_ivectors(c["N"], Fc, c["T"], c["invvar"], tett)
    capi.set_tv_gemm(0, 6)
    tv = capi.TV(c["C"], c["D"], c["R"], c["U"], c["mean"], c["invvar"])
    tv.set_stats(c["N"], c["F"])
    tv.set_T(c["T"])
    tv.subtract_m()
    tv.estimate_tett()
    tv.estimate_w()
    assert _rel(tv.get_W(), W_ref) < 1e-6
