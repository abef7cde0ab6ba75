"""ctypes binding of liblia_ral_b200.so (the C ABI in include/lia_ral_b200.h).

numpy in / numpy out; every failure of the library is raised as `LrError` carrying
`lr_last_error()` (the reference throws alize::Exception at the same points).  There is no CPU
path behind these calls: without the built CUDA library, or without a B200, they raise.
"""
import ctypes as ct
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblia_ral_b200.so")

c_dp = ct.POINTER(ct.c_double)
c_fp = ct.POINTER(ct.c_float)
c_ip = ct.POINTER(ct.c_int32)
c_up = ct.POINTER(ct.c_uint32)


class LrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"lia_ral_b200 error {code}: {msg}")
        self.code = code


class LrSeg(ct.Structure):
    _fields_ = [("begin", ct.c_int64), ("length", ct.c_int64), ("row", ct.c_int32),
                ("pad_", ct.c_int32)]


_lib = None


def build(force=False):
    """Compile the CUDA library in-tree with nvcc for sm_100a (csrc/Makefile)."""
    import subprocess
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-s", "-j", str(os.cpu_count() or 4)]
    if force:
        cmd.append("-B")
    subprocess.check_call(cmd)
    return LIB_PATH


def lib():
    """Load the library (once).  Missing library is a hard error, never a fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LrError(-1, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; "
                          "g.build()'` (nvcc, sm_100a); this engine has no CPU fallback")
    L = ct.CDLL(LIB_PATH)
    L.lr_last_error.restype = ct.c_char_p
    L.lr_version.restype = ct.c_char_p
    L.lr_stream_handle.restype = ct.c_uint64
    L.lr_launch_count.restype = ct.c_uint64
    L.lr_reset_launch_count.restype = None
    for name in ("lr_gmm_create", "lr_feats_upload", "lr_feats_wrap_device", "lr_tv_create",
                 "lr_tv_dev_N", "lr_tv_dev_F", "lr_tv_dev_acc"):
        getattr(L, name).restype = ct.c_void_p
    for name in ("lr_gmm_em_stats_len", "lr_tv_acc_len", "lr_tv_acc_a_stride"):
        getattr(L, name).restype = ct.c_size_t
    for name in ("lr_gmm_destroy", "lr_feats_destroy", "lr_tv_destroy"):
        getattr(L, name).restype = None
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise LrError(rc, lib().lr_last_error().decode("utf-8", "replace"))


def _handle(p):
    if not p:
        raise LrError(-1, lib().lr_last_error().decode("utf-8", "replace"))
    return ct.c_void_p(p)


def _d(a):
    return a.ctypes.data_as(c_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _f32(a):
    a = np.asarray(a)
    assert a.dtype == np.float32 and a.ndim == 2 and a.strides[1] == 4, "frames must be float32 [T, D]"
    return a


def _segs(segs):
    """[(begin, length, row), ...] -> (LrSeg array or None, count)"""
    if segs is None:
        return None, 0
    arr = (LrSeg * len(segs))()
    for i, s in enumerate(segs):
        arr[i].begin, arr[i].length = int(s[0]), int(s[1])
        arr[i].row = int(s[2]) if len(s) > 2 else 0
    return arr, len(segs)


def init(device=0):
    _check(lib().lr_init(int(device)))


def shutdown():
    _check(lib().lr_shutdown())


def synchronize():
    _check(lib().lr_synchronize())


def stream_handle():
    return int(lib().lr_stream_handle())


def launch_count():
    return int(lib().lr_launch_count())


def reset_launch_count():
    lib().lr_reset_launch_count()


def profile(enable):
    _check(lib().lr_profile(int(bool(enable))))


def profile_read(kind):
    """-> (total device ms, launches) of kernel kind 0 (LLK pass) / 1 (statistics pass)."""
    ms, n = ct.c_double(0.0), ct.c_uint64(0)
    _check(lib().lr_profile_read(int(kind), ct.byref(ms), ct.byref(n)))
    return ms.value, int(n.value)


def set_gmm_kernel(which):
    """0 = auto, 1 = fp32 SIMT, 2 = tcgen05."""
    _check(lib().lr_set_gmm_kernel(int(which)))


def get_gmm_kernel():
    return int(lib().lr_get_gmm_kernel())


def set_gmm_products(level):
    """0 = five fp16 products per tile (default), 1 = four, 2 = three (see include/lia_ral_b200.h)."""
    _check(lib().lr_set_gmm_products(int(level)))


def get_gmm_products():
    return int(lib().lr_get_gmm_products())


def set_tv_gemm(which=0, planes=0):
    """Contraction kernel of the TV rows: 0 = INT8 digit GEMM (default), 1 = cuBLAS fp64 cross-check;
    planes = digit planes per operand (3..7, 0 = keep).  Effective at the next estimate_tett()."""
    _check(lib().lr_set_tv_gemm(int(which), int(planes)))


def gemm_digits(A, B, C=None, alpha=1.0, beta=0.0, planes=0):
    """C = beta C + alpha A B^T through the INT8 digit GEMM (A [M x K], B [N x K], fp64)."""
    A, B = _f64(A), _f64(B)
    M, K = A.shape
    N = B.shape[0]
    assert B.shape[1] == K
    out = np.zeros((M, N)) if C is None else _f64(C).copy()
    _check(lib().lr_gemm_digits(ct.c_size_t(M), ct.c_size_t(N), ct.c_size_t(K), _d(A), _d(B), _d(out),
                                ct.c_double(alpha), ct.c_double(beta), int(planes)))
    return out


class Feats:
    """Frames resident in HBM (FeatureServer buffer twin)."""

    def __init__(self, X=None, device_ptr=None, T=None, ldx=None, D=None):
        L = lib()
        if X is not None:
            X = _f32(X)
            self.T, self.D = X.shape
            self.ldx = X.strides[0] // 4
            self.h = _handle(L.lr_feats_upload(X.ctypes.data_as(c_fp), ct.c_size_t(self.T),
                                               ct.c_size_t(self.ldx), self.D))
        else:
            self.T, self.ldx, self.D = int(T), int(ldx), int(D)
            self.h = _handle(L.lr_feats_wrap_device(ct.c_void_p(int(device_ptr)),
                                                    ct.c_size_t(self.T), ct.c_size_t(self.ldx),
                                                    self.D))

    def invalidate(self):
        """The wrapped frames changed: drop the cached tensor-core operand."""
        _check(lib().lr_feats_invalidate(self.h))

    def close(self):
        if self.h:
            lib().lr_feats_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GMM:
    """Device-resident MixtureGD twin."""

    def __init__(self, w, mean, cov):
        mean = _f64(mean)
        self.C, self.D = mean.shape
        w, cov = _f64(w), _f64(cov)
        assert w.shape == (self.C,) and cov.shape == mean.shape
        self.h = _handle(lib().lr_gmm_create(self.C, self.D, _d(w), _d(mean), _d(cov)))

    def close(self):
        if self.h:
            lib().lr_gmm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set(self, w, mean, cov):
        _check(lib().lr_gmm_set(self.h, _d(_f64(w)), _d(_f64(mean)), _d(_f64(cov))))

    def set_cst(self, cst):
        _check(lib().lr_gmm_set_cst(self.h, _d(_f64(cst))))

    def get(self):
        C, D = self.C, self.D
        out = dict(w=np.empty(C), mean=np.empty((C, D)), cov=np.empty((C, D)),
                   covinv=np.empty((C, D)), cst=np.empty(C), det=np.empty(C))
        _check(lib().lr_gmm_get(self.h, _d(out["w"]), _d(out["mean"]), _d(out["cov"]),
                                _d(out["covinv"]), _d(out["cst"]), _d(out["det"])))
        return out

    # ---- a4/a5
    def em_accumulate(self, X, segs=None, weight=1.0, occ=None, m1=None, m2=None):
        X = _f32(X)
        C, D = self.C, self.D
        occ = np.zeros(C) if occ is None else occ
        m1 = np.zeros((C, D)) if m1 is None else m1
        m2 = np.zeros((C, D)) if m2 is None else m2
        llk, nfr = ct.c_double(0.0), ct.c_double(0.0)
        sa, ns = _segs(segs)
        _check(lib().lr_gmm_em_accumulate(self.h, X.ctypes.data_as(c_fp), ct.c_size_t(X.shape[0]),
                                          ct.c_size_t(X.strides[0] // 4), sa, ct.c_size_t(ns),
                                          ct.c_double(weight), _d(occ), _d(m1), _d(m2),
                                          ct.byref(llk), ct.byref(nfr)))
        return llk.value, nfr.value, occ, m1, m2

    def em_stats_len(self):
        return int(lib().lr_gmm_em_stats_len(self.h))

    def em_accumulate_dev(self, feats, t0, T, weight, d_stats_ptr):
        _check(lib().lr_gmm_em_accumulate_dev(self.h, feats.h, ct.c_size_t(t0), ct.c_size_t(T),
                                              ct.c_double(weight), ct.c_void_p(int(d_stats_ptr))))

    def em_update_dev(self, d_stats_ptr, flooring=0.0, ceiling=0.0, d_cov_signal_ptr=None):
        sig = ct.c_void_p(int(d_cov_signal_ptr)) if d_cov_signal_ptr else None
        _check(lib().lr_gmm_em_update_dev(self.h, ct.c_void_p(int(d_stats_ptr)),
                                          ct.c_double(flooring), ct.c_double(ceiling), sig))

    def em_update(self, occ, m1, m2, flooring=0.0, ceiling=0.0, cov_signal=None):
        sig = _d(_f64(cov_signal)) if cov_signal is not None else None
        _check(lib().lr_gmm_em_update(self.h, _d(_f64(occ)), _d(_f64(m1)), _d(_f64(m2)),
                                      ct.c_double(flooring), ct.c_double(ceiling), sig))

    # ---- a2/a3
    def bwstats(self, X, segs, U, N=None, F=None):
        X = _f32(X)
        N = np.zeros((U, self.C)) if N is None else N
        F = np.zeros((U, self.C * self.D)) if F is None else F
        sa, ns = _segs(segs)
        _check(lib().lr_gmm_bwstats(self.h, X.ctypes.data_as(c_fp), ct.c_size_t(X.shape[0]),
                                    ct.c_size_t(X.strides[0] // 4), sa, ct.c_size_t(ns),
                                    ct.c_size_t(U), _d(N), _d(F)))
        return N, F

    def jfa_bwstats(self, X, segs, speaker_of_session, n_speakers):
        """(N_h, F_h, N, F): per-session and per-speaker Baum-Welch statistics of JFAAcc; segs rows = sessions."""
        X = _f32(X)
        spk = np.ascontiguousarray(speaker_of_session, dtype=np.int32)
        nh = len(spk)
        N_h, F_h = np.zeros((nh, self.C)), np.zeros((nh, self.C * self.D))
        N, F = np.zeros((n_speakers, self.C)), np.zeros((n_speakers, self.C * self.D))
        sa, ns = _segs(segs)
        _check(lib().lr_jfa_bwstats(self.h, X.ctypes.data_as(c_fp), ct.c_size_t(X.shape[0]),
                                    ct.c_size_t(X.strides[0] // 4), sa, ct.c_size_t(ns), ct.c_size_t(nh),
                                    spk.ctypes.data_as(c_ip), ct.c_size_t(n_speakers), _d(N_h), _d(F_h), _d(N), _d(F)))
        return N_h, F_h, N, F

    def jfa_normalize_features(self, ux, X, segs):
        """JFAAcc::normalizeFeatures with self as the session model: returns the compensated copy of X."""
        X = np.array(_f32(X), copy=True, order="C")
        ux = np.ascontiguousarray(ux, dtype=np.float64).reshape(-1)
        assert ux.size == self.C * self.D
        sa, ns = _segs(segs)
        _check(lib().lr_jfa_normalize_features(self.h, _d(ux), X.ctypes.data_as(c_fp), ct.c_size_t(X.shape[0]),
                                               ct.c_size_t(X.strides[0] // 4), sa, ct.c_size_t(ns)))
        return X

    def bwstats_dev(self, feats, segs, U, d_N_ptr, d_F_ptr):
        sa, ns = _segs(segs)
        _check(lib().lr_gmm_bwstats_dev(self.h, feats.h, sa, ct.c_size_t(ns), ct.c_size_t(U),
                                        ct.c_void_p(int(d_N_ptr)), ct.c_void_p(int(d_F_ptr))))

    # ---- a7
    def llk_topk(self, X, K, complete=True, min_llk=-200.0, max_llk=200.0):
        X = _f32(X)
        T = X.shape[0]
        llk = np.empty(T)
        idx = np.empty((T, K), dtype=np.uint32)
        top_lk = np.empty((T, K))
        rest_lk, rest_w = np.empty(T), np.empty(T)
        _check(lib().lr_gmm_llk_topk(self.h, X.ctypes.data_as(c_fp), ct.c_size_t(T),
                                     ct.c_size_t(X.strides[0] // 4), K, int(complete),
                                     ct.c_double(min_llk), ct.c_double(max_llk), _d(llk),
                                     idx.ctypes.data_as(c_up), _d(top_lk), _d(rest_lk),
                                     _d(rest_w)))
        return llk, idx, top_lk, rest_lk, rest_w

    def llk_use_topk(self, X, idx, rest_lk, complete=True, min_llk=-200.0, max_llk=200.0):
        X = _f32(X)
        T = X.shape[0]
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        K = idx.shape[1]
        llk = np.empty(T)
        rl = _d(_f64(rest_lk)) if rest_lk is not None else None
        _check(lib().lr_gmm_llk_use_topk(self.h, X.ctypes.data_as(c_fp), ct.c_size_t(T),
                                         ct.c_size_t(X.strides[0] // 4), K,
                                         idx.ctypes.data_as(c_up), rl, int(complete),
                                         ct.c_double(min_llk), ct.c_double(max_llk), _d(llk)))
        return llk

    def llk(self, X, min_llk=-200.0, max_llk=200.0):
        X = _f32(X)
        T = X.shape[0]
        out = np.empty(T)
        _check(lib().lr_gmm_llk(self.h, X.ctypes.data_as(c_fp), ct.c_size_t(T),
                                ct.c_size_t(X.strides[0] // 4), ct.c_double(min_llk),
                                ct.c_double(max_llk), _d(out)))
        return out


def frames_mean_cov(X):
    X = _f32(X)
    D = X.shape[1]
    mean, cov = np.empty(D), np.empty(D)
    _check(lib().lr_frames_mean_cov(X.ctypes.data_as(c_fp), ct.c_size_t(X.shape[0]),
                                    ct.c_size_t(X.strides[0] // 4), D, _d(mean), _d(cov)))
    return mean, cov


def compute_test(world, clients, X, segs=None, K=10, complete=True, min_llk=-200.0, max_llk=200.0,
                 per_segment=False, world_decime=1):
    """Frame loop of ComputeTest() for one test file -> (mean_llk_world[n_out],
    mean_llk_client[n_clients, n_out])."""
    X = _f32(X)
    sa, ns = _segs(segs)
    n_out = (ns if segs is not None else 1) if per_segment else 1
    mw = np.empty(n_out)
    mc = np.empty((len(clients), n_out))
    arr = (ct.c_void_p * max(1, len(clients)))(*[c.h.value for c in clients])
    _check(lib().lr_compute_test_decime(world.h, arr, len(clients), X.ctypes.data_as(c_fp),
                                        ct.c_size_t(X.shape[0]), ct.c_size_t(X.strides[0] // 4), sa,
                                        ct.c_size_t(ns), K, int(complete), ct.c_double(min_llk),
                                        ct.c_double(max_llk), int(per_segment), int(world_decime), _d(mw), _d(mc)))
    return mw, mc


class TV:
    """Device twin of the reference's TVAcc object (AccumulateTVStat.h)."""

    def __init__(self, C, D, R, U, ubm_mean, ubm_invvar):
        self.C, self.D, self.R, self.U = int(C), int(D), int(R), int(U)
        m, iv = _f64(ubm_mean).reshape(-1), _f64(ubm_invvar).reshape(-1)
        assert m.size == C * D and iv.size == C * D
        self.h = _handle(lib().lr_tv_create(self.C, self.D, self.R, ct.c_size_t(self.U), _d(m),
                                            _d(iv)))

    def close(self):
        if self.h:
            lib().lr_tv_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stats(self, N, F):
        N, F = _f64(N), _f64(F)
        assert N.shape == (self.U, self.C) and F.size == self.U * self.C * self.D
        _check(lib().lr_tv_set_stats(self.h, _d(N), _d(F)))

    def get_stats(self):
        N, F = np.empty((self.U, self.C)), np.empty((self.U, self.C * self.D))
        _check(lib().lr_tv_get_stats(self.h, _d(N), _d(F)))
        return N, F

    def dev_N(self):
        return int(lib().lr_tv_dev_N(self.h))

    def dev_F(self):
        return int(lib().lr_tv_dev_F(self.h))

    def set_T(self, T):
        T = _f64(T)
        assert T.shape == (self.R, self.C * self.D)
        _check(lib().lr_tv_set_T(self.h, _d(T)))

    def get_T(self):
        T = np.empty((self.R, self.C * self.D))
        _check(lib().lr_tv_get_T(self.h, _d(T)))
        return T

    def get_mean(self):
        m = np.empty(self.C * self.D)
        _check(lib().lr_tv_get_mean(self.h, _d(m)))
        return m

    def set_mean(self, mean):
        m = _f64(mean).reshape(-1)
        assert m.size == self.C * self.D
        _check(lib().lr_tv_set_mean(self.h, _d(m)))

    def get_W(self):
        W = np.empty((self.U, self.R))
        _check(lib().lr_tv_get_W(self.h, _d(W)))
        return W

    def get_acc(self, want_A=True):
        R, C, D = self.R, self.C, self.D
        A = np.empty((C, R * R)) if want_A else None
        Cmx, Rm, r, mw = np.empty((R, C * D)), np.empty((R, R)), np.empty(R), np.empty(R)
        _check(lib().lr_tv_get_acc(self.h, _d(A) if want_A else None, _d(Cmx), _d(Rm), _d(r),
                                   _d(mw)))
        return A, Cmx, Rm, r, mw

    def reset_tmp_acc(self):
        _check(lib().lr_tv_reset_tmp_acc(self.h))

    def subtract_m(self):
        _check(lib().lr_tv_subtract_m(self.h))

    def estimate_tett(self):
        _check(lib().lr_tv_estimate_tett(self.h))

    def estimate_w(self):
        _check(lib().lr_tv_estimate_w(self.h))

    def estimate_a_and_c(self):
        _check(lib().lr_tv_estimate_a_and_c(self.h))

    def update_t(self):
        _check(lib().lr_tv_update_t(self.h))

    def min_divergence(self, n_sessions):
        _check(lib().lr_tv_min_divergence(self.h, ct.c_double(n_sessions)))

    def orthonormalize_t(self):
        _check(lib().lr_tv_orthonormalize_t(self.h))

    # ---- approximate i-vector modes (IvExtractor.cpp:151-363)
    def norm_t(self):
        _check(lib().lr_tv_norm_t(self.h))

    def norm_statistics(self):
        _check(lib().lr_tv_norm_statistics(self.h))

    def weighted_cov(self, weight):
        w = _f64(weight).reshape(-1)
        assert w.size == self.C
        W = np.empty((self.R, self.R))
        _check(lib().lr_tv_weighted_cov(self.h, _d(w), _d(W)))
        return W

    def approximate_tctc(self, Q):
        Q = _f64(Q)
        assert Q.shape == (self.R, self.R)
        Dm = np.empty((self.C, self.R))
        _check(lib().lr_tv_approximate_tctc(self.h, _d(Q), _d(Dm)))
        return Dm

    def estimate_w_ubm_weight(self, Wcov):
        Wcov = _f64(Wcov)
        assert Wcov.shape == (self.R, self.R)
        _check(lib().lr_tv_estimate_w_ubm_weight(self.h, _d(Wcov)))

    def estimate_w_eigen_decomposition(self, Dm, Q):
        Dm, Q = _f64(Dm), _f64(Q)
        assert Dm.shape == (self.C, self.R) and Q.shape == (self.R, self.R)
        _check(lib().lr_tv_estimate_w_eigen_decomposition(self.h, _d(Dm), _d(Q)))

    # ---- component-sharded M-step (multi-GPU)
    def update_t_range(self, c0, c1):
        _check(lib().lr_tv_update_t_range(self.h, int(c0), int(c1)))

    def pack_t(self, c0, c1, dev_ptr):
        _check(lib().lr_tv_pack_t(self.h, int(c0), int(c1), ct.c_void_p(int(dev_ptr))))

    def unpack_t(self, c0, c1, dev_ptr):
        _check(lib().lr_tv_unpack_t(self.h, int(c0), int(c1), ct.c_void_p(int(dev_ptr))))

    def acc_a_stride(self):
        return int(lib().lr_tv_acc_a_stride(self.h))

    def dev_acc(self):
        return int(lib().lr_tv_dev_acc(self.h))

    def acc_len(self):
        return int(lib().lr_tv_acc_len(self.h))

    def finish_estep(self, n_speakers_total):
        _check(lib().lr_tv_finish_estep(self.h, ct.c_double(n_speakers_total)))


def eigen_problem(EP, rank=None):
    """computeEigenProblem (AccumulateTVStat.cpp:2999): eigvec[n x rank] (column j = j-th largest), eigval."""
    EP = _f64(EP)
    n = EP.shape[0]
    rank = n if rank is None else int(rank)
    vec, val = np.empty((n, rank)), np.empty(rank)
    _check(lib().lr_eigen_problem(n, _d(EP), rank, _d(vec), _d(val)))
    return vec, val


def plda_native_scoring(F, G, Sigma, models, model_of, segments):
    F, Sigma = _f64(F), _f64(Sigma)
    d, rF = F.shape
    rG = 0 if G is None else G.shape[1]
    Gp = _d(_f64(G)) if rG else None
    models, segments = _f64(models), _f64(segments)
    model_of = np.ascontiguousarray(model_of, dtype=np.int32)
    n_models = len(np.unique(model_of))
    scores = np.empty((n_models, segments.shape[1]))
    _check(lib().lr_plda_native_scoring(d, rF, rG, _d(F), Gp, _d(Sigma), _d(models),
                                        ct.c_size_t(models.shape[1]),
                                        model_of.ctypes.data_as(c_ip), ct.c_size_t(n_models),
                                        _d(segments), ct.c_size_t(segments.shape[1]), _d(scores)))
    return scores


def plda_native_scoring_dev(F, G, Sigma, d_models_ptr, n_enrol, model_of, d_segments_ptr, n_test, d_scores_ptr,
                            ld_scores=None):
    """device-resident variant: models [d x n_enrol] / segments [d x n_test] fp64 and the fp32 score block
    [n_models x ld_scores] live in device memory (pointers as integers)."""
    F, Sigma = _f64(F), _f64(Sigma)
    d, rF = F.shape
    rG = 0 if G is None else G.shape[1]
    Gp = _d(_f64(G)) if rG else None
    model_of = np.ascontiguousarray(model_of, dtype=np.int32)
    n_models = len(np.unique(model_of))
    _check(lib().lr_plda_native_scoring_dev(d, rF, rG, _d(F), Gp, _d(Sigma), ct.c_void_p(d_models_ptr),
                                            ct.c_size_t(n_enrol), model_of.ctypes.data_as(c_ip),
                                            ct.c_size_t(n_models), ct.c_void_p(d_segments_ptr),
                                            ct.c_size_t(n_test), ct.c_void_p(d_scores_ptr),
                                            ct.c_size_t(ld_scores or n_test)))
    return n_models


# ---- i-vector back-end (PldaDev statistics / normalisation, non-PLDA scorings) ----------------
def _cls(class_of):
    return np.ascontiguousarray(class_of, dtype=np.int32)


def iv_cov_mat(data, class_of, n_spk):
    """PldaDev::computeAll + computeCovMat -> (mean, speaker_means, Sigma, W, B)."""
    data, cls = _f64(data), _cls(class_of)
    d, n = data.shape
    mean, sm = np.empty(d), np.empty((d, n_spk))
    S, W, B = np.empty((d, d)), np.empty((d, d)), np.empty((d, d))
    _check(lib().lr_iv_cov_mat(d, ct.c_size_t(n), _d(data), cls.ctypes.data_as(c_ip), ct.c_size_t(n_spk),
                               _d(mean), _d(sm), _d(S), _d(W), _d(B)))
    return mean, sm, S, W, B


def iv_wccn_chol(data, class_of, n_spk):
    data, cls = _f64(data), _cls(class_of)
    d, n = data.shape
    out = np.empty((d, d))
    _check(lib().lr_iv_wccn_chol(d, ct.c_size_t(n), _d(data), cls.ctypes.data_as(c_ip),
                                 ct.c_size_t(n_spk), _d(out)))
    return out


def iv_mahalanobis_matrix(data, class_of, n_spk):
    data, cls = _f64(data), _cls(class_of)
    d, n = data.shape
    out = np.empty((d, d))
    _check(lib().lr_iv_mahalanobis_matrix(d, ct.c_size_t(n), _d(data), cls.ctypes.data_as(c_ip),
                                          ct.c_size_t(n_spk), _d(out)))
    return out


def iv_efr_matrix(cov):
    cov = _f64(cov)
    out = np.empty_like(cov)
    _check(lib().lr_iv_efr_matrix(cov.shape[0], _d(cov), _d(out)))
    return out


def iv_lda(W, B, rank):
    W, B = _f64(W), _f64(B)
    out = np.empty((rank, W.shape[0]))
    _check(lib().lr_iv_lda(W.shape[0], _d(W), _d(B), int(rank), _d(out)))
    return out


def iv_normalize(data, mu=None, M=None, length_norm=False):
    """center -> rotateLeft -> lengthNorm (each optional), vectors in columns."""
    data = _f64(data)
    d, n = data.shape
    r = 0
    if M is not None:
        M = _f64(M)
        r = M.shape[0]
        assert M.shape[1] == d
    out = np.empty((r if M is not None else d, n))
    _check(lib().lr_iv_normalize(d, ct.c_size_t(n), _d(data), _d(_f64(mu)) if mu is not None else None,
                                 _d(M) if M is not None else None, r, int(bool(length_norm)), _d(out)))
    return out


def _trials(trials, nm, nt):
    if trials is None:
        return None, None
    t = np.ascontiguousarray(trials, dtype=np.uint8)
    assert t.shape == (nm, nt)
    return t, t.ctypes.data_as(ct.POINTER(ct.c_uint8))


def iv_cosine_scoring(models, segments, trials=None):
    models, segments = _f64(models), _f64(segments)
    d, nm = models.shape
    nt = segments.shape[1]
    keep, tp = _trials(trials, nm, nt)
    sc = np.empty((nm, nt))
    _check(lib().lr_iv_cosine_scoring(d, ct.c_size_t(nm), ct.c_size_t(nt), _d(models), _d(segments), tp,
                                      _d(sc)))
    return sc


def iv_mahalanobis_scoring(models, segments, Mah, trials=None):
    models, segments, Mah = _f64(models), _f64(segments), _f64(Mah)
    d, nm = models.shape
    nt = segments.shape[1]
    keep, tp = _trials(trials, nm, nt)
    sc = np.empty((nm, nt))
    _check(lib().lr_iv_mahalanobis_scoring(d, ct.c_size_t(nm), ct.c_size_t(nt), _d(models), _d(segments),
                                           _d(Mah), tp, _d(sc)))
    return sc


def iv_two_cov_scoring(models, segments, W, B):
    models, segments = _f64(models), _f64(segments)
    d, nm = models.shape
    nt = segments.shape[1]
    sc = np.empty((nm, nt))
    _check(lib().lr_iv_two_cov_scoring(d, ct.c_size_t(nm), ct.c_size_t(nt), _d(models), _d(segments),
                                       _d(_f64(W)), _d(_f64(B)), _d(sc)))
    return sc


def plda_em_iteration(data, class_of, n_spk, F, G, Sigma, Delta):
    """One PldaModel::em_iteration -> (data centred by Delta, F, G, Sigma, Delta)."""
    data, F, Sigma, Delta = _f64(data).copy(), _f64(F).copy(), _f64(Sigma).copy(), _f64(Delta).copy()
    d, n = data.shape
    rF = F.shape[1]
    rG = 0 if G is None else G.shape[1]
    G = np.zeros((d, 0)) if rG == 0 else _f64(G).copy()
    cls = _cls(class_of)
    _check(lib().lr_plda_em_iteration(d, rF, rG, ct.c_size_t(n), _d(data), cls.ctypes.data_as(c_ip),
                                      ct.c_size_t(n_spk), _d(F), _d(G) if rG else None, _d(Sigma), _d(Delta)))
    return data, F, G, Sigma, Delta
