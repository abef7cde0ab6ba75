"""ctypes binding of oracle/lr_oracle.c (numpy in / numpy out).

TEST INFRASTRUCTURE ONLY (see lr_oracle.h).  `Oracle()` loads the IEEE parity build,
`Oracle(fast=True)` the -O3 -ffast-math + pthreads build that plays the reference's
`--enable-MT` CPU path in bench.py.
"""
import ctypes as ct
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")

c_dp = ct.POINTER(ct.c_double)
c_fp = ct.POINTER(ct.c_float)
c_ip = ct.POINTER(ct.c_int32)
c_up = ct.POINTER(ct.c_uint32)


def build(force=False):
    """Compile the oracle with gcc (oracle/Makefile). Building the checker is not using it."""
    libs = [os.path.join(_BUILD, n) for n in ("liblr_oracle.so", "liblr_oracle_fast.so")]
    src = [os.path.join(_HERE, n) for n in ("lr_oracle.c", "lr_oracle.h")]
    stale = force or any(
        not os.path.exists(l) or os.path.getmtime(l) < max(os.path.getmtime(s) for s in src)
        for l in libs
    )
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return libs


def _d(a):
    return a.ctypes.data_as(c_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _f32(a):
    a = np.asarray(a)
    assert a.dtype == np.float32 and a.ndim == 2 and a.strides[1] == 4
    return a


class GMM:
    """Plain container: weights, means, cov (+ computeAll products)."""

    def __init__(self, w, mean, cov, oracle):
        self.w = _f64(w)
        self.mean = _f64(mean)
        self.cov = _f64(cov)
        self.C, self.D = self.mean.shape
        self.covinv = np.empty_like(self.cov)
        self.det = np.empty(self.C)
        self.cst = np.empty(self.C)
        oracle.lib.orc_gmm_compute_all(self.C, self.D, _d(self.cov), _d(self.covinv),
                                       _d(self.det), _d(self.cst))

    def args(self):
        return (self.C, self.D, _d(self.w), _d(self.mean), _d(self.covinv), _d(self.cst))


class Oracle:
    def __init__(self, fast=False):
        build()
        name = "liblr_oracle_fast.so" if fast else "liblr_oracle.so"
        self.lib = ct.CDLL(os.path.join(_BUILD, name))
        L = self.lib
        L.orc_frame_likelihoods.restype = ct.c_double
        L.orc_em_accumulate.restype = ct.c_double
        L.orc_set_it_parameter.restype = ct.c_double
        L.orc_set_it_parameter.argtypes = [ct.c_double, ct.c_double, ct.c_int, ct.c_int]
        L.orc_distrib_lk.restype = ct.c_double

    # ---- GMM -----------------------------------------------------------------
    def gmm(self, w, mean, cov):
        return GMM(w, mean, cov, self)

    def frame_likelihoods(self, g, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        p = np.empty(g.C)
        s = self.lib.orc_frame_likelihoods(*g.args(), x.ctypes.data_as(c_fp), _d(p))
        return s, p

    def bwstats(self, g, X, frame2row, U, N=None, F=None, threads=1):
        X = _f32(X)
        T = X.shape[0]
        N = np.zeros((U, g.C)) if N is None else N
        F = np.zeros((U, g.C * g.D)) if F is None else F
        f2r = None
        if frame2row is not None:
            f2r = np.ascontiguousarray(frame2row, dtype=np.int32)
        self.lib.orc_bwstats(*g.args(), X.ctypes.data_as(c_fp), ct.c_size_t(T),
                             ct.c_size_t(X.strides[0] // 4),
                             f2r.ctypes.data_as(c_ip) if f2r is not None else None,
                             ct.c_size_t(U), _d(N), _d(F), threads)
        return N, F

    def jfa_normalize_features(self, g, ux, X, segs):
        """JFAAcc::normalizeFeatures under the session model g; segs = [(begin, length), ...]; returns a copy."""
        X = np.array(_f32(X), copy=True, order="C")
        ux = np.ascontiguousarray(ux, dtype=np.float64).reshape(-1)
        b = np.ascontiguousarray([s[0] for s in segs], dtype=np.int64)
        n = np.ascontiguousarray([s[1] for s in segs], dtype=np.int64)
        self.lib.orc_jfa_normalize_features(*g.args(), _d(ux), X.ctypes.data_as(c_fp), ct.c_size_t(X.strides[0] // 4),
                                            b.ctypes.data_as(ct.POINTER(ct.c_int64)),
                                            n.ctypes.data_as(ct.POINTER(ct.c_int64)), ct.c_size_t(len(segs)))
        return X

    def em_accumulate(self, g, X, weight=1.0, occ=None, m1=None, m2=None, threads=1):
        X = _f32(X)
        T = X.shape[0]
        occ = np.zeros(g.C) if occ is None else occ
        m1 = np.zeros((g.C, g.D)) if m1 is None else m1
        m2 = np.zeros((g.C, g.D)) if m2 is None else m2
        nfr = ct.c_double(0.0)
        llk = self.lib.orc_em_accumulate(*g.args(), X.ctypes.data_as(c_fp), ct.c_size_t(T),
                                         ct.c_size_t(X.strides[0] // 4), ct.c_double(weight),
                                         _d(occ), _d(m1), _d(m2), ct.byref(nfr), threads)
        return llk, nfr.value, occ, m1, m2

    def em_get(self, g, occ, m1, m2):
        w, mean, cov = g.w.copy(), g.mean.copy(), g.cov.copy()
        self.lib.orc_em_get(g.C, g.D, _d(_f64(occ)), _d(_f64(m1)), _d(_f64(m2)), _d(w), _d(mean),
                            _d(cov))
        return w, mean, cov

    def variance_control(self, cov, flooring, ceiling, cov_signal):
        cov = _f64(cov).copy()
        C, D = cov.shape
        nf, nc = ct.c_long(0), ct.c_long(0)
        self.lib.orc_variance_control(C, D, _d(cov), ct.c_double(flooring), ct.c_double(ceiling),
                                      _d(_f64(cov_signal)), ct.byref(nf), ct.byref(nc))
        return cov, nf.value, nc.value

    def set_it_parameter(self, begin, end, nb_it, it):
        return self.lib.orc_set_it_parameter(begin, end, nb_it, it)

    def mean_cov(self, X):
        X = _f32(X)
        D = X.shape[1]
        mean, cov = np.empty(D), np.empty(D)
        self.lib.orc_mean_cov(D, X.ctypes.data_as(c_fp), ct.c_size_t(X.shape[0]),
                              ct.c_size_t(X.strides[0] // 4), _d(mean), _d(cov))
        return mean, cov

    def llk_determine_top(self, g, X, K, complete=True, min_llk=-200.0, max_llk=200.0):
        X = _f32(X)
        T = X.shape[0]
        llk = np.empty(T)
        idx = np.empty((T, K), dtype=np.uint32)
        top_lk = np.empty((T, K))
        rest_lk, rest_w = np.empty(T), np.empty(T)
        self.lib.orc_llk_determine_top(*g.args(), X.ctypes.data_as(c_fp), ct.c_size_t(T),
                                       ct.c_size_t(X.strides[0] // 4), K, int(complete),
                                       ct.c_double(min_llk), ct.c_double(max_llk), _d(llk),
                                       idx.ctypes.data_as(c_up), _d(top_lk), _d(rest_lk),
                                       _d(rest_w))
        return llk, idx, top_lk, rest_lk, rest_w

    def llk_use_top(self, g, X, idx, rest_lk, complete=True, min_llk=-200.0, max_llk=200.0):
        X = _f32(X)
        T = X.shape[0]
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        K = idx.shape[1]
        llk = np.empty(T)
        rl = _f64(rest_lk)
        self.lib.orc_llk_use_top(*g.args(), X.ctypes.data_as(c_fp), ct.c_size_t(T),
                                 ct.c_size_t(X.strides[0] // 4), K, idx.ctypes.data_as(c_up),
                                 _d(rl), int(complete), ct.c_double(min_llk),
                                 ct.c_double(max_llk), _d(llk))
        return llk

    def llk_all(self, g, X, min_llk=-200.0, max_llk=200.0):
        X = _f32(X)
        T = X.shape[0]
        llk = np.empty(T)
        self.lib.orc_llk_all(*g.args(), X.ctypes.data_as(c_fp), ct.c_size_t(T),
                             ct.c_size_t(X.strides[0] // 4), ct.c_double(min_llk),
                             ct.c_double(max_llk), _d(llk))
        return llk

    # ---- Total variability -----------------------------------------------------
    def tv_subtract_m(self, N, F, ubm_mean):
        N, F = _f64(N), _f64(F).copy()
        U, C = N.shape
        D = F.shape[1] // C
        self.lib.orc_tv_subtract_m(ct.c_size_t(U), C, D, _d(N), _d(_f64(ubm_mean)), _d(F))
        return F

    def tv_tett(self, T, invvar, C, D, threads=1):
        T = _f64(T)
        R = T.shape[0]
        out = np.empty((C, R, R))
        self.lib.orc_tv_tett(C, D, R, _d(T), _d(_f64(invvar)), _d(out), threads)
        return out

    def tv_ivectors(self, N, F, T, invvar, tett, threads=1):
        N, F, T = _f64(N), _f64(F), _f64(T)
        U, C = N.shape
        R = T.shape[0]
        D = T.shape[1] // C
        W = np.empty((U, R))
        self.lib.orc_tv_ivectors(ct.c_size_t(U), C, D, R, _d(N), _d(F), _d(T), _d(_f64(invvar)),
                                 _d(_f64(tett)), _d(W), threads)
        return W

    def tv_estep(self, N, F, T, invvar, tett, Cmx=None, threads=1):
        N, F, T = _f64(N), _f64(F), _f64(T)
        U, C = N.shape
        R = T.shape[0]
        D = T.shape[1] // C
        W = np.empty((U, R))
        A = np.empty((C, R * R))
        Cmx = np.zeros((R, C * D)) if Cmx is None else Cmx
        Rm, r, meanW = np.empty((R, R)), np.empty(R), np.empty(R)
        self.lib.orc_tv_estep(ct.c_size_t(U), C, D, R, _d(N), _d(F), _d(T), _d(_f64(invvar)),
                              _d(_f64(tett)), _d(W), _d(A), _d(Cmx), _d(Rm), _d(r), _d(meanW),
                              threads)
        return W, A, Cmx, Rm, r, meanW

    def tv_mstep(self, A, Cmx, C, D):
        A, Cmx = _f64(A), _f64(Cmx)
        R = Cmx.shape[0]
        T = np.empty((R, C * D))
        self.lib.orc_tv_mstep(C, D, R, _d(A), _d(Cmx), _d(T))
        return T

    def tv_mindiv(self, Rm, r, meanW, ubm_mean, T, n_sessions, C, D):
        Rm, r, T = _f64(Rm).copy(), _f64(r).copy(), _f64(T).copy()
        mean = _f64(ubm_mean).copy()
        R = T.shape[0]
        rc = self.lib.orc_tv_mindiv(C, D, R, ct.c_double(n_sessions), _d(Rm), _d(r),
                                    _d(_f64(meanW)), _d(mean), _d(T))
        if rc != 0:
            raise ArithmeticError("upperCholesky failed")
        return mean, T

    def tv_orthonormalize(self, T):
        T = _f64(T).copy()
        self.lib.orc_tv_orthonormalize(T.shape[0], ct.c_size_t(T.shape[1]), _d(T))
        return T

    # ---- approximate i-vector modes ------------------------------------------
    def tv_norm_t(self, T, invvar):
        T = _f64(T).copy()
        self.lib.orc_tv_norm_t(T.shape[0], ct.c_size_t(T.shape[1]), _d(_f64(invvar)), _d(T))
        return T

    def tv_norm_statistics(self, N, F, ubm_mean, invvar):
        N, F = _f64(N), _f64(F).copy()
        U, C = N.shape
        D = F.size // (U * C)
        self.lib.orc_tv_norm_statistics(ct.c_size_t(U), C, D, _d(N), _d(_f64(ubm_mean)),
                                        _d(_f64(invvar)), _d(F))
        return F

    def tv_weighted_cov(self, T, weight, C, D):
        T = _f64(T)
        R = T.shape[0]
        W = np.empty((R, R))
        self.lib.orc_tv_weighted_cov(C, D, R, _d(T), _d(_f64(weight)), _d(W))
        return W

    def eigen_sym(self, EP, rank=None):
        EP = _f64(EP)
        n = EP.shape[0]
        rank = n if rank is None else rank
        vec, val = np.empty((n, rank)), np.empty(rank)
        self.lib.orc_eigen_sym(n, _d(EP), rank, _d(vec), _d(val))
        return vec, val

    def tv_approximate_tctc(self, T, Q, C, D):
        T, Q = _f64(T), _f64(Q)
        R = T.shape[0]
        Dm = np.empty((C, R))
        self.lib.orc_tv_approximate_tctc(C, D, R, _d(T), _d(Q), _d(Dm))
        return Dm

    def tv_ivectors_ubm_weight(self, N, F, T, Wcov):
        N, F, T = _f64(N), _f64(F), _f64(T)
        U, C = N.shape
        R = T.shape[0]
        D = T.shape[1] // C
        W = np.empty((U, R))
        self.lib.orc_tv_ivectors_ubm_weight(ct.c_size_t(U), C, D, R, _d(N), _d(F), _d(T),
                                            _d(_f64(Wcov)), _d(W))
        return W

    def tv_ivectors_eigen(self, N, F, T, Dm, Q, W0=None):
        N, F, T = _f64(N), _f64(F), _f64(T)
        U, C = N.shape
        R = T.shape[0]
        D = T.shape[1] // C
        W = np.zeros((U, R)) if W0 is None else _f64(W0).copy()
        self.lib.orc_tv_ivectors_eigen(ct.c_size_t(U), C, D, R, _d(N), _d(F), _d(T),
                                       _d(_f64(Dm)), _d(_f64(Q)), _d(W))
        return W

    # ---- i-vector back-end (PldaDev / PldaTest non-PLDA scorings) --------------
    def iv_compute_all(self, data, class_of, n_spk):
        data = _f64(data)
        d, n = data.shape
        cls = np.ascontiguousarray(class_of, dtype=np.int32)
        mean, sm = np.empty(d), np.empty((d, n_spk))
        self.lib.orc_iv_compute_all(d, ct.c_size_t(n), _d(data), cls.ctypes.data_as(c_ip),
                                    ct.c_size_t(n_spk), _d(mean), _d(sm))
        return mean, sm

    def iv_cov_mat(self, data, class_of, n_spk):
        data = _f64(data)
        d, n = data.shape
        cls = np.ascontiguousarray(class_of, dtype=np.int32)
        mean, sm = self.iv_compute_all(data, cls, n_spk)
        S, W, B = np.empty((d, d)), np.empty((d, d)), np.empty((d, d))
        self.lib.orc_iv_cov_mat(d, ct.c_size_t(n), _d(data), cls.ctypes.data_as(c_ip),
                                ct.c_size_t(n_spk), _d(mean), _d(sm), _d(S), _d(W), _d(B))
        return mean, sm, S, W, B

    def iv_wccn_chol(self, data, class_of, n_spk):
        data = _f64(data)
        d, n = data.shape
        cls = np.ascontiguousarray(class_of, dtype=np.int32)
        _, sm = self.iv_compute_all(data, cls, n_spk)
        out = np.empty((d, d))
        rc = self.lib.orc_iv_wccn_chol(d, ct.c_size_t(n), _d(data), cls.ctypes.data_as(c_ip),
                                       ct.c_size_t(n_spk), _d(sm), _d(out))
        if rc != 0:
            raise ArithmeticError("WCCN: singular within-class covariance")
        return out

    def iv_length_norm(self, data):
        data = _f64(data).copy()
        self.lib.orc_iv_length_norm(data.shape[0], ct.c_size_t(data.shape[1]), _d(data))
        return data

    def iv_center(self, data, mu):
        data = _f64(data).copy()
        self.lib.orc_iv_center(data.shape[0], ct.c_size_t(data.shape[1]), _d(_f64(mu)), _d(data))
        return data

    def iv_rotate_left(self, M, data):
        M, data = _f64(M), _f64(data)
        out = np.empty((M.shape[0], data.shape[1]))
        self.lib.orc_iv_rotate_left(M.shape[0], M.shape[1], ct.c_size_t(data.shape[1]), _d(M),
                                    _d(data), _d(out))
        return out

    def iv_efr_matrix(self, cov):
        cov = _f64(cov)
        out = np.empty_like(cov)
        if self.lib.orc_iv_efr_matrix(cov.shape[0], _d(cov), _d(out)) != 0:
            raise ArithmeticError("EFR: covariance is not positive definite")
        return out

    def iv_lda(self, W, B, rank):
        W, B = _f64(W), _f64(B)
        out = np.empty((rank, W.shape[0]))
        if self.lib.orc_iv_lda(W.shape[0], _d(W), _d(B), rank, _d(out)) != 0:
            raise ArithmeticError("LDA: singular within-class covariance")
        return out

    def _trials(self, trials, nm, nt):
        if trials is None:
            return None, None
        t = np.ascontiguousarray(trials, dtype=np.uint8)
        assert t.shape == (nm, nt)
        return t, t.ctypes.data_as(ct.POINTER(ct.c_uint8))

    def iv_cosine(self, models, segments, trials=None):
        models, segments = _f64(models), _f64(segments)
        d, nm = models.shape
        nt = segments.shape[1]
        keep, tp = self._trials(trials, nm, nt)
        sc = np.empty((nm, nt))
        self.lib.orc_iv_cosine(d, ct.c_size_t(nm), ct.c_size_t(nt), _d(models), _d(segments), tp, _d(sc))
        return sc

    def iv_mahalanobis(self, models, segments, Mah, trials=None):
        models, segments = _f64(models), _f64(segments)
        d, nm = models.shape
        nt = segments.shape[1]
        keep, tp = self._trials(trials, nm, nt)
        sc = np.empty((nm, nt))
        self.lib.orc_iv_mahalanobis(d, ct.c_size_t(nm), ct.c_size_t(nt), _d(models), _d(segments),
                                    _d(_f64(Mah)), tp, _d(sc))
        return sc

    def iv_two_cov(self, models, segments, W, B):
        models, segments = _f64(models), _f64(segments)
        d, nm = models.shape
        nt = segments.shape[1]
        sc = np.empty((nm, nt))
        if self.lib.orc_iv_two_cov(d, ct.c_size_t(nm), ct.c_size_t(nt), _d(models), _d(segments),
                                   _d(_f64(W)), _d(_f64(B)), _d(sc)) != 0:
            raise ArithmeticError("2cov: singular covariance")
        return sc

    # ---- PLDA ---------------------------------------------------------------
    def plda_native_scoring(self, F, G, Sigma, models, model_of, segments):
        F, Sigma = _f64(F), _f64(Sigma)
        d, rF = F.shape
        rG = 0 if G is None else G.shape[1]
        Gp = _d(_f64(G)) if rG else None
        models, segments = _f64(models), _f64(segments)
        model_of = np.ascontiguousarray(model_of, dtype=np.int32)
        n_models = len(np.unique(model_of))
        scores = np.empty((n_models, segments.shape[1]))
        rc = self.lib.orc_plda_native_scoring(d, rF, rG, _d(F), Gp, _d(Sigma), _d(models),
                                              ct.c_size_t(models.shape[1]),
                                              model_of.ctypes.data_as(c_ip),
                                              ct.c_size_t(n_models), _d(segments),
                                              ct.c_size_t(segments.shape[1]), _d(scores))
        if rc != 0:
            raise ArithmeticError("singular Sigma")
        return scores

    def plda_em_iteration(self, data, class_of, n_spk, F, G, Sigma, Delta):
        """One PldaModel::em_iteration -> (data centred, F, G, Sigma, Delta)."""
        data, F, Sigma, Delta = _f64(data).copy(), _f64(F).copy(), _f64(Sigma).copy(), _f64(Delta).copy()
        d, n = data.shape
        rF = F.shape[1]
        rG = 0 if G is None else G.shape[1]
        G = np.zeros((d, 0)) if G is None else _f64(G).copy()
        cls = np.ascontiguousarray(class_of, dtype=np.int32)
        rc = self.lib.orc_plda_em_iteration(d, rF, rG, ct.c_size_t(n), _d(data), cls.ctypes.data_as(c_ip),
                                            ct.c_size_t(n_spk), _d(F), _d(G) if rG else None, _d(Sigma), _d(Delta))
        if rc != 0:
            raise ArithmeticError("PLDA EM: singular matrix")
        return data, F, G, Sigma, Delta

    def invert(self, a):
        a = _f64(a)
        out = np.empty_like(a)
        self.lib.orc_invert(a.shape[0], _d(a), _d(out))
        return out

    def upper_cholesky(self, a):
        a = _f64(a)
        out = np.empty_like(a)
        rc = self.lib.orc_upper_cholesky(a.shape[0], _d(a), _d(out))
        if rc != 0:
            raise ArithmeticError("not positive definite")
        return out
