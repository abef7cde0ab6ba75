"""Independent numpy/scipy formulation of the same path -- cross-checks lr_oracle.c.

TEST INFRASTRUCTURE ONLY.  Vectorised, log-domain where the C restatement is linear-domain,
solve-based where the C restatement inverts explicitly: two routes to the same numbers.
"""
import numpy as np


def compute_all(cov):
    D = cov.shape[1]
    det = np.prod(cov, axis=1)
    return 1.0 / cov, det, 1.0 / ((2 * np.pi) ** (D / 2) * np.sqrt(det))


def log_joint(w, mean, cov, X):
    """log(w_c * lk_c(x)) for all frames/components, [T, C] (log domain)."""
    X = np.asarray(X, dtype=np.float64)
    D = X.shape[1]
    a = 1.0 / cov
    const = np.log(w) - 0.5 * (D * np.log(2 * np.pi) + np.log(cov).sum(1)) - 0.5 * (mean ** 2 * a).sum(1)
    return const[None, :] + X @ (mean * a).T - 0.5 * (X ** 2) @ a.T


def posteriors(w, mean, cov, X):
    lj = log_joint(w, mean, cov, X)
    m = lj.max(1, keepdims=True)
    p = np.exp(lj - m)
    s = p.sum(1, keepdims=True)
    return p / s, (m + np.log(s))[:, 0]


def bwstats(w, mean, cov, X, frame2row, U):
    g, _ = posteriors(w, mean, cov, X)
    C, D = mean.shape
    N = np.zeros((U, C))
    F = np.zeros((U, C, D))
    Xd = np.asarray(X, dtype=np.float64)
    for u in range(U):
        sel = frame2row == u
        N[u] = g[sel].sum(0)
        F[u] = g[sel].T @ Xd[sel]
    return N, F.reshape(U, C * D)


def em_stats(w, mean, cov, X):
    g, llk = posteriors(w, mean, cov, X)
    Xd = np.asarray(X, dtype=np.float64)
    return llk.sum(), g.sum(0), g.T @ Xd, g.T @ (Xd ** 2)


def ivectors(N, F, T, invvar, ubm_mean=None):
    """w_u = (I + T diag(N_u (x) invvar) T^T)^-1 T (invvar o F_u) via solve()."""
    U, C = N.shape
    R, sv = T.shape
    D = sv // C
    W = np.zeros((U, R))
    Linvs = []
    for u in range(U):
        nn = np.repeat(N[u], D) * invvar
        L = np.eye(R) + (T * nn[None, :]) @ T.T
        b = T @ (invvar * F[u])
        W[u] = np.linalg.solve(L, b)
        Linvs.append(np.linalg.inv(L))
    return W, Linvs


def tv_estep(N, F, T, invvar):
    U, C = N.shape
    R, sv = T.shape
    W, Linvs = ivectors(N, F, T, invvar)
    A = np.zeros((C, R, R))
    Cmx = np.zeros((R, sv))
    Rm = np.zeros((R, R))
    for u in range(U):
        E = Linvs[u] + np.outer(W[u], W[u])
        A += N[u][:, None, None] * E[None]
        Cmx += np.outer(W[u], F[u])
        Rm += E
    return W, A.reshape(C, R * R), Cmx, Rm, W.sum(0), W.mean(0)


def tv_mstep(A, Cmx, C, D):
    R = Cmx.shape[0]
    T = np.zeros_like(Cmx)
    for c in range(C):
        T[:, c * D:(c + 1) * D] = np.linalg.solve(A[c].reshape(R, R), Cmx[:, c * D:(c + 1) * D])
    return T


def plda_scores(F, G, Sigma, models, model_of, segments):
    """Two-covariance closed form expanded: only the cross term needs an [Nt x r][r x Nm] product."""
    iS = np.linalg.inv(Sigma)
    if G is not None and G.shape[1] > 0:
        J = iS - iS @ G @ np.linalg.inv(G.T @ iS @ G + np.eye(G.shape[1])) @ G.T @ iS
    else:
        J = iS
    FTJ = F.T @ J
    phi = FTJ @ F
    r = phi.shape[0]
    pm, ps = FTJ @ models, FTJ @ segments
    ids = np.unique(model_of)
    out = np.zeros((len(ids), segments.shape[1]))
    K = lambda n: np.linalg.inv(n * phi + np.eye(r))
    ld = lambda M: np.linalg.slogdet(M)[1]
    K1 = K(1)
    for i, mid in enumerate(ids):
        cols = np.nonzero(model_of == mid)[0]
        L = len(cols)
        m = pm[:, cols].sum(1)
        KL, KL1 = K(L), K(L + 1)
        const = 0.5 * (ld(KL1) - ld(KL) - ld(K1))
        a = 0.5 * m @ (KL1 - KL) @ m + const
        b = 0.5 * np.einsum("it,ij,jt->t", ps, KL1 - K1, ps)
        out[i] = ps.T @ (KL1 @ m) + a + b
    return out


def map_occ_dep(w0, mean0, cov0, w_ml, mean_ml, cov_ml, frame_count, r_mean=None, r_var=None, r_weight=None):
    """computeMAPOccDep (TrainTools.cpp:445-489): occupation-dependent MAP of the ML estimate
    (w_ml, mean_ml, cov_ml) towards the a-priori model; r_* = None disables that adaptation."""
    alpha = w_ml * frame_count
    w, mean, cov = w0.copy(), mean0.copy(), cov0.copy()
    if r_mean is not None:
        a = (alpha / (alpha + r_mean))[:, None]
        mean = (1 - a) * mean0 + a * mean_ml
    if r_var is not None:
        a = (alpha / (alpha + r_var))[:, None]
        cov = (1 - a) * cov0 + a * cov_ml + (1 - a) * a * (mean0 - mean_ml) ** 2
    if r_weight is not None:
        a = alpha / (alpha + r_weight)
        w = a * w_ml + (1 - a) * w0
        w = w / w.sum()
    return w, mean, cov
