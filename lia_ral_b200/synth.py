"""Seeded synthetic workloads: UBMs, frames, BW statistics, PLDA models.

numpy's Philox counter-based generator everywhere (never libc rand()).  Used by tests/ and
bench.py on both the GPU arm and the CPU-oracle arm so both see identical inputs.
"""
import numpy as np


def _rng(seed):
    return np.random.Generator(np.random.Philox(seed))


def make_ubm(C=2048, D=60, seed=1):
    """means ~ N(0, 2^2); variances ~ LogNormal(0, 0.5^2) clipped to [0.05, 20];
    weights = softmax(N(0,1)) floored at 1e-5 and renormalised."""
    g = _rng(seed)
    mean = g.normal(0.0, 2.0, size=(C, D))
    cov = np.clip(np.exp(g.normal(0.0, 0.5, size=(C, D))), 0.05, 20.0)
    z = g.normal(0.0, 1.0, size=C)
    w = np.exp(z - z.max())
    w /= w.sum()
    w = np.maximum(w, 1e-5)
    w /= w.sum()
    return w, mean, cov


def make_frames(w, mean, cov, T, seed=2, chunk=1 << 20):
    """Draw component ~ weights, x = mu_c + sigma_c * N(0,1); float32 row-major [T, D]."""
    g = _rng(seed)
    C, D = mean.shape
    X = np.empty((T, D), dtype=np.float32)
    sd = np.sqrt(cov)
    cdf = np.cumsum(w)
    cdf[-1] = 1.0
    for s in range(0, T, chunk):
        n = min(chunk, T - s)
        comp = np.searchsorted(cdf, g.random(n), side="right").clip(0, C - 1)
        X[s:s + n] = (mean[comp] + sd[comp] * g.standard_normal((n, D))).astype(np.float32)
    return X


def perturb_ubm(w, mean, cov, seed=3, frac=0.1, scale=0.3):
    """Client model / perturbed EM start: means of `frac` of the components moved by N(0, scale^2)."""
    g = _rng(seed)
    C, D = mean.shape
    m = mean.copy()
    pick = g.random(C) < frac
    m[pick] += g.normal(0.0, scale, size=(int(pick.sum()), D))
    return w.copy(), m, cov.copy()


def make_T(R, C, D, invvar, seed=4, scale=None):
    """T ~ N(0,1) * (sum invvar) * 1e-3, the scale TVAcc::initT uses (AccumulateTVStat.cpp:733-746);
    `scale` replaces that factor (a trained T has entries of a few 1e-2 sigma)."""
    g = _rng(seed)
    if scale is None:
        scale = float(np.sum(invvar)) * 1e-3
    return g.standard_normal((R, C * D)) * scale


def make_bw_stats(U, w, mean, cov, R=None, frames_per_utt=3000, active=64, seed=5):
    """Synthesised BW statistics for TV EM (cfg4): sparse Dirichlet occupancies summing to
    frames_per_utt, F_u = N_u o (mu + noise). Returns N [U,C], F [U,C*D] float64."""
    g = _rng(seed)
    C, D = mean.shape
    N = np.zeros((U, C))
    F = np.zeros((U, C * D))
    sd = np.sqrt(cov)
    for u in range(U):
        act = g.choice(C, size=min(active, C), replace=False)
        occ = g.dirichlet(np.full(len(act), 0.5)) * frames_per_utt
        N[u, act] = occ
        off = g.normal(0.0, 0.3, size=D)  # utterance-level shift shared across components
        mu = mean[act] + off[None, :] * sd[act]
        mu += g.standard_normal((len(act), D)) * sd[act] / np.sqrt(np.maximum(occ, 1.0))[:, None]
        Fu = np.zeros((C, D))
        Fu[act] = occ[:, None] * mu
        F[u] = Fu.reshape(-1)
    return N, F


def make_plda(d=400, rF=200, rG=0, n_models=64, n_test=128, sessions=None, seed=6):
    """Random PLDA model (F, G, SPD Sigma) + i-vectors. sessions: list of enrolment counts."""
    g = _rng(seed)
    F = g.standard_normal((d, rF)) / np.sqrt(d)
    G = g.standard_normal((d, rG)) / np.sqrt(d) if rG else None
    A = g.standard_normal((d, d)) / np.sqrt(d)
    Sigma = A @ A.T * 0.5 + np.eye(d) * 0.5
    if sessions is None:
        sessions = [1] * n_models
    model_of = np.concatenate([np.full(s, i, dtype=np.int32) for i, s in enumerate(sessions)])
    spk = g.standard_normal((rF, len(sessions)))
    models = F @ spk[:, model_of] + 0.3 * g.standard_normal((d, len(model_of)))
    tspk = g.standard_normal((rF, n_test))
    segments = F @ tspk + 0.3 * g.standard_normal((d, n_test))
    return F, G, Sigma, models, model_of, segments
