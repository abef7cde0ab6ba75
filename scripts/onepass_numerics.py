"""Numerical model (numpy, CPU) of the posterior path planned for the round-2 one-pass kernel
(DESIGN.md §4.6): per 128-component slice the posteriors are held as fp16(2^14 * 2^(S - m_slice))
while the log-sum-exp of the 16 slices is exchanged, then rescaled by f = 2^(m_slice - lse) with one
fp16 multiply; the slice sums are taken in 24-bit fixed point (redux.sync.add.u32).
Compared with today's path, fp16(2^14 * 2^(S - lse)) (one rounding), and with exact posteriors."""
import numpy as np

rng = np.random.default_rng(0)
T, C, SL = 4096, 2048, 128
# log2 joint likelihoods with a realistic spread: a few competitive components per frame
S = -120.0 - np.abs(rng.standard_normal((T, C))) * 40.0
top = rng.integers(0, C, (T, 6))
S[np.arange(T)[:, None], top] = -110.0 + rng.standard_normal((T, 6)) * 3.0
S = S.astype(np.float32)

lse = (S.max(1) + np.log2(np.exp2(S - S.max(1, keepdims=True)).astype(np.float64).sum(1))).astype(np.float64)
gamma = np.exp2(S.astype(np.float64) - lse[:, None])

# today: one rounding
p_now = np.exp2((S - lse[:, None].astype(np.float32) + 14).astype(np.float32)).astype(np.float16).astype(np.float64) / 2 ** 14

# planned: slice-local reference, fixed-point slice sums, delayed fp16 rescale
Ss = S.reshape(T, C // SL, SL)
m_loc = Ss.max(2)
e = np.exp2(Ss - m_loc[:, :, None]).astype(np.float32)
z_fix = np.round(e.astype(np.float64) * 2 ** 24).sum(2) / 2 ** 24                  # redux.sync.add.u32
m_glob = m_loc.max(1)
lse_new = m_glob + np.log2((z_fix * np.exp2((m_loc - m_glob[:, None]).astype(np.float64))).sum(1))
e16 = (e * np.float32(2 ** 14)).astype(np.float16)
d = (m_loc - lse_new[:, None]).astype(np.float32)                                  # <= 0
f = np.exp2(d).astype(np.float16)                                                   # naive: one fp16 factor
p_naive = (e16 * f[:, :, None]).astype(np.float16).astype(np.float64).reshape(T, C) / 2 ** 14
# a factor below 2^-14 is a SUBNORMAL fp16 with few significand bits: split it into a mantissa in
# [0.5, 1] and an exact power of two (two HMUL2 instead of one)
k = np.ceil(d)
f_m = np.exp2(d - k).astype(np.float16)
f_p = np.exp2(np.maximum(k, -24.0)).astype(np.float16)
p_new = ((e16 * f_m[:, :, None]).astype(np.float16) * f_p[:, :, None]).astype(np.float16).astype(np.float64).reshape(T, C) / 2 ** 14


def report(name, p):
    big = gamma > 1e-6
    rel = np.abs(p - gamma)[big] / gamma[big]
    print(f"{name:28s} rel.err of posteriors > 1e-6: rms {np.sqrt((rel ** 2).mean()):.2e} max {rel.max():.2e}; "
          f"|sum_c p - 1| max {np.abs(p.sum(1) - 1).max():.2e}; N_c rel.err max "
          f"{(np.abs(p.sum(0) - gamma.sum(0)) / np.maximum(gamma.sum(0), 1e-9)).max():.2e}")


print(f"lse: |new - exact| max {np.abs(lse_new - lse).max():.2e} log2 units")
report("today (one fp16 rounding)", p_now)
report("planned, one fp16 factor", p_naive)
report("planned, mantissa x 2^k", p_new)
flush = (gamma > 2.0 ** -38) & (p_new == 0)
print(f"posteriors > 2^-38 flushed to zero by the delayed rescale: {flush.sum()} of {(gamma > 2.0 ** -38).sum()}")
