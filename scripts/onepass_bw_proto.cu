// onepass_bw_proto.cu -- stand-alone prototype of the ONE-PASS likelihood + Baum-Welch statistics
// kernel planned for round 2 (cluster-free variant): the per-frame
// log-sum-exp is no longer precomputed by a first pass over the frames, it is exchanged between the
// slice-CTAs of a frame group while the posteriors wait in TMEM.
//
// Per CTA (slice of 128 components, one frame group) and half tile h (64 frames):
//   G1(h)   : S[c, t] = W A^T, HI weights in TMEM (TS UMMA), LO weights SS      (as csrc/gmm_tc.cu pass 2)
//   epi-1(h): slice-local per-frame max / sum over the 128 lanes (redux.sync on the ordered-int image,
//             24-bit fixed-point sums), (max, sum) -> global, release counter; posteriors relative to
//             the slice max as fp16(2^14 e) into one of 6 waiting TMEM slots
//   epi-2(h): acquire the counter of the group, combine the 16 partials -> lse, rescale the waiting
//             posteriors by 2^(m_slice - lse) (fp16 mantissa x exact power of two)
//   G2(h)   : F[c, :] += P[c, t] A[t, :]  (TS UMMA, 128 accumulator columns = [xh hi, 1 | xh lo])
// TMEM: S 2 x 64 | P 6 x 32 | HI weights 64 | accumulator 128 = 512 columns.
// Each epilogue team runs epi-1(h) then epi-2 of its PREVIOUS half tile, so an exchange has two
// half-tile periods to complete before anybody waits on it.
//
// Synthetic 2048c/60d model and frames are built on the host in the library's operand layout (swizzled
// fp16 hi / lo panels); the kernel's N / F accumulators and log-sum-exps are checked against fp64.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/onepass_bw_proto scripts/onepass_bw_proto.cu
//   ./scripts/onepass_bw_proto [tiles_per_group]
//
// Written and compile-checked in round 1 without GPU time left to run it: expect bring-up work.
// Not part of the product library.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int kTile = 128, kHF = 64, kSlice = 128, kD = 60, kOneCol = 60;
constexpr int kPanelBytes = 128 * 128, kTileBytes = 4 * kPanelBytes;
constexpr int kHalfPanel = kPanelBytes / 2, kHalfBytes = 4 * kHalfPanel;
constexpr int kHStages = 5, kNP = 6;
constexpr int kThreads = 384;
constexpr int kColS = 0, kColP = 128, kColW = 320, kColAcc = 384;  // TMEM column map
constexpr size_t kSmem = 1024 + 64 * 1024 + kHStages * (size_t)kHalfBytes + 512;

// ---- PTX wrappers (the validated set of csrc/gmm_tc.cu) -------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, 0xffffffff;\n"
      "selp.b32 %0, 1, 0, px;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint64_t desc_add(uint64_t d, uint32_t bytes) { return d + (uint64_t)(bytes >> 4); }
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__host__ __device__ __forceinline__ uint32_t panel_off(int row, int col) {
  return (uint32_t)row * 128u + (uint32_t)((((col >> 3) ^ (row & 7)) << 4) | ((col & 7) << 1));
}
__device__ __forceinline__ int f2ord(float f) {
  int i = __float_as_int(f);
  return i ^ ((i >> 31) & 0x7FFFFFFF);
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7FFFFFFF)); }

struct Bars {
  uint32_t full[kHStages], empty[kHStages], s_full[2], s_free[2], p_ready[kNP], p_free[kNP], w_full, w_tmem, f_full,
      tmem_slot;
};

// The likelihood GEMM of one half tile (24 UMMAs): HI weights from TMEM (TS), LO weights from smem (SS)
__device__ __forceinline__ void issue_g1_ts(uint32_t d_tmem, uint32_t w_tmem, uint64_t wlo_desc0, uint64_t x_desc0) {
  constexpr uint32_t idesc = make_idesc(128, 64, 0, 0);
  constexpr int kXPanel = 64 * 128;
  constexpr int wp[6] = {0, 1, 0, 1, 0, 1};
  constexpr int xp[6] = {0, 1, 2, 3, 0, 1};
  uint32_t acc = 0;
#pragma unroll
  for (int q = 0; q < 6; q++) {
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
      uint64_t xd = desc_add(x_desc0, xp[q] * kXPanel + kk * 32);
      if (q < 4)
        umma_ts(d_tmem, w_tmem + wp[q] * 32 + kk * 8, xd, idesc, acc);
      else
        umma_ss(d_tmem, desc_add(wlo_desc0, wp[q] * kPanelBytes + kk * 32), xd, idesc, acc);
      acc = 1;
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1)
k_onepass_bw(int n_slices, const unsigned char *__restrict__ Wp, const unsigned char *__restrict__ Xh,
             const int *__restrict__ group_tiles, long P_pad, float2 *part /*[slices][P_pad]*/,
             unsigned *cnt /*[groups][max_half]*/, int max_half, float *__restrict__ lse_out /*[P_pad]*/,
             float *__restrict__ acc_out /*[groups][slices][128][128]*/) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char *base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t s_w = base;
  uint32_t hstage[kHStages];
  for (int i = 0; i < kHStages; i++) hstage[i] = base + 64 * 1024 + i * kHalfBytes;
  Bars sm;
  {
    uint32_t b = base + 64 * 1024 + kHStages * kHalfBytes;
    for (int i = 0; i < kHStages; i++) {
      sm.full[i] = b;
      b += 8;
      sm.empty[i] = b;
      b += 8;
    }
    for (int i = 0; i < 2; i++) {
      sm.s_full[i] = b;
      b += 8;
      sm.s_free[i] = b;
      b += 8;
    }
    for (int i = 0; i < kNP; i++) {
      sm.p_ready[i] = b;
      b += 8;
      sm.p_free[i] = b;
      b += 8;
    }
    sm.w_full = b;
    sm.w_tmem = b + 8;
    sm.f_full = b + 16;
    sm.tmem_slot = b + 24;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.x % n_slices, group = blockIdx.x / n_slices;
  const int t_begin = group_tiles[group], t_end = group_tiles[group + 1];
  const int n_half = 2 * (t_end - t_begin);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kHStages; i++) {
      mbar_init(sm.full[i], 1);
      mbar_init(sm.empty[i], 1);
    }
    for (int i = 0; i < 2; i++) {
      mbar_init(sm.s_full[i], 1);
      mbar_init(sm.s_free[i], 4);
    }
    for (int i = 0; i < kNP; i++) {
      mbar_init(sm.p_ready[i], 4);
      mbar_init(sm.p_free[i], 1);
    }
    mbar_init(sm.w_full, 1);
    mbar_init(sm.w_tmem, 8);
    mbar_init(sm.f_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc(sm.tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(sm.tmem_slot));

  if (warp == 0) {
    // ---- bulk-copy producer
    const bool leader = elect_one();
    if (leader) {
      mbar_expect_tx(sm.w_full, 4 * kPanelBytes);
      for (int p = 0; p < 4; p++)
        bulk_g2s(s_w + p * kPanelBytes, Wp + (size_t)slice * 4 * kPanelBytes + (size_t)p * kPanelBytes, kPanelBytes,
                 sm.w_full);
    }
    for (int h = 0; h < n_half; h++) {
      const int st = h % kHStages;
      mbar_wait(sm.empty[st], ((h / kHStages) & 1) ^ 1);
      if (leader) {
        mbar_expect_tx(sm.full[st], kHalfBytes);
        const unsigned char *src = Xh + (size_t)(t_begin + (h >> 1)) * kTileBytes + (size_t)(h & 1) * kHalfPanel;
        for (int p = 0; p < 4; p++)
          bulk_g2s(hstage[st] + p * kHalfPanel, src + (size_t)p * kPanelBytes, kHalfPanel, sm.full[st]);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---- likelihood-GEMM issuer: at most two half tiles ahead of epi-1 (two S buffers)
    const bool leader = elect_one();
    const uint64_t wlo_desc0 = make_desc(s_w + 2 * kPanelBytes, 16, 1024);
    const uint64_t x_desc0 = make_desc(hstage[0], 16, 1024);
    mbar_wait(sm.w_full, 0);
    mbar_wait(sm.w_tmem, 0);
    for (int h = 0; h < n_half; h++) {
      const int st = h % kHStages, sb = h & 1;
      mbar_wait(sm.full[st], (h / kHStages) & 1);
      if (h >= 2) mbar_wait(sm.s_free[sb], ((h >> 1) - 1) & 1);
      tc_fence_after();
      if (leader) {
        issue_g1_ts(tmem_base + kColS + sb * kHF, tmem_base + kColW, wlo_desc0, desc_add(x_desc0, st * kHalfBytes));
        umma_commit(sm.s_full[sb]);
      }
      __syncwarp();
    }
  } else if (warp == 3) {
    // ---- statistics-GEMM issuer: F[c, :] (+)= P[c, t] A[t, :], B = [xh hi, 1 | xh lo] read MN-major
    const bool leader = elect_one();
    constexpr uint32_t idesc2 = make_idesc(128, 128, 0, 1);
    const uint64_t b_desc0 = make_desc(hstage[0], 2u * kHalfPanel, 1024);
    for (int h = 0; h < n_half; h++) {
      const int st = h % kHStages, ps = h % kNP;
      mbar_wait(sm.full[st], (h / kHStages) & 1);
      mbar_wait(sm.p_ready[ps], (h / kNP) & 1);
      tc_fence_after();
      if (leader) {
        const uint64_t bd0 = desc_add(b_desc0, st * kHalfBytes);
#pragma unroll
        for (int kk = 0; kk < kHF / 16; kk++)
          umma_ts(tmem_base + kColAcc, tmem_base + kColP + ps * 32 + kk * 8, desc_add(bd0, kk * 2048), idesc2,
                  (h > 0 || kk > 0) ? 1u : 0u);
        umma_commit(sm.empty[st]);
        umma_commit(sm.p_free[ps]);
        if (h == n_half - 1) umma_commit(sm.f_full);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ---- two epilogue teams of four warps (one per TMEM lane quarter); team t owns the half tiles
    // h = t, t + 2, ...
    const int q = warp & 3, team = (warp - 4) >> 2;
    const int et = threadIdx.x - 128 - team * 128;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    // team-local exchange area inside the HI-weights region (free once they live in TMEM): panel `team`
    unsigned char *xa = base_ptr + (size_t)team * kPanelBytes;
    float *wmax = reinterpret_cast<float *>(xa);                  // [64][4]
    unsigned *wsum = reinterpret_cast<unsigned *>(xa + 1024);      // [64][4]
    __half *fm = reinterpret_cast<__half *>(xa + 2048);            // [64] mantissa of the rescale factor
    __half *fp = reinterpret_cast<__half *>(xa + 2048 + 128);      // [64] power of two of the rescale factor
    {
      mbar_wait(sm.w_full, 0);
      const int r = q * 32 + lane;
#pragma unroll
      for (int half16 = 0; half16 < 2; half16++) {
        uint32_t v[16];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int chunk = half16 * 4 + j;
          const uint32_t a = s_w + team * kPanelBytes + r * 128 + ((chunk ^ (r & 7)) << 4);
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(v[4 * j]), "=r"(v[4 * j + 1]), "=r"(v[4 * j + 2]), "=r"(v[4 * j + 3])
                       : "r"(a));
        }
        tmem_st16(tmem_base + lane_addr + kColW + team * 32 + half16 * 16, v);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sm.w_tmem);
      named_bar_sync(1 + team, 128);  // the team's panel is read completely before it is reused below
    }
    const long frame0 = (long)t_begin * kTile;
    unsigned *my_cnt = cnt + (size_t)group * max_half;

    auto epi1 = [&](int h) {
      const int sb = h & 1, ps = h % kNP;
      mbar_wait(sm.s_full[sb], (h >> 1) & 1);
      if (h >= kNP) mbar_wait(sm.p_free[ps], ((h / kNP) - 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int hf = 0; hf < 2; hf++) {  // 32 frame columns at a time (register budget)
        uint32_t r[32];
        tmem_ld32(tmem_base + lane_addr + kColS + sb * kHF + hf * 32, r);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; j++) {
          const int m = __reduce_max_sync(0xFFFFFFFFu, f2ord(__uint_as_float(r[j])));
          if (lane == j) wmax[(hf * 32 + j) * 4 + q] = ord2f(m);
        }
        named_bar_sync(1 + team, 128);
#pragma unroll
        for (int j = 0; j < 32; j++) {
          const float4 v = *reinterpret_cast<const float4 *>(wmax + (hf * 32 + j) * 4);
          const float m = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
          const float e = ex2f(__uint_as_float(r[j]) - m);
          r[j] = __float_as_uint(e);
          const unsigned t = __reduce_add_sync(0xFFFFFFFFu, __float2uint_rn(e * 16777216.f));
          if (lane == j) wsum[(hf * 32 + j) * 4 + q] = t;
        }
        uint32_t pk[16];
#pragma unroll
        for (int e2 = 0; e2 < 16; e2++) {
          __half2 hh = __floats2half2_rn(__uint_as_float(r[2 * e2]) * 16384.f, __uint_as_float(r[2 * e2 + 1]) * 16384.f);
          pk[e2] = *reinterpret_cast<uint32_t *>(&hh);
        }
        tmem_st16(tmem_base + lane_addr + kColP + ps * 32 + hf * 16, pk);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sm.s_free[sb]);
      named_bar_sync(1 + team, 128);  // every warp's sums are in shared memory
      if (et < kHF) {
        const float4 v = *reinterpret_cast<const float4 *>(wmax + et * 4);
        const uint4 s4 = *reinterpret_cast<const uint4 *>(wsum + et * 4);
        const float m = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
        const float z = (float)((double)s4.x + (double)s4.y + (double)s4.z + (double)s4.w) * (1.f / 16777216.f);
        part[(size_t)slice * P_pad + frame0 + (long)h * kHF + et] = make_float2(m, z);
        __threadfence();
      }
      named_bar_sync(1 + team, 128);  // partials written (and fenced) before the counter moves
      if (et == 0) {
        __threadfence();
        atomicAdd(my_cnt + h, 1u);
      }
    };

    auto epi2 = [&](int h) {
      const int ps = h % kNP;
      if (et < kHF) {
        while (ld_acquire_u32(my_cnt + h) < (unsigned)n_slices) {
        }
        const long f = frame0 + (long)h * kHF + et;
        float m = -3.0e38f;
        for (int sl = 0; sl < n_slices; sl++) m = fmaxf(m, __ldcg(&part[(size_t)sl * P_pad + f]).x);
        float z = 0.f;
        for (int sl = 0; sl < n_slices; sl++) {
          const float2 v = __ldcg(&part[(size_t)sl * P_pad + f]);
          z += v.y * ex2f(v.x - m);
        }
        const float lse = m + log2f(z);
        if (slice == 0) lse_out[f] = lse;
        const float d = __ldcg(&part[(size_t)slice * P_pad + f]).x - lse;  // <= 0
        const float k = ceilf(d);
        fm[et] = __float2half_rn(ex2f(d - k));
        fp[et] = __float2half_rn(ex2f(fmaxf(k, -24.f)));
      }
      named_bar_sync(1 + team, 128);
      uint32_t r[32];
      tmem_ld32(tmem_base + lane_addr + kColP + ps * 32, r);
      tmem_wait_ld();
      const __half2 *fm2 = reinterpret_cast<const __half2 *>(fm), *fp2 = reinterpret_cast<const __half2 *>(fp);
#pragma unroll
      for (int e2 = 0; e2 < 32; e2++) {  // packed column e2 holds the frames (2 e2, 2 e2 + 1)
        __half2 p = *reinterpret_cast<__half2 *>(&r[e2]);
        p = __hmul2(__hmul2(p, fm2[e2]), fp2[e2]);
        r[e2] = *reinterpret_cast<uint32_t *>(&p);
      }
      {
        uint32_t lo16[16], hi16[16];
#pragma unroll
        for (int e2 = 0; e2 < 16; e2++) {
          lo16[e2] = r[e2];
          hi16[e2] = r[16 + e2];
        }
        tmem_st16(tmem_base + lane_addr + kColP + ps * 32, lo16);
        tmem_st16(tmem_base + lane_addr + kColP + ps * 32 + 16, hi16);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sm.p_ready[ps]);
      named_bar_sync(1 + team, 128);  // fm / fp may be rewritten by the next epi-2
    };

    int prev = -1;
    for (int h = team; h < n_half; h += 2) {
      epi1(h);
      if (prev >= 0) epi2(prev);
      prev = h;
    }
    if (prev >= 0) epi2(prev);

    // ---- dump the accumulator of the run: lane = component, 128 statistics columns
    if (n_half > 0) {
      mbar_wait(sm.f_full, 0);
      tc_fence_after();
      float *out = acc_out + (((size_t)group * n_slices + slice) * kSlice + q * 32 + lane) * 128;
      for (int ch = team * 2; ch < team * 2 + 2; ch++) {
        uint32_t r[32];
        tmem_ld32(tmem_base + lane_addr + kColAcc + ch * 32, r);
        tmem_wait_ld();
#pragma unroll
        for (int e2 = 0; e2 < 32; e2++) out[ch * 32 + e2] = __uint_as_float(r[e2]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------ host: synthetic problem + check
#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e__ = (x);                                                            \
    if (e__ != cudaSuccess) {                                                         \
      std::printf("%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e__)); \
      return 1;                                                                       \
    }                                                                                 \
  } while (0)

static void put_split(unsigned char *panel_hi, unsigned char *panel_lo, int row, int col, double v, bool lo_too = true) {
  const __half hi = __float2half_rn((float)v);
  const __half lo = __float2half_rn((float)(v - (double)__half2float(hi)));
  *reinterpret_cast<__half *>(panel_hi + panel_off(row, col)) = hi;
  if (panel_lo && lo_too) *reinterpret_cast<__half *>(panel_lo + panel_off(row, col)) = lo;
}

int main(int argc, char **argv) {
  const int C = 2048, D = kD, n_slices = C / kSlice;
  int dev_sms = 0;
  CK(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0));
  const int groups = std::max(1, dev_sms / n_slices);
  const int tiles_per_group = argc > 1 ? std::atoi(argv[1]) : 16;  // <= 64: fp32 accumulator run length
  const int n_tiles = groups * tiles_per_group;
  const long P = (long)n_tiles * kTile;
  std::printf("one-pass BW prototype: %d components, %d groups x %d tiles (%ld frames), %d CTAs\n", C, groups,
              tiles_per_group, P, groups * n_slices);

  uint64_t s = 0x9E3779B97F4A7C15ull;
  auto rnd = [&]() {
    s ^= s << 13;
    s ^= s >> 7;
    s ^= s << 17;
    return (double)(s >> 11) / 9007199254740992.0;
  };
  auto gauss = [&]() { return std::sqrt(-2.0 * std::log(rnd() + 1e-300)) * std::cos(6.283185307179586 * rnd()); };
  std::vector<double> mean((size_t)C * D), var((size_t)C * D), lw(C);
  for (auto &v : mean) v = gauss();
  for (auto &v : var) v = 0.5 + 1.5 * rnd();
  for (auto &v : lw) v = std::log2(1.0 / C);
  const double kHalfLog2e = 0.72134752044448170368;
  // weights panels: per slice [hi a | hi b | lo a | lo b]; normalisation g = 0, s = 1 (xh = x)
  std::vector<unsigned char> W((size_t)n_slices * 4 * kPanelBytes, 0);
  for (int c = 0; c < C; c++) {
    unsigned char *b = W.data() + (size_t)(c / kSlice) * 4 * kPanelBytes;
    const int row = c % kSlice;
    double kc = lw[c];
    for (int i = 0; i < D; i++) {
      const double iv = 1.0 / var[(size_t)c * D + i], mu = mean[(size_t)c * D + i];
      kc += -0.5 * std::log2(6.283185307179586 * var[(size_t)c * D + i]) - kHalfLog2e * iv * mu * mu;
      put_split(b, b + 2 * kPanelBytes, row, i, 2.0 * kHalfLog2e * iv * mu);       // coefficient of xh
      put_split(b + kPanelBytes, b + 3 * kPanelBytes, row, i, -kHalfLog2e * iv);   // coefficient of xh^2
    }
    const __half k0 = __float2half_rn((float)kc);
    const double r1 = kc - (double)__half2float(k0);
    const __half k1 = __float2half_rn((float)r1);
    const __half k2 = __float2half_rn((float)(r1 - (double)__half2float(k1)));
    *reinterpret_cast<__half *>(b + panel_off(row, kOneCol)) = k0;
    *reinterpret_cast<__half *>(b + panel_off(row, kOneCol + 1)) = k1;
    *reinterpret_cast<__half *>(b + panel_off(row, kOneCol + 2)) = k2;
  }
  // frames: x = mean_c + sigma_c N(0, 1); tiles of four panels [xh hi, 1 | xh^2 hi | xh lo | xh^2 lo]
  std::vector<float> X((size_t)P * D);
  std::vector<unsigned char> Xh((size_t)n_tiles * kTileBytes, 0);
  for (long p = 0; p < P; p++) {
    const int c = (int)(rnd() * C) % C;
    unsigned char *t = Xh.data() + (size_t)(p / kTile) * kTileBytes;
    const int row = (int)(p % kTile);
    for (int i = 0; i < D; i++) {
      const float x = (float)(mean[(size_t)c * D + i] + std::sqrt(var[(size_t)c * D + i]) * gauss());
      X[(size_t)p * D + i] = x;
      put_split(t, t + 2 * kPanelBytes, row, i, (double)x);
      put_split(t + kPanelBytes, t + 3 * kPanelBytes, row, i, (double)x * (double)x);
    }
    for (int k = kOneCol; k < kOneCol + 3; k++) *reinterpret_cast<__half *>(t + panel_off(row, k)) = __float2half_rn(1.f);
  }
  std::vector<int> cuts(groups + 1);
  for (int g = 0; g <= groups; g++) cuts[g] = g * tiles_per_group;
  const int max_half = 2 * tiles_per_group;

  unsigned char *dW, *dXh;
  int *dCuts;
  float2 *dPart;
  unsigned *dCnt;
  float *dLse, *dAcc;
  CK(cudaMalloc(&dW, W.size()));
  CK(cudaMalloc(&dXh, Xh.size()));
  CK(cudaMalloc(&dCuts, cuts.size() * sizeof(int)));
  CK(cudaMalloc(&dPart, (size_t)n_slices * P * sizeof(float2)));
  CK(cudaMalloc(&dCnt, (size_t)groups * max_half * sizeof(unsigned)));
  CK(cudaMalloc(&dLse, (size_t)P * sizeof(float)));
  CK(cudaMalloc(&dAcc, (size_t)groups * n_slices * kSlice * 128 * sizeof(float)));
  CK(cudaMemcpy(dW, W.data(), W.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dXh, Xh.data(), Xh.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dCuts, cuts.data(), cuts.size() * sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(k_onepass_bw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float ms = 0.f;
  for (int rep = 0; rep < 2; rep++) {
    CK(cudaMemset(dCnt, 0, (size_t)groups * max_half * sizeof(unsigned)));
    CK(cudaEventRecord(e0));
    // all CTAs of a group must be co-resident (they wait on each other's counters): 1 CTA per SM,
    // grid <= SM count; a production version launches cooperatively
    k_onepass_bw<<<groups * n_slices, kThreads, kSmem>>>(n_slices, dW, dXh, dCuts, P, dPart, dCnt, max_half, dLse, dAcc);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaEventElapsedTime(&ms, e0, e1));
  }
  std::vector<float> lse(P), acc((size_t)groups * n_slices * kSlice * 128);
  CK(cudaMemcpy(lse.data(), dLse, lse.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(acc.data(), dAcc, acc.size() * 4, cudaMemcpyDeviceToHost));

  // ---- fp64 reference on a subset of groups (all components)
  double worst_lse = 0.0, worst_n = 0.0, worst_f = 0.0, max_n = 0.0, max_f = 0.0;
  const int check_groups = std::min(groups, 2);
  std::vector<double> S(C), Nref((size_t)C), Fref((size_t)C * D);
  for (int g = 0; g < check_groups; g++) {
    std::fill(Nref.begin(), Nref.end(), 0.0);
    std::fill(Fref.begin(), Fref.end(), 0.0);
    for (long p = (long)cuts[g] * kTile; p < (long)cuts[g + 1] * kTile; p++) {
      double m = -1e300;
      for (int c = 0; c < C; c++) {
        double a = lw[c];
        for (int i = 0; i < D; i++) {
          const double v = var[(size_t)c * D + i], dlt = (double)X[(size_t)p * D + i] - mean[(size_t)c * D + i];
          a += -0.5 * std::log2(6.283185307179586 * v) - kHalfLog2e * dlt * dlt / v;
        }
        S[c] = a;
        m = std::max(m, a);
      }
      double z = 0.0;
      for (int c = 0; c < C; c++) z += std::exp2(S[c] - m);
      const double l = m + std::log2(z);
      worst_lse = std::max(worst_lse, std::fabs(l - (double)lse[p]));
      for (int c = 0; c < C; c++) {
        const double gam = std::exp2(S[c] - l);
        if (gam < 1e-12) continue;
        Nref[c] += gam;
        for (int i = 0; i < D; i++) Fref[(size_t)c * D + i] += gam * (double)X[(size_t)p * D + i];
      }
    }
    for (int c = 0; c < C; c++) {
      const float *a = acc.data() + (((size_t)g * n_slices + c / kSlice) * kSlice + c % kSlice) * 128;
      const double n = (double)a[kOneCol] / 16384.0;
      worst_n = std::max(worst_n, std::fabs(n - Nref[c]));
      max_n = std::max(max_n, Nref[c]);
      for (int i = 0; i < D; i++) {
        const double f = ((double)a[i] + (double)a[64 + i]) / 16384.0;
        worst_f = std::max(worst_f, std::fabs(f - Fref[(size_t)c * D + i]));
        max_f = std::max(max_f, std::fabs(Fref[(size_t)c * D + i]));
      }
    }
  }
  std::printf("max |lse error| %.3e log2 units; N: max |err| %.3e of max %.3e; F: max |err| %.3e of max %.3e\n", worst_lse,
              worst_n, max_n, worst_f, max_f);
  std::printf("kernel %.3f ms for %ld frames = %.1f M frames/s (two-pass library kernels: ~420 M frames/s for BW)\n", ms,
              P, P / (ms * 1e-3) / 1e6);
  const bool ok = worst_lse < 1e-3 && worst_n < 2e-3 * std::max(max_n, 1.0) && worst_f < 2e-3 * std::max(max_f, 1.0);
  std::printf(ok ? "PASS\n" : "FAIL\n");
  return ok ? 0 : 3;
}
