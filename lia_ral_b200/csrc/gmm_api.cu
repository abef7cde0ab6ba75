// gmm_api.cu -- extern "C" entry points of the frames x components path: they own the frame
// and segment loops that LIA_SpkTools runs one frame at a time (accumulateStatEM,
// computeAndAccumulateTVStat, the ComputeTest frame loop), staging host frames through two
// device buffers so the PCIe copy of block k+1 overlaps the kernels of block k.
#include <algorithm>
#include <cstdlib>

#include "gmm_topk.cuh"

namespace lr {

namespace {

constexpr long kBlockFrames = 1L << 18;  // frames per staged block (63 MB at D = 60)

// positions [pos, pos + len) of the frame list map to the frames begin, begin + 1, ...; positions
// covered by no span are padding.  The per-position index is expanded ON THE DEVICE from the spans
// (k_expand_index): a host loop over every frame cost more than the kernels it feeds.
struct Span {
  long long pos, begin, len;
};

struct Plan {
  std::vector<Span> spans;  // empty -> identity
  std::vector<LrChunk> chunks;
  long P = 0;
};

__global__ void k_expand_index(const Span *__restrict__ spans, int n, long P,
                               unsigned *__restrict__ index) {
  long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  int lo = 0, hi = n - 1;  // last span starting at or before p
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (spans[mid].pos <= p)
      lo = mid;
    else
      hi = mid - 1;
  }
  const long long off = p - spans[lo].pos;
  index[p] = (off >= 0 && off < spans[lo].len) ? (unsigned)(spans[lo].begin + off) : kPadFrame;
}

struct Clip {
  int row;
  long begin;  // relative to the block base
  long len;
};

// Frame list + chunk list of the frames [b0, b1) (absolute), positions relative to `base`.
// segs == nullptr selects every frame, row 0.  Chunks never straddle rows.
// pad = true (tensor-core path): every row's run of positions is padded with kPadFrame entries
// to a whole number of 128-frame tiles, so chunks are tile aligned.
void build_plan(const lr_seg *segs, size_t n_segs, long b0, long b1, long base, bool use_rows,
                bool pad, Plan &plan) {
  plan.spans.clear();
  plan.chunks.clear();
  plan.P = 0;
  if (!segs) {
    plan.P = b1 - b0;
    if (b0 != base) plan.spans.push_back({0, b0 - base, plan.P});
    for (long p = 0; p < plan.P; p += kChunkFrames)
      plan.chunks.push_back({p, (int)std::min<long>(kChunkFrames, plan.P - p), 0});
    return;
  }
  std::vector<Clip> clips;
  for (size_t s = 0; s < n_segs; s++) {
    long sb = std::max<long>(segs[s].begin, b0), se = std::min<long>(segs[s].begin + segs[s].length, b1);
    if (se > sb) clips.push_back({use_rows ? segs[s].row : 0, sb - base, se - sb});
  }
  std::stable_sort(clips.begin(), clips.end(),
                   [](const Clip &a, const Clip &b) { return a.row < b.row; });
  plan.spans.reserve(clips.size());
  long pos = 0;
  size_t i = 0;
  while (i < clips.size()) {
    int row = clips[i].row;
    long row_start = pos;
    while (i < clips.size() && clips[i].row == row) {
      plan.spans.push_back({pos, clips[i].begin, clips[i].len});
      pos += clips[i].len;
      i++;
    }
    for (long p = row_start; p < pos; p += kChunkFrames)
      plan.chunks.push_back({p, (int)std::min<long>(kChunkFrames, pos - p), row});
    if (pad) pos = (pos + 127) / 128 * 128;  // the gap up to the next tile is padding
  }
  plan.P = pos;
  if (plan.spans.empty() && plan.P > 0) plan.spans.push_back({0, 0, 0});  // all padding
}

lr_status check_segs(const lr_seg *segs, size_t n_segs, size_t T, size_t U, bool use_rows) {
  for (size_t s = 0; s < n_segs; s++) {
    LR_REQUIRE(segs[s].begin >= 0 && segs[s].length >= 0 &&
                   (size_t)(segs[s].begin + segs[s].length) <= T,
               "segment %zu [%lld, +%lld) outside the %zu frames", s, (long long)segs[s].begin,
               (long long)segs[s].length, T);
    LR_REQUIRE(!use_rows || (segs[s].row >= 0 && (size_t)segs[s].row < U),
               "segment %zu: row %d outside [0, %zu)", s, segs[s].row, U);
  }
  return LR_OK;
}

// Upload the plan and run pass 1 (+ pass 2 when any output is requested) over it.
lr_status run_plan(lr_gmm *g, const float *dX, size_t ldx, const Plan &plan, bool tc, double fw,
                   double *dN, double *dF, double *dS2, double *d_llk_sum, unsigned char *conv = nullptr,
                   bool conv_valid = false) {
  if (plan.P == 0) return LR_OK;
  Engine &e = engine();
  unsigned *d_index = nullptr;
  if (!plan.spans.empty()) {
    d_index = (unsigned *)scratch_get(kSlotIndex, (size_t)plan.P * sizeof(unsigned));
    Span *d_spans = (Span *)scratch_get(kSlotSpans, plan.spans.size() * sizeof(Span));
    if (!d_index || !d_spans) return LR_ERR_CUDA;
    LR_CUDA(cudaMemcpyAsync(d_spans, plan.spans.data(), plan.spans.size() * sizeof(Span),
                            cudaMemcpyHostToDevice, e.stream));
    k_expand_index<<<(unsigned)((plan.P + 255) / 256), 256, 0, e.stream>>>(d_spans, (int)plan.spans.size(),
                                                                          plan.P, d_index);
    LR_CHECK_LAUNCH();
  }
  FrameList fl{dX, ldx, d_index, plan.P};
  if (tc) return tc_run_stats(g, fl, plan.chunks, fw, dN, dF, dS2, d_llk_sum, conv, conv_valid);
  float *d_lse = (float *)scratch_get(kSlotLse, (plan.P + 128) * sizeof(float));
  if (!d_lse) return LR_ERR_CUDA;
  lr_status st = gmm_pass_lse(g, fl, d_lse, nullptr, d_llk_sum);
  if (st != LR_OK) return st;
  if (dN || dF || dS2) {
    LrChunk *d_chunks = (LrChunk *)scratch_get(kSlotChunks, plan.chunks.size() * sizeof(LrChunk));
    if (!d_chunks) return LR_ERR_CUDA;
    LR_CUDA(cudaMemcpyAsync(d_chunks, plan.chunks.data(), plan.chunks.size() * sizeof(LrChunk),
                            cudaMemcpyHostToDevice, e.stream));
    st = gmm_pass_acc(g, fl, d_lse, d_chunks, (int)plan.chunks.size(), fw, dN, dF, dS2);
  }
  return st;
}

// Device-resident frames are processed in blocks of at most 2^22 frames (2 GB of fp16 hi/lo
// operand + 0.5 GB of per-slice partials); the range is cut into EQUAL blocks so that no launch is
// a short tail (multiple of 128 frames: tile aligned for the identity frame list).
long dev_block_step(long lo, long hi) {
  const long kMax = 1L << 22;
  const long total = std::max<long>(hi - lo, 1);
  const long n_blocks = (total + kMax - 1) / kMax;
  const long step = (total + n_blocks - 1) / n_blocks;
  return (step + 127) / 128 * 128;
}

// Frame range touched by the segments (or [0, T) without segments).
void seg_range(const lr_seg *segs, size_t n_segs, size_t T, long &lo, long &hi) {
  if (!segs) {
    lo = 0;
    hi = (long)T;
    return;
  }
  lo = (long)T;
  hi = 0;
  for (size_t s = 0; s < n_segs; s++) {
    if (segs[s].length <= 0) continue;
    lo = std::min<long>(lo, segs[s].begin);
    hi = std::max<long>(hi, segs[s].begin + segs[s].length);
  }
  if (hi < lo) lo = hi = 0;
}

// Host frames -> staged blocks -> fn(dX_block, b0, b1).  Copies run on the copy stream and
// alternate between two device buffers guarded by events.
template <typename Fn>
lr_status for_each_host_block(const float *X, size_t ldx, long lo, long hi, Fn fn) {
  Engine &e = engine();
  int k = 0;
  for (long b0 = lo; b0 < hi; b0 += kBlockFrames, k++) {
    long b1 = std::min(hi, b0 + kBlockFrames);
    int s = k & 1;
    size_t bytes = (size_t)(b1 - b0) * ldx * sizeof(float);
    float *buf = (float *)scratch_get(s ? kSlotX1 : kSlotX0,
                                      (size_t)std::min<long>(kBlockFrames, hi - lo) * ldx *
                                          sizeof(float));
    if (!buf) return LR_ERR_CUDA;
    if (k >= 2) LR_CUDA(cudaStreamWaitEvent(e.copy_stream, e.ev_consumed[s], 0));
    LR_CUDA(cudaMemcpyAsync(buf, X + (size_t)b0 * ldx, bytes, cudaMemcpyHostToDevice,
                            e.copy_stream));
    LR_CUDA(cudaEventRecord(e.ev_copied[s], e.copy_stream));
    LR_CUDA(cudaStreamWaitEvent(e.stream, e.ev_copied[s], 0));
    lr_status st = fn(buf, b0, b1);
    if (st != LR_OK) return st;
    LR_CUDA(cudaEventRecord(e.ev_consumed[s], e.stream));
  }
  return LR_OK;
}

__global__ void k_add_scalar(double *dst, double v) { *dst += v; }

// session rows -> += session accumulator, += speaker accumulator (several sessions per speaker: RED adds)
__global__ void k_jfa_fold(size_t n_sessions, size_t width, const int *__restrict__ spk,
                           const double *__restrict__ t, double *__restrict__ acc_h, double *__restrict__ acc) {
  const size_t total = n_sessions * width;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t h = i / width, k = i - h * width;
    const double v = t[i];
    acc_h[i] += v;
    if (v != 0.0) atomicAdd(acc + (size_t)spk[h] * width + k, v);
  }
}

}  // namespace
}  // namespace lr

using namespace lr;

extern "C" {

lr_status lr_gmm_em_accumulate(lr_gmm *g, const float *X, size_t T, size_t ldx,
                               const lr_seg *segs, size_t n_segs, double frame_weight,
                               double *occ, double *m1, double *m2, double *sum_log_lk,
                               double *n_frames) {
  LR_READY();
  LR_REQUIRE(g && X && occ && m1 && m2, "lr_gmm_em_accumulate: null argument");
  LR_REQUIRE(ldx >= (size_t)g->D && T < (1ull << 32), "lr_gmm_em_accumulate: bad T / ldx");
  if (segs) {
    lr_status st = check_segs(segs, n_segs, T, 1, false);
    if (st != LR_OK) return st;
  }
  Engine &e = engine();
  size_t C = g->C, cd = (size_t)g->C * g->D, n = lr_gmm_em_stats_len(g);
  double *d_stats = (double *)scratch_get(kSlotStats, n * sizeof(double));
  if (!d_stats) return LR_ERR_CUDA;
  LR_CUDA(cudaMemsetAsync(d_stats, 0, n * sizeof(double), e.stream));
  long lo, hi;
  seg_range(segs, n_segs, T, lo, hi);
  Plan plan;
  double total_pos = 0.0;
  lr_status sel = LR_OK;
  const bool tc = tc_selected(g, &sel);
  if (sel != LR_OK) return sel;
  lr_status st = for_each_host_block(X, ldx, lo, hi, [&](const float *dX, long b0, long b1) {
    build_plan(segs, n_segs, b0, b1, b0, false, tc, plan);
    for (auto &c : plan.chunks) total_pos += (double)c.len;
    return run_plan(g, dX, ldx, plan, tc, frame_weight, d_stats, d_stats + C, d_stats + C + cd,
                    d_stats + C + 2 * cd);
  });
  if (st != LR_OK) return st;
  std::vector<double> h(n);
  LR_CUDA(cudaMemcpyAsync(h.data(), d_stats, n * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  LR_CUDA(cudaStreamSynchronize(e.stream));
  for (size_t i = 0; i < C; i++) occ[i] += h[i];
  for (size_t i = 0; i < cd; i++) {
    m1[i] += h[C + i];
    m2[i] += h[C + cd + i];
  }
  if (sum_log_lk) *sum_log_lk += h[C + 2 * cd];
  if (n_frames) *n_frames += frame_weight * total_pos;
  return LR_OK;
}

lr_status lr_gmm_em_accumulate_dev(lr_gmm *g, const lr_feats *f, size_t t0, size_t T,
                                   double frame_weight, double *d_stats) {
  LR_READY();
  LR_REQUIRE(g && f && d_stats, "lr_gmm_em_accumulate_dev: null argument");
  LR_REQUIRE(f->D == g->D && t0 + T <= f->T, "lr_gmm_em_accumulate_dev: frame range / vectSize");
  size_t C = g->C, cd = (size_t)g->C * g->D;
  Plan plan;
  lr_status sel = LR_OK;
  const bool tc = tc_selected(g, &sel);
  if (sel != LR_OK) return sel;
  const long step = dev_block_step((long)t0, (long)(t0 + T));
  // The converted frame operand (512 B per frame) is kept with the handle between calls over the same
  // range: EM iterations re-read the same frames, and tc_derive keeps the normalised space stable.
  // Budget: LR_CONV_CACHE_GB (default 24) -- beyond it, or if the allocation fails, every call converts.
  bool cache = false, cache_valid = false;
  if (tc && engine().gmm_kernel != 3) {
    size_t tiles = 0;
    for (long b0 = (long)t0; b0 < (long)(t0 + T); b0 += step)
      tiles += (size_t)((std::min<long>((long)(t0 + T), b0 + step) - b0 + 127) / 128);
    const size_t need = tiles * kTcTileBytesPub;
    static const double budget_gb = getenv("LR_CONV_CACHE_GB") ? atof(getenv("LR_CONV_CACHE_GB")) : 24.0;
    if ((double)need <= budget_gb * 1e9) {
      if (f->conv_cap < need) {
        cudaFree(f->d_conv);
        f->d_conv = nullptr;
        f->conv_cap = 0;
        f->conv_norm = 0;
        if (cudaMalloc(&f->d_conv, need) == cudaSuccess) f->conv_cap = need;
        else cudaGetLastError();
      }
      cache = f->d_conv != nullptr;
      const unsigned long long norm = tc_norm_id(g);
      cache_valid = cache && norm != 0 && f->conv_norm == norm && f->conv_t0 == t0 && f->conv_T == T;
      if (cache && !cache_valid) {
        f->conv_norm = norm;
        f->conv_t0 = t0;
        f->conv_T = T;
      }
    }
  }
  size_t tile_off = 0;
  for (long b0 = (long)t0; b0 < (long)(t0 + T); b0 += step) {
    long b1 = std::min<long>((long)(t0 + T), b0 + step);
    build_plan(nullptr, 0, b0, b1, b0, false, tc, plan);
    lr_status st = run_plan(g, f->d_x + (size_t)b0 * f->ldx, f->ldx, plan, tc, frame_weight, d_stats,
                            d_stats + C, d_stats + C + cd, d_stats + C + 2 * cd,
                            cache ? f->d_conv + tile_off * kTcTileBytesPub : nullptr, cache_valid);
    if (st != LR_OK) {
      f->conv_norm = 0;
      return st;
    }
    tile_off += (size_t)((b1 - b0 + 127) / 128);
  }
  // n_frames accumulates on the device too (no host sync in this variant)
  k_add_scalar<<<1, 1, 0, engine().stream>>>(d_stats + C + 2 * cd + 1, frame_weight * (double)T);
  LR_CHECK_LAUNCH();
  return LR_OK;
}

lr_status lr_gmm_bwstats(lr_gmm *g, const float *X, size_t T, size_t ldx, const lr_seg *segs,
                         size_t n_segs, size_t U, double *N, double *F) {
  LR_READY();
  LR_REQUIRE(g && X && N && F && U > 0, "lr_gmm_bwstats: null argument");
  LR_REQUIRE(ldx >= (size_t)g->D && T < (1ull << 32), "lr_gmm_bwstats: bad T / ldx");
  if (segs) {
    lr_status st = check_segs(segs, n_segs, T, U, true);
    if (st != LR_OK) return st;
  }
  Engine &e = engine();
  size_t nN = U * (size_t)g->C, nF = nN * g->D;
  DevBuf<double> dN, dF;
  LR_CUDA(dN.alloc(nN));
  LR_CUDA(dF.alloc(nF));
  LR_CUDA(cudaMemcpyAsync(dN.p, N, nN * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  LR_CUDA(cudaMemcpyAsync(dF.p, F, nF * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  long lo, hi;
  seg_range(segs, n_segs, T, lo, hi);
  Plan plan;
  lr_status sel = LR_OK;
  const bool tc = tc_selected(g, &sel);
  if (sel != LR_OK) return sel;
  lr_status st = for_each_host_block(X, ldx, lo, hi, [&](const float *dX, long b0, long b1) {
    build_plan(segs, n_segs, b0, b1, b0, true, tc, plan);
    return run_plan(g, dX, ldx, plan, tc, 1.0, dN.p, dF.p, nullptr, nullptr);
  });
  if (st != LR_OK) return st;
  LR_CUDA(cudaMemcpyAsync(N, dN.p, nN * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  LR_CUDA(cudaMemcpyAsync(F, dF.p, nF * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  LR_CUDA(cudaStreamSynchronize(e.stream));
  return LR_OK;
}

lr_status lr_gmm_bwstats_dev(lr_gmm *g, const lr_feats *f, const lr_seg *segs, size_t n_segs,
                             size_t U, double *d_N, double *d_F) {
  LR_READY();
  LR_REQUIRE(g && f && d_N && d_F && U > 0, "lr_gmm_bwstats_dev: null argument");
  LR_REQUIRE(f->D == g->D && f->T < (1ull << 32), "lr_gmm_bwstats_dev: vectSize / frame count");
  if (segs) {
    lr_status st = check_segs(segs, n_segs, f->T, U, true);
    if (st != LR_OK) return st;
  }
  long lo, hi;
  seg_range(segs, n_segs, f->T, lo, hi);
  Plan plan;
  lr_status sel = LR_OK;
  const bool tc = tc_selected(g, &sel);
  if (sel != LR_OK) return sel;
  const long step = dev_block_step(lo, hi);
  for (long b0 = lo; b0 < hi; b0 += step) {
    long b1 = std::min(hi, b0 + step);
    build_plan(segs, n_segs, b0, b1, 0, true, tc, plan);
    lr_status st = run_plan(g, f->d_x, f->ldx, plan, tc, 1.0, d_N, d_F, nullptr, nullptr);
    if (st != LR_OK) return st;
  }
  return LR_OK;
}

// JFAAcc::computeAndAccumulateJFAStat (AccumulateJFAStat.cpp:520-576): the same posteriors feed a
// per-SESSION accumulator (N_h, F_X_h) and a per-SPEAKER one (N, F_X).  The frames x components kernel
// runs once with one statistics row per session; the speaker rows are folded from the session rows.
lr_status lr_jfa_bwstats(lr_gmm *g, const float *X, size_t T, size_t ldx, const lr_seg *segs, size_t n_segs,
                         size_t n_sessions, const int32_t *speaker_of_session, size_t n_speakers, double *N_h,
                         double *F_h, double *N, double *F) {
  LR_READY();
  LR_REQUIRE(g && X && segs && speaker_of_session && N_h && F_h && N && F && n_sessions > 0 && n_speakers > 0,
             "lr_jfa_bwstats: null argument");
  LR_REQUIRE(ldx >= (size_t)g->D && T < (1ull << 32), "lr_jfa_bwstats: bad T / ldx");
  for (size_t i = 0; i < n_sessions; i++)
    LR_REQUIRE(speaker_of_session[i] >= 0 && (size_t)speaker_of_session[i] < n_speakers,
               "lr_jfa_bwstats: session %zu belongs to speaker %d outside [0, %zu)", i, speaker_of_session[i],
               n_speakers);
  lr_status st = check_segs(segs, n_segs, T, n_sessions, true);
  if (st != LR_OK) return st;
  Engine &e = engine();
  const size_t C = g->C, sv = C * g->D;
  DevBuf<double> tN, tF, dNh, dFh, dN, dF;
  DevBuf<int> dSpk;
  LR_CUDA(tN.alloc(n_sessions * C));
  LR_CUDA(tF.alloc(n_sessions * sv));
  LR_CUDA(dNh.alloc(n_sessions * C));
  LR_CUDA(dFh.alloc(n_sessions * sv));
  LR_CUDA(dN.alloc(n_speakers * C));
  LR_CUDA(dF.alloc(n_speakers * sv));
  LR_CUDA(dSpk.alloc(n_sessions));
  LR_CUDA(cudaMemsetAsync(tN.p, 0, n_sessions * C * sizeof(double), e.stream));
  LR_CUDA(cudaMemsetAsync(tF.p, 0, n_sessions * sv * sizeof(double), e.stream));
  LR_CUDA(cudaMemcpyAsync(dNh.p, N_h, n_sessions * C * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  LR_CUDA(cudaMemcpyAsync(dFh.p, F_h, n_sessions * sv * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  LR_CUDA(cudaMemcpyAsync(dN.p, N, n_speakers * C * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  LR_CUDA(cudaMemcpyAsync(dF.p, F, n_speakers * sv * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  LR_CUDA(cudaMemcpyAsync(dSpk.p, speaker_of_session, n_sessions * sizeof(int), cudaMemcpyHostToDevice, e.stream));
  long lo, hi;
  seg_range(segs, n_segs, T, lo, hi);
  Plan plan;
  lr_status sel = LR_OK;
  const bool tc = tc_selected(g, &sel);
  if (sel != LR_OK) return sel;
  st = for_each_host_block(X, ldx, lo, hi, [&](const float *dX, long b0, long b1) {
    build_plan(segs, n_segs, b0, b1, b0, true, tc, plan);
    return run_plan(g, dX, ldx, plan, tc, 1.0, tN.p, tF.p, nullptr, nullptr);
  });
  if (st != LR_OK) return st;
  k_jfa_fold<<<(unsigned)std::min<size_t>(4096, (n_sessions * C + 255) / 256), 256, 0, e.stream>>>(
      n_sessions, C, dSpk.p, tN.p, dNh.p, dN.p);
  LR_CHECK_LAUNCH();
  k_jfa_fold<<<(unsigned)std::min<size_t>(65535, (n_sessions * sv + 255) / 256), 256, 0, e.stream>>>(
      n_sessions, sv, dSpk.p, tF.p, dFh.p, dF.p);
  LR_CHECK_LAUNCH();
  LR_CUDA(cudaMemcpyAsync(N_h, dNh.p, n_sessions * C * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  LR_CUDA(cudaMemcpyAsync(F_h, dFh.p, n_sessions * sv * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  LR_CUDA(cudaMemcpyAsync(N, dN.p, n_speakers * C * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  LR_CUDA(cudaMemcpyAsync(F, dF.p, n_speakers * sv * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  LR_CUDA(cudaStreamSynchronize(e.stream));
  return LR_OK;
}

// -------------------------------------------------------------------------- JFA feature compensation
// JFAAcc::normalizeFeatures (AccumulateJFAStat.cpp:4623-4680): every selected frame loses the posterior-
// weighted channel offset,  x_t -= sum_k P(k | x_t) (U x)_k,  the posteriors taken under the session model
// (means M + U x, the world's weights and variances).  The frames x components scores come from the same
// likelihood pass that feeds the top-K path (tcgen05 for a tensor-core-served model); this kernel is the
// contraction of the [P x C] posteriors with the [C x D] offset (k_jfa_compensate below).
constexpr int kJfaChunk = 64, kJfaWarps = 8, kJfaFr = 4;  // components per chunk, warps per block, frames per warp
constexpr long kJfaBlock = 1L << 15;  // frames per likelihood pass (S is P x Cp floats)

// One warp per FOUR frames, the lanes over the dimensions (d = lane, lane + 32): per chunk of 64 components the
// block stages the offset rows, every warp turns its frames' scores into posteriors once (two exp2 per lane and
// frame, stored [c][frame] so one 16-byte broadcast load fetches the four), then each component costs three
// shared-memory loads for eight FMAs and nothing has to be reduced across lanes at the end.
__global__ void __launch_bounds__(kJfaWarps * 32)
k_jfa_compensate(int C, int Cp, int D, const float *__restrict__ S, const float *__restrict__ lse2,
                 const float *__restrict__ ux /*[Cp][64], zero padded*/, const unsigned *__restrict__ index,
                 long P, float *__restrict__ X, size_t ldx) {
  __shared__ __align__(16) float su[kJfaChunk][64];
  __shared__ __align__(16) float gs[kJfaWarps][kJfaChunk][kJfaFr];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long p0 = ((long)blockIdx.x * kJfaWarps + warp) * kJfaFr;
  float l2[kJfaFr];
#pragma unroll
  for (int f = 0; f < kJfaFr; f++) l2[f] = p0 + f < P ? lse2[p0 + f] : 0.f;
  float acc[kJfaFr][2];
#pragma unroll
  for (int f = 0; f < kJfaFr; f++) acc[f][0] = acc[f][1] = 0.f;
  for (int c0 = 0; c0 < C; c0 += kJfaChunk) {
    __syncthreads();  // the previous chunk has been consumed
    for (int i = threadIdx.x; i < kJfaChunk * 64; i += kJfaWarps * 32) {
      const int c = i >> 6, d = i & 63;
      su[c][d] = (c0 + c < C) ? ux[(size_t)(c0 + c) * 64 + d] : 0.f;
    }
#pragma unroll
    for (int f = 0; f < kJfaFr; f++) {
      const bool live = p0 + f < P;
      const float *Sp = S + (size_t)(live ? p0 + f : 0) * Cp + c0;
#pragma unroll
      for (int h = 0; h < kJfaChunk / 32; h++) {
        const int c = 32 * h + lane;
        gs[warp][c][f] = (live && c0 + c < C) ? exp2f(Sp[c] - l2[f]) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int c = 0; c < kJfaChunk; c++) {
      const float4 g4 = *reinterpret_cast<const float4 *>(&gs[warp][c][0]);
      const float u0 = su[c][lane], u1 = su[c][lane + 32];
      acc[0][0] = fmaf(g4.x, u0, acc[0][0]);
      acc[0][1] = fmaf(g4.x, u1, acc[0][1]);
      acc[1][0] = fmaf(g4.y, u0, acc[1][0]);
      acc[1][1] = fmaf(g4.y, u1, acc[1][1]);
      acc[2][0] = fmaf(g4.z, u0, acc[2][0]);
      acc[2][1] = fmaf(g4.z, u1, acc[2][1]);
      acc[3][0] = fmaf(g4.w, u0, acc[3][0]);
      acc[3][1] = fmaf(g4.w, u1, acc[3][1]);
    }
  }
#pragma unroll
  for (int f = 0; f < kJfaFr; f++) {
    if (p0 + f >= P) continue;
    const size_t fr = index ? index[p0 + f] : (size_t)(p0 + f);
    float *x = X + fr * ldx;
    if (lane < D) x[lane] -= acc[f][0];
    if (lane + 32 < D) x[lane + 32] -= acc[f][1];
  }
}

lr_status lr_jfa_normalize_features(lr_gmm *session_model, const double *ux, float *X, size_t T, size_t ldx,
                                    const lr_seg *segs, size_t n_segs) {
  LR_READY();
  LR_REQUIRE(session_model && ux && X && segs, "lr_jfa_normalize_features: null argument");
  lr_gmm *g = session_model;
  LR_REQUIRE(ldx >= (size_t)g->D && T < (1ull << 32) && g->D <= 64, "lr_jfa_normalize_features: bad T / ldx / D");
  lr_status st = check_segs(segs, n_segs, T, 1, false);
  if (st != LR_OK) return st;
  long lo, hi;
  seg_range(segs, n_segs, T, lo, hi);
  if (hi <= lo) return LR_OK;
  Engine &e = engine();
  // the frames are visited segment by segment like the reference's loop: a frame covered by several
  // segments is compensated once per occurrence, each time from its CURRENT value.  Layer r holds the
  // r-th occurrence of every frame; within a layer the frames are distinct (in-place update, no race)
  std::vector<std::vector<unsigned>> layers;
  {
    std::vector<unsigned char> seen((size_t)(hi - lo), 0);
    for (size_t s = 0; s < n_segs; s++)
      for (long t = segs[s].begin; t < segs[s].begin + segs[s].length; t++) {
        const unsigned r = seen[(size_t)(t - lo)]++;
        LR_REQUIRE(r < 255, "lr_jfa_normalize_features: frame %ld is covered by more than 255 segments", t);
        if (layers.size() <= r) layers.emplace_back();
        layers[r].push_back((unsigned)(t - lo));
      }
  }
  const int C = g->C, Cp = g->Cp, D = g->D;
  DevBuf<float> dX, dU;
  DevBuf<unsigned> dIdx;
  LR_CUDA(dX.alloc((size_t)(hi - lo) * ldx));
  LR_CUDA(dU.alloc((size_t)Cp * 64));
  {
    std::vector<float> uf((size_t)Cp * 64, 0.f);
    for (int c = 0; c < C; c++)
      for (int d = 0; d < D; d++) uf[(size_t)c * 64 + d] = (float)ux[(size_t)c * D + d];
    LR_CUDA(cudaMemcpyAsync(dU.p, uf.data(), uf.size() * sizeof(float), cudaMemcpyHostToDevice, e.stream));
    LR_CUDA(cudaMemcpyAsync(dX.p, X + (size_t)lo * ldx, (size_t)(hi - lo) * ldx * sizeof(float),
                            cudaMemcpyHostToDevice, e.stream));
    LR_CUDA(cudaStreamSynchronize(e.stream));  // uf leaves scope
  }
  lr_status sel = LR_OK;
  const bool tc = tc_selected(g, &sel);
  if (sel != LR_OK) return sel;
  for (const auto &layer : layers) {
    LR_CUDA(dIdx.alloc(layer.size()));
    LR_CUDA(cudaMemcpyAsync(dIdx.p, layer.data(), layer.size() * sizeof(unsigned), cudaMemcpyHostToDevice, e.stream));
    for (long b0 = 0; b0 < (long)layer.size(); b0 += kJfaBlock) {
      const long P = std::min<long>(kJfaBlock, (long)layer.size() - b0);
      float *d_lse = (float *)scratch_get(kSlotLse, (P + 128) * sizeof(float));
      float *d_S = (float *)scratch_get(kSlotS, (size_t)P * Cp * sizeof(float));
      if (!d_lse || !d_S) return LR_ERR_CUDA;
      FrameList fl{dX.p, ldx, dIdx.p + b0, P};
      st = tc ? tc_pass_lse(g, fl, d_lse, nullptr, d_S) : gmm_pass_lse(g, fl, d_lse, d_S, nullptr);
      if (st != LR_OK) return st;
      k_jfa_compensate<<<(unsigned)((P + kJfaWarps * kJfaFr - 1) / (kJfaWarps * kJfaFr)), kJfaWarps * 32, 0, e.stream>>>(
          C, Cp, D, d_S, d_lse, dU.p, dIdx.p + b0, P, dX.p, ldx);
      LR_CHECK_LAUNCH();
    }
    LR_CUDA(cudaStreamSynchronize(e.stream));  // layer.data() / dIdx are reused
  }
  LR_CUDA(cudaMemcpyAsync(X + (size_t)lo * ldx, dX.p, (size_t)(hi - lo) * ldx * sizeof(float),
                          cudaMemcpyDeviceToHost, e.stream));
  LR_CUDA(cudaStreamSynchronize(e.stream));
  return LR_OK;
}

// -------------------------------------------------------------------------- LLK / top-K
lr_status lr_gmm_llk(lr_gmm *g, const float *X, size_t T, size_t ldx, double min_llk,
                     double max_llk, double *llk) {
  LR_READY();
  LR_REQUIRE(g && X && llk && ldx >= (size_t)g->D, "lr_gmm_llk: bad argument");
  Engine &e = engine();
  lr_status sel = LR_OK;
  const bool tc = tc_selected(g, &sel);
  if (sel != LR_OK) return sel;
  return for_each_host_block(X, ldx, 0, (long)T, [&](const float *dX, long b0, long b1) {
    long P = b1 - b0;
    float *d_lse = (float *)scratch_get(kSlotLse, (P + 128) * sizeof(float));
    double *d_llk = (double *)scratch_get(kSlotLlk, P * sizeof(double));
    if (!d_lse || !d_llk) return (lr_status)LR_ERR_CUDA;
    FrameList fl{dX, ldx, nullptr, P};
    lr_status st = tc ? tc_pass_lse(g, fl, d_lse, nullptr) : gmm_pass_lse(g, fl, d_lse, nullptr, nullptr);
    if (st != LR_OK) return st;
    st = gmm_llk_from_lse(P, d_lse, min_llk, max_llk, d_llk);
    if (st != LR_OK) return st;
    LR_CUDA(cudaMemcpyAsync(llk + b0, d_llk, P * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
    LR_CUDA(cudaStreamSynchronize(e.stream));
    return (lr_status)LR_OK;
  });
}

// world top-K over a device block; outputs stay on the device in scratch slots
// worldDecime (ComputeTest.cpp:111-113, 162): frame t of a block that starts on the grid
__global__ void k_decime_propagate(long P, int K, int decime, unsigned *__restrict__ idx, double *__restrict__ rest) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P) return;
  const long off = t % decime;
  if (off == 0) return;
  const long src = t - off;
  for (int k = 0; k < K; k++) idx[t * K + k] = idx[src * K + k];
  rest[t] = rest[src];
}
__global__ void k_decime_select(long P, int decime, const double *__restrict__ use, double *__restrict__ llk) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < P && t % decime != 0) llk[t] = use[t];
}

static lr_status topk_block(lr_gmm *world, const float *dX, size_t ldx, long P, int K, int complete,
                            double min_llk, double max_llk, double **d_llk, unsigned **d_idx,
                            double **d_top, double **d_rest, double **d_restw) {
  float *d_lse = (float *)scratch_get(kSlotLse, (P + 128) * sizeof(float));
  float *d_S = (float *)scratch_get(kSlotS, (size_t)P * world->Cp * sizeof(float));
  *d_llk = (double *)scratch_get(kSlotLlk, P * sizeof(double));
  *d_idx = (unsigned *)scratch_get(kSlotIdx, (size_t)P * K * sizeof(unsigned));
  *d_top = (double *)scratch_get(kSlotTmpA, (size_t)P * K * sizeof(double));
  double *rest = (double *)scratch_get(kSlotRest, (size_t)P * 2 * sizeof(double));
  if (!d_lse || !d_S || !*d_llk || !*d_idx || !*d_top || !rest) return LR_ERR_CUDA;
  *d_rest = rest;
  *d_restw = rest + P;
  FrameList fl{dX, ldx, nullptr, P};
  // candidate nomination needs the full S matrix: from the tcgen05 likelihood pass when the world model is
  // served by it (a 2048-component world: the same contraction as a1), else from the fp32 SIMT pass.  Either
  // way the candidates are re-evaluated in fp64 in the reference's operation order (gmm_topk.cu).
  lr_status sel = LR_OK;
  const bool tc = tc_selected(world, &sel);
  if (sel != LR_OK) return sel;
  lr_status st = tc ? tc_pass_lse(world, fl, d_lse, nullptr, d_S) : gmm_pass_lse(world, fl, d_lse, d_S, nullptr);
  if (st != LR_OK) return st;
  return gmm_topk(world, fl, d_S, K, complete, min_llk, max_llk, *d_llk, *d_idx, *d_top, *d_rest,
                  *d_restw);
}

constexpr long kTopkBlock = 1L << 15;  // frames per block on the top-K path (S is P x Cp floats)

lr_status lr_gmm_llk_topk(lr_gmm *world, const float *X, size_t T, size_t ldx, int K,
                          int complete, double min_llk, double max_llk, double *llk,
                          uint32_t *idx, double *top_lk, double *rest_lk, double *rest_w) {
  LR_READY();
  LR_REQUIRE(world && X && idx && ldx >= (size_t)world->D, "lr_gmm_llk_topk: bad argument");
  LR_REQUIRE(K >= 1 && K <= kMaxTopK && K <= world->C, "lr_gmm_llk_topk: K=%d outside [1, min(%d, C)]",
             K, kMaxTopK);
  Engine &e = engine();
  for (long b0 = 0; b0 < (long)T; b0 += kTopkBlock) {
    long b1 = std::min<long>((long)T, b0 + kTopkBlock), P = b1 - b0;
    float *dX = (float *)scratch_get(kSlotX0, (size_t)std::min<long>(kTopkBlock, (long)T) * ldx * sizeof(float));
    if (!dX) return LR_ERR_CUDA;
    LR_CUDA(cudaMemcpyAsync(dX, X + (size_t)b0 * ldx, (size_t)P * ldx * sizeof(float),
                            cudaMemcpyHostToDevice, e.stream));
    double *d_llk, *d_top, *d_rest, *d_restw;
    unsigned *d_idx;
    lr_status st = topk_block(world, dX, ldx, P, K, complete, min_llk, max_llk, &d_llk, &d_idx,
                              &d_top, &d_rest, &d_restw);
    if (st != LR_OK) return st;
    if (llk) LR_CUDA(cudaMemcpyAsync(llk + b0, d_llk, P * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
    LR_CUDA(cudaMemcpyAsync(idx + (size_t)b0 * K, d_idx, (size_t)P * K * sizeof(unsigned),
                            cudaMemcpyDeviceToHost, e.stream));
    if (top_lk)
      LR_CUDA(cudaMemcpyAsync(top_lk + (size_t)b0 * K, d_top, (size_t)P * K * sizeof(double),
                              cudaMemcpyDeviceToHost, e.stream));
    if (rest_lk) LR_CUDA(cudaMemcpyAsync(rest_lk + b0, d_rest, P * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
    if (rest_w) LR_CUDA(cudaMemcpyAsync(rest_w + b0, d_restw, P * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
    LR_CUDA(cudaStreamSynchronize(e.stream));
  }
  return LR_OK;
}

lr_status lr_gmm_llk_use_topk(lr_gmm *client, const float *X, size_t T, size_t ldx, int K,
                              const uint32_t *idx, const double *rest_lk, int complete,
                              double min_llk, double max_llk, double *llk) {
  LR_READY();
  LR_REQUIRE(client && X && idx && llk && ldx >= (size_t)client->D, "lr_gmm_llk_use_topk: bad argument");
  LR_REQUIRE(K >= 1 && K <= client->C, "lr_gmm_llk_use_topk: bad K");
  for (size_t i = 0; i < T * (size_t)K; i++)
    LR_REQUIRE(idx[i] < (uint32_t)client->C, "lr_gmm_llk_use_topk: index %u out of range", idx[i]);
  Engine &e = engine();
  for (long b0 = 0; b0 < (long)T; b0 += kBlockFrames) {
    long b1 = std::min<long>((long)T, b0 + kBlockFrames), P = b1 - b0;
    float *dX = (float *)scratch_get(kSlotX0, (size_t)std::min<long>(kBlockFrames, (long)T) * ldx * sizeof(float));
    unsigned *d_idx = (unsigned *)scratch_get(kSlotIdx, (size_t)P * K * sizeof(unsigned));
    double *d_rest = (double *)scratch_get(kSlotRest, P * sizeof(double));
    double *d_llk = (double *)scratch_get(kSlotLlk, P * sizeof(double));
    if (!dX || !d_idx || !d_rest || !d_llk) return LR_ERR_CUDA;
    LR_CUDA(cudaMemcpyAsync(dX, X + (size_t)b0 * ldx, (size_t)P * ldx * sizeof(float),
                            cudaMemcpyHostToDevice, e.stream));
    LR_CUDA(cudaMemcpyAsync(d_idx, idx + (size_t)b0 * K, (size_t)P * K * sizeof(unsigned),
                            cudaMemcpyHostToDevice, e.stream));
    if (rest_lk)
      LR_CUDA(cudaMemcpyAsync(d_rest, rest_lk + b0, P * sizeof(double), cudaMemcpyHostToDevice, e.stream));
    FrameList fl{dX, ldx, nullptr, P};
    lr_status st = gmm_use_topk(client, fl, K, d_idx, rest_lk ? d_rest : nullptr, complete,
                                min_llk, max_llk, d_llk);
    if (st != LR_OK) return st;
    LR_CUDA(cudaMemcpyAsync(llk + b0, d_llk, P * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
    LR_CUDA(cudaStreamSynchronize(e.stream));
  }
  return LR_OK;
}

lr_status lr_compute_test(lr_gmm *world, lr_gmm *const *clients, int n_clients, const float *X,
                          size_t T, size_t ldx, const lr_seg *segs, size_t n_segs, int K,
                          int complete, double min_llk, double max_llk, int per_segment,
                          double *mean_llk_world, double *mean_llk_client) {
  return lr_compute_test_decime(world, clients, n_clients, X, T, ldx, segs, n_segs, K, complete, min_llk,
                                max_llk, per_segment, 1, mean_llk_world, mean_llk_client);
}

lr_status lr_compute_test_decime(lr_gmm *world, lr_gmm *const *clients, int n_clients, const float *X,
                                 size_t T, size_t ldx, const lr_seg *segs, size_t n_segs, int K,
                                 int complete, double min_llk, double max_llk, int per_segment,
                                 int world_decime, double *mean_llk_world, double *mean_llk_client) {
  LR_READY();
  LR_REQUIRE(world_decime >= 1 && world_decime <= 4096, "lr_compute_test: worldDecime %d outside [1, 4096]",
             world_decime);
  // blocks start on the decimation grid of their segment (idxFrame % worldDecime == 0, ComputeTest.cpp:162)
  const long block = kTopkBlock - kTopkBlock % world_decime;
  LR_REQUIRE(world && X && mean_llk_world && (n_clients == 0 || (clients && mean_llk_client)),
             "lr_compute_test: null argument");
  LR_REQUIRE(ldx >= (size_t)world->D && T > 0, "lr_compute_test: bad T / ldx");
  LR_REQUIRE(K >= 1 && K <= kMaxTopK && K <= world->C, "lr_compute_test: K=%d outside [1, min(%d, C)]",
             K, kMaxTopK);
  for (int i = 0; i < n_clients; i++)
    LR_REQUIRE(clients[i] && clients[i]->C == world->C && clients[i]->D == world->D,
               "lr_compute_test: client %d does not share the world's shape", i);
  lr_seg all = {0, (int64_t)T, 0, 0};
  if (!segs) {
    segs = &all;
    n_segs = 1;
  }
  lr_status st0 = check_segs(segs, n_segs, T, 1, false);
  if (st0 != LR_OK) return st0;
  Engine &e = engine();
  size_t n_out = per_segment ? n_segs : 1;
  std::vector<double> sum_w(n_out, 0.0), sum_c((size_t)n_clients * n_out, 0.0), cnt(n_out, 0.0);
  std::vector<double> h_w, h_c;
  // frames are scored segment by segment in blocks (frames outside the segments are never read)
  for (size_t s = 0; s < n_segs; s++) {
    size_t o = per_segment ? s : 0;
    for (long b0 = segs[s].begin; b0 < segs[s].begin + segs[s].length; b0 += block) {
      long b1 = std::min<long>(segs[s].begin + segs[s].length, b0 + block), P = b1 - b0;
      float *dX = (float *)scratch_get(kSlotX0, (size_t)kTopkBlock * ldx * sizeof(float));
      if (!dX) return LR_ERR_CUDA;
      LR_CUDA(cudaMemcpyAsync(dX, X + (size_t)b0 * ldx, (size_t)P * ldx * sizeof(float),
                              cudaMemcpyHostToDevice, e.stream));
      double *d_llk, *d_top, *d_rest, *d_restw;
      unsigned *d_idx;
      lr_status st = topk_block(world, dX, ldx, P, K, complete, min_llk, max_llk, &d_llk, &d_idx,
                                &d_top, &d_rest, &d_restw);
      if (st != LR_OK) return st;
      FrameList fl{dX, ldx, nullptr, P};
      if (world_decime > 1) {
        // frames off the grid keep the top list (and the COMPLETE rest) of the last grid frame, and the
        // world itself is scored through USE_TOP_DISTRIBS there (ComputeTest.cpp:162-165)
        k_decime_propagate<<<ceil_div(P, 256), 256, 0, e.stream>>>(P, K, world_decime, d_idx, d_rest);
        LR_CHECK_LAUNCH();
        double *d_use = (double *)scratch_get(kSlotSpans, (size_t)P * sizeof(double));
        if (!d_use) return LR_ERR_CUDA;
        st = gmm_use_topk(world, fl, K, d_idx, d_rest, complete, min_llk, max_llk, d_use);
        if (st != LR_OK) return st;
        k_decime_select<<<ceil_div(P, 256), 256, 0, e.stream>>>(P, world_decime, d_use, d_llk);
        LR_CHECK_LAUNCH();
      }
      h_w.resize(P);
      LR_CUDA(cudaMemcpyAsync(h_w.data(), d_llk, P * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
      double *d_cl = (double *)scratch_get(kSlotTmpB, (size_t)P * std::max(1, n_clients) * sizeof(double));
      if (!d_cl) return LR_ERR_CUDA;
      for (int i = 0; i < n_clients; i++) {
        st = gmm_use_topk(clients[i], fl, K, d_idx, d_rest, complete, min_llk, max_llk,
                          d_cl + (size_t)i * P);
        if (st != LR_OK) return st;
      }
      h_c.resize((size_t)P * n_clients);
      if (n_clients)
        LR_CUDA(cudaMemcpyAsync(h_c.data(), d_cl, (size_t)P * n_clients * sizeof(double),
                                cudaMemcpyDeviceToHost, e.stream));
      LR_CUDA(cudaStreamSynchronize(e.stream));
      // getMeanLLK: accumulated llk / accumulated weight (weight 1 per frame)
      for (long t = 0; t < P; t++) sum_w[o] += h_w[t];
      for (int i = 0; i < n_clients; i++)
        for (long t = 0; t < P; t++) sum_c[(size_t)i * n_out + o] += h_c[(size_t)i * P + t];
      cnt[o] += (double)P;
    }
  }
  for (size_t o = 0; o < n_out; o++) {
    mean_llk_world[o] = cnt[o] > 0 ? sum_w[o] / cnt[o] : 0.0;
    for (int i = 0; i < n_clients; i++)
      mean_llk_client[(size_t)i * n_out + o] = cnt[o] > 0 ? sum_c[(size_t)i * n_out + o] / cnt[o] : 0.0;
  }
  return LR_OK;
}

}  // extern "C"
