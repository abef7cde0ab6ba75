// engine.cpp -- the LIA_SpkTools batch functions re-expressed over the C ABI: Gmm handle,
// accumulateStatEM / trainModel (TrainTools.cpp), TVAcc (AccumulateTVStat.cpp).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <set>

#include "lia_host.h"

namespace lia {

// ------------------------------------------------------------------ Gmm
Gmm::Gmm(const MixtureGD &m, bool use_file_cst) {
  h_ = lr_gmm_create(m.C, m.D, m.w.data(), m.mean.data(), m.cov.data());
  if (!h_) LIA_THROW(std::string("engine: ") + lr_last_error());
  // RAW / XML files carry their own cst records; the reference scores with the stored value
  if (use_file_cst) LIA_CHECK(lr_gmm_set_cst(h_, m.cst.data()));
}
Gmm::~Gmm() { lr_gmm_destroy(h_); }
void Gmm::set(const MixtureGD &m) { LIA_CHECK(lr_gmm_set(h_, m.w.data(), m.mean.data(), m.cov.data())); }
void Gmm::get(MixtureGD &m) const {
  LIA_CHECK(lr_gmm_get(h_, m.w.data(), m.mean.data(), m.cov.data(), m.covinv.data(), m.cst.data(),
                       m.det.data()));
}

// ------------------------------------------------------------------ TrainTools
TrainCfg::TrainCfg(const Config &c) {
  initVarianceFlooring = c.getDouble("initVarianceFlooring");
  initVarianceCeiling = c.getDouble("initVarianceCeiling");
  finalVarianceFlooring = c.getDouble("finalVarianceFlooring");
  finalVarianceCeiling = c.getDouble("finalVarianceCeiling");
  nbTrainIt = c.getLong("nbTrainIt");
  baggedFrameProbability = c.getDouble("baggedFrameProbability");
  normalizeModel = c.getBool("normalizeModel", false);
  if (normalizeModel) {
    normalizeModelMeanOnly = c.getBool("normalizeModelMeanOnly", false);
    if (normalizeModelMeanOnly) normalizeModelNbIt = c.getLong("normalizeModelNbIt");
  }
  componentReduction = c.getBool("componentReduction", false);
  if (componentReduction) targetDistribCount = c.getLong("targetMixtureDistribCount");
}

void normalizeMixture(MixtureGD &m, long nbIt, bool meanOnly) {
  const int C = m.C, D = m.D;
  std::vector<double> gm(D), gc(D);
  for (long it = 0; it < nbIt; it++) {
    // mixtureFusion: sequential gaussianFusion (:258-283) -- exact moment matching, accumulated in the same order
    for (int i = 0; i < D; i++) {
      gm[i] = m.mean[i];
      gc[i] = m.cov[i];
    }
    double wacc = m.w[0];
    for (int k = 1; k < C; k++) {
      const double a1 = m.w[k] / (m.w[k] + wacc), a2 = 1.0 - a1;
      for (int i = 0; i < D; i++) {
        const double d = m.mean[(size_t)k * D + i] - gm[i];
        gc[i] = a1 * m.cov[(size_t)k * D + i] + a2 * gc[i] + a1 * a2 * d * d;
        gm[i] = a1 * m.mean[(size_t)k * D + i] + a2 * gm[i];
      }
      wacc += m.w[k];
    }
    for (int k = 0; k < C; k++)
      for (int i = 0; i < D; i++) {
        const size_t e = (size_t)k * D + i;
        m.mean[e] = (m.mean[e] - gm[i]) / std::sqrt(gc[i]);
        if (!meanOnly) m.cov[e] /= gc[i];
      }
  }
  m.computeAll();
}

void reduceToTopWeights(MixtureGD &m, size_t nbTop) {
  const size_t C = (size_t)m.C, D = (size_t)m.D;
  if (nbTop >= C) return;
  std::vector<size_t> order(C);
  for (size_t k = 0; k < C; k++) order[k] = k;
  // TabWeight sorts by decreasing weight with qsort (GeneralTools.cpp:277-280; ties are unordered there, by index here)
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return m.w[a] > m.w[b]; });
  std::vector<char> keep(C, 0);
  for (size_t i = 0; i < nbTop; i++) keep[order[i]] = 1;
  MixtureGD out;
  out.id = m.id;
  out.resize((int)nbTop, m.D);
  size_t o = 0;
  double tot = 0.0;
  for (size_t k = 0; k < C; k++)
    if (keep[k]) {
      out.w[o] = m.w[k];
      tot += m.w[k];
      std::copy(m.mean.begin() + k * D, m.mean.begin() + (k + 1) * D, out.mean.begin() + o * D);
      std::copy(m.cov.begin() + k * D, m.cov.begin() + (k + 1) * D, out.cov.begin() + o * D);
      o++;
    }
  for (size_t k = 0; k < nbTop; k++) out.w[k] /= tot;
  out.computeAll();
  m = out;
}

double setItParameter(double begin, double end, int nbIt, int it) {
  if (nbIt < 2) return begin;
  double itVal = (begin - end) / ((double)nbIt - 1.0);
  return begin - itVal * it;
}

static bool baggedFrame(double p) { return ((double)rand() / (double)RAND_MAX) < p; }

SegCluster baggedSegments(const SegCluster &in, double p, long minLen, long maxLen) {
  SegCluster out;
  for (const Seg &seg : in) {
    long begin = seg.begin, left = seg.length;
    while (true) {
      long verify = std::min(std::max(left, minLen), maxLen);
      bool move = left <= verify;
      long length = move ? left : verify;
      if (length > 0 && baggedFrame(p)) out.push_back({seg.source, begin, length, seg.label});
      if (move) break;
      left -= length;
      begin += length;
    }
  }
  return out;
}

void EmAcc::reset(int C, int D) {
  occ.assign(C, 0.0);
  m1.assign((size_t)C * D, 0.0);
  m2.assign((size_t)C * D, 0.0);
  n = 0;
}

double accumulateStatEM(const FeatureServer &fs, const Gmm &g, const SegCluster &segs, EmAcc &acc,
                        double weight) {
  std::vector<lr_seg> es = toEngineSegs(fs, segs);
  double llk = 0.0;
  LIA_CHECK(lr_gmm_em_accumulate(g.h(), fs.data(), fs.getFeatureCount(), fs.ld(), es.data(), es.size(),
                                 weight, acc.occ.data(), acc.m1.data(), acc.m2.data(), &llk, &acc.n));
  return llk;
}

void computeMeanCov(const std::vector<TrainStream> &streams, std::vector<double> &mean, std::vector<double> &cov) {
  // one FrameAccGD over the selected frames of every stream (TrainTools.cpp:593-602): gather them
  // (labels may skip frames), then one device pass
  if (streams.empty()) LIA_THROW("computeMeanCov: no input stream");
  const int D = streams[0].fs->getVectSize();
  std::vector<float> sel;
  for (const TrainStream &st : streams)
    for (auto &s : st.segs) {
      const float *p = st.fs->data() + (st.fs->getFirstFeatureIndexOfASource(s.source) + s.begin) * st.fs->ld();
      if ((int)st.fs->ld() == D) {
        sel.insert(sel.end(), p, p + (size_t)s.length * D);
      } else {
        for (long t = 0; t < s.length; t++) sel.insert(sel.end(), p + (size_t)t * st.fs->ld(), p + (size_t)t * st.fs->ld() + D);
      }
    }
  if (sel.empty()) LIA_THROW("computeMeanCov: no selected frame");
  mean.assign(D, 0.0);
  cov.assign(D, 0.0);
  LIA_CHECK(lr_frames_mean_cov(sel.data(), sel.size() / D, D, D, mean.data(), cov.data()));
}

void computeMeanCov(const FeatureServer &fs, const SegCluster &segs, std::vector<double> &mean,
                    std::vector<double> &cov) {
  computeMeanCov(std::vector<TrainStream>{TrainStream{&fs, segs, 1.0}}, mean, cov);
}

void mixtureInit(const std::vector<TrainStream> &streams, const std::vector<double> &globalCov,
                 const Config &c, MixtureGD &world) {
  // mean of randomly picked 3..7-frame chunks per component, cov = global cov, equal weights
  // (TrainTools.cpp:674-766; every stream contributes nbFrameToSelect x weight frames per component)
  const long minLen = c.getLong("baggedMinimalLength", 3), maxLen = c.getLong("baggedMaximalLength", 7);
  const double nbFrameToSelect = (double)c.getLong("nbFrameToSelect", 50);
  const int C = world.C, D = world.D;
  std::vector<double> sum((size_t)C * D, 0.0), cnt(C, 0.0);
  for (size_t stream = 0; stream < streams.size(); stream++) {
    const FeatureServer &fs = *streams[stream].fs;
    const SegCluster &segs = streams[stream].segs;
    const long total = totalFrame(segs);
    double proba = nbFrameToSelect * streams[stream].weight / (double)std::max(1L, total);
    // a probability above 1 becomes several passes (the reference's loop at :707-711 does not terminate
    // in that case; this is trainModelStream's rule, :1062-1066)
    long baggedIt = 1;
    if (proba > 1) {
      baggedIt = (long)proba + 1;
      proba /= (double)baggedIt;
    }
    for (long it = 0; it < baggedIt; it++) {
      srand((unsigned)(((stream + 1) * 100) + (it + 1)));  // :737
      // one pass over the segments, each chunk assigned to a random component when selected
      for (const Seg &seg : segs) {
        long begin = seg.begin, left = seg.length;
        const size_t first = fs.getFirstFeatureIndexOfASource(seg.source);
        while (left > 0) {
          long length = std::min(std::min(std::max(left, minLen), maxLen), left);
          for (int k = 0; k < C; k++) {
            if (!baggedFrame(proba)) continue;
            for (long t = 0; t < length; t++) {
              const float *x = fs.data() + (first + begin + t) * fs.ld();
              for (int i = 0; i < D; i++) sum[(size_t)k * D + i] += x[i];
            }
            cnt[k] += (double)length;
          }
          left -= length;
          begin += length;
        }
      }
    }
  }
  for (int k = 0; k < C; k++) {
    for (int i = 0; i < D; i++) {
      world.mean[(size_t)k * D + i] = cnt[k] > 0 ? sum[(size_t)k * D + i] / cnt[k] : 0.0;
      world.cov[(size_t)k * D + i] = globalCov[i];
    }
    world.w[k] = 1.0 / C;
  }
  world.computeAll();
}

void mixtureInit(const FeatureServer &fs, const SegCluster &segs, const std::vector<double> &globalCov,
                 const Config &c, MixtureGD &world) {
  mixtureInit(std::vector<TrainStream>{TrainStream{&fs, segs, 1.0}}, globalCov, c, world);
}

void trainModel(const Config &c, const std::vector<TrainStream> &streams, const std::vector<double> &globalCov,
                MixtureGD &world, const TrainCfg &cfg) {
  // trainModelStream (TrainTools.cpp:1030-1110): every iteration draws, per stream, bagged 3..7-frame
  // chunks with probability nbFrameToSelect x weight / totalFrame(stream) -- several passes when that
  // exceeds 1 -- and accumulates them all into ONE EM accumulator
  const long minLen = c.getLong("baggedMinimalLength", 3), maxLen = c.getLong("baggedMaximalLength", 7);
  const long initRand = c.getLong("initRand", 0);
  const bool verbose = c.getBool("verbose", false);
  std::unique_ptr<Gmm> gp(new Gmm(world));
  EmAcc acc;
  const Shard &sh = Shard::get();
  const long initialDistribCount = world.C;
  for (long it = 0; it < cfg.nbTrainIt; it++) {
    Gmm &g = *gp;
    const double flooring = setItParameter(cfg.initVarianceFlooring, cfg.finalVarianceFlooring, (int)cfg.nbTrainIt, (int)it);
    const double ceiling = setItParameter(cfg.initVarianceCeiling, cfg.finalVarianceCeiling, (int)cfg.nbTrainIt, (int)it);
    unsigned long nbTotalFrame = 0;  // :1054 (truncated per stream like the reference)
    for (const TrainStream &st : streams) nbTotalFrame += (unsigned long)((double)totalFrame(st.segs) * st.weight);
    const double nbFrameToSelect = cfg.baggedFrameProbability * (double)nbTotalFrame;
    acc.reset(world.C, world.D);  // emAcc.resetEM()
    double llk = 0.0;
    for (size_t stream = 0; stream < streams.size(); stream++) {
      const TrainStream &st = streams[stream];
      long nbBaggedIt = 1;
      double baggedProba = nbFrameToSelect * st.weight / (double)std::max(1L, totalFrame(st.segs));
      if (baggedProba > 1) {
        nbBaggedIt = (long)baggedProba + 1;
        baggedProba /= (double)nbBaggedIt;
      }
      for (long baggedIt = 0; baggedIt < nbBaggedIt; baggedIt++) {
        srand((unsigned)(((it + 1 + initRand) * 200) + (((stream + 1) * 20) + (baggedIt + 1))));  // :1070
        SegCluster bagged = baggedProba >= 1.0 ? st.segs : baggedSegments(st.segs, baggedProba, minLen, maxLen);
        // several ranks: each accumulates a contiguous range of the (identically bagged) segments balanced
        // by frames; ONE all-reduce of {occ, m1, m2, llk, n} per iteration follows (emAcc.addAccEM,
        // AccumulateStat.cpp:286-292)
        if (sh.world > 1) {
          std::vector<double> wgt(bagged.size());
          for (size_t i = 0; i < bagged.size(); i++) wgt[i] = (double)bagged[i].length;
          auto r = sh.rangeByWeight(wgt);
          SegCluster mine(bagged.begin() + r.first, bagged.begin() + r.second);
          if (!mine.empty()) llk += accumulateStatEM(*st.fs, g, mine, acc);
        } else if (!bagged.empty()) {
          llk += accumulateStatEM(*st.fs, g, bagged, acc);
        }
      }
    }
    if (sh.world > 1) {
      const size_t cC = world.C, cd = (size_t)world.C * world.D;
      std::vector<double> pack(cC + 2 * cd + 2);
      std::copy(acc.occ.begin(), acc.occ.end(), pack.begin());
      std::copy(acc.m1.begin(), acc.m1.end(), pack.begin() + cC);
      std::copy(acc.m2.begin(), acc.m2.end(), pack.begin() + cC + cd);
      pack[cC + 2 * cd] = llk;
      pack[cC + 2 * cd + 1] = acc.n;
      LIA_CHECK(lr_allreduce_host(pack.data(), pack.size()));
      std::copy(pack.begin(), pack.begin() + cC, acc.occ.begin());
      std::copy(pack.begin() + cC, pack.begin() + cC + cd, acc.m1.begin());
      std::copy(pack.begin() + cC + cd, pack.begin() + cC + 2 * cd, acc.m2.begin());
      llk = pack[cC + 2 * cd];
      acc.n = pack[cC + 2 * cd + 1];
    }
    // *world = emAcc.getEM(); varianceControl(world, flooring, ceiling, globalCov)  (:1076-1077)
    LIA_CHECK(lr_gmm_em_update(g.h(), acc.occ.data(), acc.m1.data(), acc.m2.data(), flooring, ceiling,
                               globalCov.data()));
    if (cfg.componentReduction || cfg.normalizeModel) {  // :1078-1098, on the host: O(C D) between two EM passes
      g.get(world);
      bool rebuilt = false;
      if (cfg.componentReduction) {
        const double diff = (double)(initialDistribCount - cfg.targetDistribCount) / (double)cfg.nbTrainIt;
        long nbTop = initialDistribCount - (long)((double)(it + 1) * diff);
        if (it == cfg.nbTrainIt - 1) nbTop = cfg.targetDistribCount;
        if (nbTop < 1) LIA_THROW("componentReduction: targetMixtureDistribCount must be >= 1");
        if (nbTop < world.C) {
          reduceToTopWeights(world, (size_t)nbTop);
          rebuilt = true;
        }
      }
      if (cfg.normalizeModel) normalizeMixture(world, cfg.normalizeModelNbIt, cfg.normalizeModelMeanOnly);
      if (rebuilt)
        gp.reset(new Gmm(world));  // the device model has a fixed number of components
      else
        g.set(world);
    }
    if (verbose)
      std::cout << "ML (partial) estimate it[" << it << "] (take care, it corresponds to the previous it,0 means init likelihood) = "
                << (acc.n > 0 ? llk / acc.n : 0.0) << std::endl;
  }
  gp->get(world);
}

void trainModel(const Config &c, const FeatureServer &fs, const SegCluster &segs,
                const std::vector<double> &globalCov, MixtureGD &world, const TrainCfg &cfg) {
  trainModel(c, std::vector<TrainStream>{TrainStream{&fs, segs, 1.0}}, globalCov, world, cfg);
}

// ------------------------------------------------------------------ TopGauss
double TopGauss::compute(const MixtureGD &ubm, const FeatureServer &fs, const std::string &file, const Config &c) {
  const double topD = c.getDouble("topGauss");
  const int K = (int)c.getLong("topDistribsCount", 10);
  if (topD >= 1.0 && (long)topD > K) LIA_THROW("topGauss exceeds topDistribsCount");
  const bool complete = c.getString("computeLLKWithTopDistribs", "COMPLETE") == "COMPLETE";
  const double minLLK = c.getDouble("minLLK", -200.0), maxLLK = c.getDouble("maxLLK", 200.0);
  SegCluster all = selectedSegments(c, fs, c.getParam("labelSelectedFrames")), sel;
  for (const Seg &s : all)
    if (s.source == file) sel.push_back(s);
  nt = (uint64_t)totalFrame(sel);
  nbg.assign(nt, 0);
  idx.clear();
  snsw.clear();
  snsl.clear();
  nbgcnt = 0;
  if (nt == 0) return 0.0;
  const size_t D = fs.ld();
  std::vector<float> X((size_t)nt * D);
  size_t t = 0;
  for (const Seg &s : sel) {
    const float *p = fs.data() + (fs.getFirstFeatureIndexOfASource(s.source) + (size_t)s.begin) * D;
    std::copy(p, p + (size_t)s.length * D, X.begin() + t * D);
    t += (size_t)s.length;
  }
  Gmm g(ubm, true);
  std::vector<double> llk(nt), top((size_t)nt * K), rest(nt), restw(nt);
  std::vector<uint32_t> ix((size_t)nt * K);
  LIA_CHECK(lr_gmm_llk_topk(g.h(), X.data(), nt, D, K, complete ? 1 : 0, minLLK, maxLLK, llk.data(), ix.data(), top.data(),
                            rest.data(), restw.data()));
  double sum = 0.0;
  for (uint64_t f = 0; f < nt; f++) {
    sum += llk[f];
    const double lkTot = std::exp(llk[f]);
    uint64_t n = 0;
    if (topD < 1.0) {
      double val = 0.0;
      for (int j = 0; j < K; j++) {  // :173-179
        if (val > topD * lkTot) break;
        val += top[f * K + j];
        n++;
      }
    } else {
      n = (uint64_t)topD;
    }
    nbg[f] = n;
    nbgcnt += n;
    double w = 1.0, l = lkTot;  // :183-193
    for (uint64_t j = 0; j < n; j++) {
      idx.push_back(ix[f * K + j]);
      w -= ubm.w[ix[f * K + j]];
      l -= top[f * K + j];
    }
    snsw.push_back(w);
    snsl.push_back(l < 1e-200 ? 1e-200 : l);
  }
  return sum / (double)nt;
}
void TopGauss::write(const std::string &file, const Config &c) const {
  const std::string name = c.getParam("nbGaussianFilesDir") + file;
  std::ofstream f(name.c_str(), std::ios::out | std::ios::binary);
  if (!f) LIA_THROW("Cannot find nbGaussian file " + name);
  f.write((const char *)&nt, 8);
  f.write((const char *)&nbgcnt, 8);
  f.write((const char *)nbg.data(), 8 * nbg.size());
  f.write((const char *)idx.data(), 8 * idx.size());
  f.write((const char *)snsw.data(), 8 * snsw.size());
  f.write((const char *)snsl.data(), 8 * snsl.size());
}
void TopGauss::read(const std::string &file, const Config &c) {
  const std::string name = c.getParam("nbGaussianFilesDir") + file;
  std::ifstream f(name.c_str(), std::ios::in | std::ios::binary);
  if (!f) LIA_THROW("Cannot find nbGaussian file " + name);
  f.read((char *)&nt, 8);
  f.read((char *)&nbgcnt, 8);
  nbg.assign(nt, 0);
  idx.assign(nbgcnt, 0);
  snsw.assign(nt, 0.0);
  snsl.assign(nt, 0.0);
  f.read((char *)nbg.data(), 8 * nt);
  f.read((char *)idx.data(), 8 * nbgcnt);
  f.read((char *)snsw.data(), 8 * nt);
  f.read((char *)snsl.data(), 8 * nt);
  if (!f) LIA_THROW("nbGaussian file " + name + " is truncated");
}

// ------------------------------------------------------------------ MAP
MAPCfg::MAPCfg(const Config &c) {
  mean = c.getBool("meanAdapt", false);
  var = c.getBool("varAdapt", false);
  weight = c.getBool("weightAdapt", false);
  method = c.getParam("MAPAlgo");
  if (method == "MAPConst" || method == "MAPConst2") {  // a priori probability of the initial model (:113-117)
    if (mean) r[0] = c.getDouble("MAPAlphaMean");
    if (var) r[1] = c.getDouble("MAPAlphaVar");
    if (weight) r[2] = c.getDouble("MAPAlphaWeight");
  } else if (method == "MAPOccDep" || method == "MAPModelBased") {
    if (mean) r[0] = c.getDouble("MAPRegFactorMean");
    if (var) r[1] = c.getDouble("MAPRegFactorVar");
    if (weight) r[2] = c.getDouble("MAPRegFactorWeight");
  } else {
    LIA_THROW("mapAlgo[" + method + "] is not implemented by this engine (MAPOccDep | MAPModelBased | MAPConst | MAPConst2)");
  }
  nbTrainIt = c.getLong("nbTrainIt", 1);
  baggedFrameProbability = c.getDouble("baggedFrameProbability", 1.0);
  normalizeModel = c.getBool("normalizeModel", false);
  if (normalizeModel) {
    normalizeModelMeanOnly = c.getBool("normalizeModelMeanOnly", false);
    if (normalizeModelMeanOnly) normalizeModelNbIt = c.getLong("normalizeModelNbIt");
  }
}

void computeMAPOccDep(const MixtureGD &w, MixtureGD &client, const MAPCfg &cfg, double frameCount) {
  MixtureGD t = w;  // a priori data
  const int C = w.C, D = w.D;
  for (int k = 0; k < C; k++) {
    const double alpha = client.w[k] * frameCount;  // occupation of the component
    if (cfg.mean) {
      const double a = alpha / (alpha + cfg.r[0]);
      for (int i = 0; i < D; i++)
        t.mean[(size_t)k * D + i] = (1 - a) * w.mean[(size_t)k * D + i] + a * client.mean[(size_t)k * D + i];
    }
    if (cfg.var) {
      const double a = alpha / (alpha + cfg.r[1]);
      for (int i = 0; i < D; i++) {
        const double dm = w.mean[(size_t)k * D + i] - client.mean[(size_t)k * D + i];
        t.cov[(size_t)k * D + i] = (1 - a) * w.cov[(size_t)k * D + i] + a * client.cov[(size_t)k * D + i] +
                                   (1 - a) * a * dm * dm;
      }
    }
  }
  if (cfg.weight) {
    double sum = 0;
    for (int k = 0; k < C; k++) {
      const double alpha = client.w[k] * frameCount, a = alpha / (alpha + cfg.r[2]);
      t.w[k] = a * client.w[k] + (1 - a) * w.w[k];
      sum += t.w[k];
    }
    for (int k = 0; k < C; k++) t.w[k] /= sum;
  }
  t.computeAll();
  t.id = client.id;
  client = t;
}

// Direct mean-only interpolation with a constant a priori weight; the variance / weight branches are TODOs in the
// reference (:372-379), so the result keeps the initial model's weights and variances
void computeMAPConst(const MixtureGD &w, MixtureGD &client, const MAPCfg &cfg) {
  MixtureGD t = w;
  if (cfg.mean) {
    const double alpha = cfg.r[0];
    for (size_t e = 0; e < t.mean.size(); e++) t.mean[e] = alpha * w.mean[e] + (1 - alpha) * client.mean[e];
  }
  t.id = client.id;
  client = t;
}
// ... the same with the component weights in the interpolation (:388-419)
void computeMAPConst2(const MixtureGD &w, MixtureGD &client, const MAPCfg &cfg) {
  MixtureGD t = w;
  if (cfg.mean) {
    const double alpha = cfg.r[0];
    for (int k = 0; k < w.C; k++)
      for (int i = 0; i < w.D; i++) {
        const size_t e = (size_t)k * w.D + i;
        t.mean[e] = (alpha * w.w[k] * w.mean[e] + (1 - alpha) * client.w[k] * client.mean[e]) /
                    (w.w[k] * alpha + client.w[k] * (1 - alpha));
      }
  }
  t.id = client.id;
  client = t;
}
void computeMAP(const MixtureGD &w, MixtureGD &client, const MAPCfg &cfg, double frameCount) {
  if (cfg.method == "MAPConst") computeMAPConst(w, client, cfg);
  else if (cfg.method == "MAPConst2") computeMAPConst2(w, client, cfg);
  else computeMAPOccDep(w, client, cfg, frameCount);  // MAPOccDep and MAPModelBased share the formulas (:445-489, :491-545)
}

void adaptModel(const Config &c, const FeatureServer &fs, const SegCluster &segs, const MixtureGD &apriori,
                MixtureGD &client, const MAPCfg &cfg) {
  const long minLen = c.getLong("baggedMinimalLength", 3), maxLen = c.getLong("baggedMaximalLength", 7);
  Gmm g(client);
  EmAcc acc;
  for (long it = 0; it < cfg.nbTrainIt; it++) {
    SegCluster bagged = cfg.baggedFrameProbability >= 1.0 ? segs : baggedSegments(segs, cfg.baggedFrameProbability, minLen, maxLen);
    acc.reset(client.C, client.D);
    srand((unsigned)it);
    accumulateStatEM(fs, g, bagged, acc);
    // clientMixture = emAcc.getEM() on the device (no variance control), then MAP on the host (O(C D))
    LIA_CHECK(lr_gmm_em_update(g.h(), acc.occ.data(), acc.m1.data(), acc.m2.data(), 0.0, 0.0, nullptr));
    g.get(client);
    computeMAP(apriori, client, cfg, acc.n);
    if (cfg.normalizeModel) normalizeMixture(client, cfg.normalizeModelNbIt, cfg.normalizeModelMeanOnly);  // :898
    g.set(client);
  }
}

// ------------------------------------------------------------------ one process per GPU
Shard &Shard::get() {
  static Shard s;
  return s;
}
std::pair<size_t, size_t> Shard::range(size_t n) const {
  const size_t per = n / world, rem = n % world;
  const size_t b = rank * per + std::min<size_t>(rank, rem);
  return {b, b + per + ((size_t)rank < rem ? 1 : 0)};
}
std::pair<size_t, size_t> Shard::rangeByWeight(const std::vector<double> &w) const {
  // cut k goes where the prefix sum is closest to k / world of the total (cuts stay monotone)
  std::vector<double> prefix(w.size() + 1, 0.0);
  for (size_t i = 0; i < w.size(); i++) prefix[i + 1] = prefix[i] + w[i];
  std::vector<size_t> cuts(world + 1, w.size());
  cuts[0] = 0;
  size_t i = 0;
  for (int k = 1; k < world; k++) {
    const double target = prefix.back() * k / world;
    while (i < w.size() && prefix[i + 1] < target) i++;
    size_t cut = (i < w.size() && target - prefix[i] > prefix[i + 1] - target) ? i + 1 : i;
    cuts[k] = std::max(cut, cuts[k - 1]);
  }
  return {cuts[rank], cuts[rank + 1]};
}
void Shard::barrier() const {
  if (world > 1) {
    double one = 1.0;
    LIA_CHECK(lr_allreduce_host(&one, 1));
  }
}
void initEngine(const Config &c) {
  auto env = [](const char *k) -> const char * { const char *v = getenv(k); return (v && *v) ? v : nullptr; };
  Shard &s = Shard::get();
  s.world = (int)c.getLong("lrWorldSize", env("LR_WORLD_SIZE") ? atol(env("LR_WORLD_SIZE")) : 1);
  s.rank = (int)c.getLong("lrRank", env("LR_RANK") ? atol(env("LR_RANK")) : 0);
  if (s.world < 1 || s.rank < 0 || s.rank >= s.world) LIA_THROW("lrRank / lrWorldSize out of range");
  if (c.existsParam("device"))
    LIA_CHECK(lr_init((int)c.getLong("device")));
  else if (s.world > 1)
    LIA_CHECK(lr_init(s.rank));
  if (s.world > 1) {
    const std::string f = c.getString("lrCommFile", env("LR_COMM_FILE") ? env("LR_COMM_FILE") : "");
    if (f.empty()) LIA_THROW("lrWorldSize > 1 needs lrCommFile (rendezvous file on a shared filesystem)");
    LIA_CHECK(lr_comm_init_file(s.rank, s.world, f.c_str()));
  }
}

// ------------------------------------------------------------------ TVAcc
// With several ranks a TVAcc holds the rank's contiguous range of NDX lines (AccumulateTVStat.cpp:498-507
// splits the lines over threads the same way); the statistics files on disk stay whole.
void TVAcc::shardLines() {
  totalLines_ = lines_.size();
  totalSessions_ = 0;
  for (auto &l : lines_) totalSessions_ += l.size();
  auto r = Shard::get().range(lines_.size());
  firstLine_ = r.first;
  if (r.second - r.first < lines_.size())
    lines_ = std::vector<std::vector<std::string>>(lines_.begin() + r.first, lines_.begin() + r.second);
  if (lines_.empty()) LIA_THROW("TVAcc: fewer NDX lines than ranks");
}

TVAcc::TVAcc(const std::string &ndxFile, const Config &c) : cfg_(c) {
  XList ndx(ndxFile);
  lines_ = ndx.lines();
  if (lines_.empty()) LIA_THROW("TVAcc: empty NDX list " + ndxFile);
  shardLines();
  init(c);
}
TVAcc::TVAcc(const std::vector<std::vector<std::string>> &fileLines, const Config &c)
    : cfg_(c), lines_(fileLines) {
  if (lines_.empty()) LIA_THROW("TVAcc: empty file list");
  shardLines();
  init(c);
}
void TVAcc::init(const Config &c) {
  world_ = MixtureGD::loadFromConfig(c.getParam("inputWorldFilename"), c);
  R_ = (int)c.getLong("totalVariabilityNumber", 1);
  tv_ = lr_tv_create(world_.C, world_.D, R_, lines_.size(), world_.mean.data(), world_.covinv.data());
  if (!tv_) LIA_THROW(std::string("engine: ") + lr_last_error());
  N_ = Matrix(lines_.size(), world_.C);
  F_ = Matrix(lines_.size(), (size_t)world_.C * world_.D);
}
TVAcc::~TVAcc() { lr_tv_destroy(tv_); }

void TVAcc::computeAndAccumulateTVStat(const Config &c) {
  // every file of every NDX line, once in the FeatureServer; a file listed on several lines
  // feeds each of them (AccumulateTVStat.cpp:339-346)
  std::vector<std::string> all;
  {
    XList tmp;
    std::set<std::string> seen;
    for (auto &l : lines_)
      for (auto &f : l)
        if (seen.insert(f).second) all.push_back(f);
  }
  FeatureServer fs(c, all);
  SegCluster sel = selectedSegments(c, fs, c.getParam("labelSelectedFrames"));
  std::vector<lr_seg> segs;
  for (const Seg &s : sel)
    for (size_t line = 0; line < lines_.size(); line++)
      if (std::find(lines_[line].begin(), lines_[line].end(), s.source) != lines_[line].end()) {
        lr_seg e;
        e.begin = (int64_t)(fs.getFirstFeatureIndexOfASource(s.source) + s.begin);
        e.length = s.length;
        e.row = (int32_t)line;
        e.pad_ = 0;
        segs.push_back(e);
      }
  Gmm g(world_, true);
  LIA_CHECK(lr_gmm_bwstats(g.h(), fs.data(), fs.getFeatureCount(), fs.ld(), segs.data(), segs.size(),
                           lines_.size(), N_.data.data(), F_.data.data()));
  LIA_CHECK(lr_tv_set_stats(tv_, N_.data.data(), F_.data.data()));
}

void TVAcc::loadT(const std::string &name, const Config &c) {
  Matrix T;
  T.load(c.getString("matrixFilesPath", "") + name + c.getString("loadMatrixFilesExtension", ""),
         c.getString("loadMatrixFormat", "DB"));
  if (T.cols < T.rows) {  // :636-639
    Matrix t(T.cols, T.rows);
    for (size_t i = 0; i < T.rows; i++)
      for (size_t j = 0; j < T.cols; j++) t(j, i) = T(i, j);
    T = t;
  }
  if ((long)T.rows != c.getLong("totalVariabilityNumber") || T.cols != (size_t)world_.C * world_.D)
    LIA_THROW("Incorrect dimension of TotalVariability Matrix");
  LIA_CHECK(lr_tv_set_T(tv_, T.data.data()));
}

Matrix TVAcc::getT() {
  Matrix T(R_, (size_t)world_.C * world_.D);
  LIA_CHECK(lr_tv_get_T(tv_, T.data.data()));
  return T;
}
void TVAcc::setStats(const Matrix &N, const Matrix &F) {
  if (N.rows != N_.rows || N.cols != N_.cols || F.rows != F_.rows || F.cols != F_.cols)
    LIA_THROW("TVAcc::setStats: incorrect dimension");
  N_ = N;
  F_ = F;
  reloadStats();
}
void TVAcc::setT(const Matrix &T) {
  if ((long)T.rows != R_ || T.cols != (size_t)world_.C * world_.D) LIA_THROW("TVAcc::setT: incorrect dimension");
  LIA_CHECK(lr_tv_set_T(tv_, T.data.data()));
}

void TVAcc::initT(const Config &c) {
  const size_t sv = (size_t)world_.C * world_.D;
  Matrix T(R_, sv);
  const std::string law = c.getString("randomInitLaw", "normal");
  double norm = 0.0;
  for (double v : world_.covinv) norm += v;
  if (law == "normal") {
    // ScoreWarp.cpp:68-79 Box-Muller on libc rand(), cosine branch, float ratio
    double x1 = rand() / (float)RAND_MAX, x2;
    for (size_t i = 0; i < T.rows; i++)
      for (size_t j = 0; j < T.cols; j++) {
        double val;
        do {
          x2 = x1;
          x1 = rand() / (float)RAND_MAX;
          val = std::sqrt(-2.0 * std::log(x1)) * std::cos(2.0 * 3.14159265358979323846 * x2);
        } while (std::isnan(val) || std::isinf(val));
        T(i, j) = val * norm * 0.001;
      }
  } else if (law == "uniform") {
    srand48((long)(sv * R_));
    for (auto &v : T.data) v = drand48() * norm / (double)sv;
  } else {
    LIA_THROW("Selected random initialization law does not exist");
  }
  LIA_CHECK(lr_tv_set_T(tv_, T.data.data()));
}

void TVAcc::saveT(const std::string &name, const Config &c) {
  Matrix T(R_, (size_t)world_.C * world_.D);
  LIA_CHECK(lr_tv_get_T(tv_, T.data.data()));
  T.save(c.getString("matrixFilesPath", "") + name + c.getString("saveMatrixFilesExtension", ""),
         c.getString("saveMatrixFormat", "DB"));
}

void TVAcc::loadN(const Config &c) {
  N_.load(c.getString("matrixFilesPath", "") + c.getParam("nullOrderStatSpeaker") +
              c.getString("loadMatrixFilesExtension", ""),
          c.getString("loadMatrixFormat", "DB"));
  if (N_.rows != totalLines_ || N_.cols != (size_t)world_.C) LIA_THROW("Incorrect dimension of N Matrix");
  if (lines_.size() != totalLines_) {  // several ranks: the file holds every line, keep this rank's rows
    Matrix mine(lines_.size(), N_.cols);
    std::copy(N_.data.begin() + firstLine_ * N_.cols, N_.data.begin() + (firstLine_ + lines_.size()) * N_.cols,
              mine.data.begin());
    N_ = mine;
  }
}
void TVAcc::loadF_X(const Config &c) {
  F_.load(c.getString("matrixFilesPath", "") + c.getParam("firstOrderStatSpeaker") +
              c.getString("loadMatrixFilesExtension", ""),
          c.getString("loadMatrixFormat", "DB"));
  if (F_.rows != totalLines_ || F_.cols != (size_t)world_.C * world_.D)
    LIA_THROW("Incorrect dimension of F_X Matrix");
  if (lines_.size() != totalLines_) {
    Matrix mine(lines_.size(), F_.cols);
    std::copy(F_.data.begin() + firstLine_ * F_.cols, F_.data.begin() + (firstLine_ + lines_.size()) * F_.cols,
              mine.data.begin());
    F_ = mine;
  }
  LIA_CHECK(lr_tv_set_stats(tv_, N_.data.data(), F_.data.data()));
}
void TVAcc::saveAccs(const Config &c) {
  std::string fx = "F_X.mat", n = "N.mat";
  const std::string path = c.getString("matrixFilesPath", ""), ext = c.getString("saveMatrixFilesExtension", "");
  if (c.existsParam("nullOrderStatSpeaker")) n = path + c.getParam("nullOrderStatSpeaker") + ext;
  if (c.existsParam("firstOrderStatSpeaker")) fx = path + c.getParam("firstOrderStatSpeaker") + ext;
  const Shard &sh = Shard::get();
  if (sh.world == 1) {
    F_.save(fx, c.getString("saveMatrixFormat", "DB"));
    N_.save(n, c.getString("saveMatrixFormat", "DB"));
    return;
  }
  // several ranks: the rows are gathered (equal-size blocks of ceil(lines / world) rows, then cut) and rank 0
  // writes the files the single-process run writes
  const size_t blockRows = (totalLines_ + sh.world - 1) / sh.world;
  auto gather = [&](const Matrix &mine, const std::string &file) {
    std::vector<double> send(blockRows * mine.cols, 0.0), recv(send.size() * sh.world);
    std::copy(mine.data.begin(), mine.data.end(), send.begin());
    LIA_CHECK(lr_allgather_host(send.data(), send.size(), recv.data()));
    if (sh.rank != 0) return;
    Matrix all(totalLines_, mine.cols);
    Shard probe = sh;
    for (int r = 0; r < sh.world; r++) {
      probe.rank = r;
      auto rg = probe.range(totalLines_);
      std::copy(recv.begin() + (size_t)r * send.size(), recv.begin() + (size_t)r * send.size() + (rg.second - rg.first) * mine.cols,
                all.data.begin() + rg.first * mine.cols);
    }
    all.save(file, c.getString("saveMatrixFormat", "DB"));
  };
  gather(F_, fx);
  gather(N_, n);
}

void TVAcc::substractM() { LIA_CHECK(lr_tv_subtract_m(tv_)); }
void TVAcc::estimateTETt() { LIA_CHECK(lr_tv_estimate_tett(tv_)); }
void TVAcc::estimateW() { LIA_CHECK(lr_tv_estimate_w(tv_)); }
void TVAcc::estimateAandC() { LIA_CHECK(lr_tv_estimate_a_and_c(tv_)); }
void TVAcc::resetTmpAcc() { LIA_CHECK(lr_tv_reset_tmp_acc(tv_)); }
void TVAcc::updateTestimate() {
  // several ranks: the one exchange step of the iteration + the component-sharded M-step
  if (Shard::get().world > 1) {
    double total = 0;
    LIA_CHECK(lr_tv_exchange_sharded(tv_, (double)lines_.size(), &total));
  } else {
    LIA_CHECK(lr_tv_update_t(tv_));
  }
}
void TVAcc::minDivergence() {
  // _n_sessions (:137) counts the sessions of the WHOLE list; the accumulators were summed over the ranks
  LIA_CHECK(lr_tv_min_divergence(tv_, (double)totalSessions_));
}
void TVAcc::orthonormalizeT() { LIA_CHECK(lr_tv_orthonormalize_t(tv_)); }

void TVAcc::normTMatrix() { LIA_CHECK(lr_tv_norm_t(tv_)); }
void TVAcc::normStatistics() { LIA_CHECK(lr_tv_norm_statistics(tv_)); }
Matrix TVAcc::getWeightedCov(const std::vector<double> &weight) {
  if (weight.size() != (size_t)world_.C) LIA_THROW("getWeightedCov: weight vector size != distribCount");
  Matrix W(R_, R_);
  LIA_CHECK(lr_tv_weighted_cov(tv_, weight.data(), W.data.data()));
  return W;
}
void TVAcc::computeEigenProblem(const Matrix &EP, Matrix &eigenVect, long rank) {
  if (EP.rows != EP.cols) LIA_THROW("computeEigenProblem: matrix is not square");
  eigenVect = Matrix(EP.rows, (size_t)rank);
  std::vector<double> val((size_t)rank);
  LIA_CHECK(lr_eigen_problem((int)EP.rows, EP.data.data(), (int)rank, eigenVect.data.data(), val.data()));
}
Matrix TVAcc::approximateTcTc(const Matrix &Q) {
  if (Q.rows != (size_t)R_ || Q.cols != (size_t)R_) LIA_THROW("approximateTcTc: Q must be rankT x rankT");
  Matrix D(world_.C, R_);
  LIA_CHECK(lr_tv_approximate_tctc(tv_, Q.data.data(), D.data.data()));
  return D;
}
void TVAcc::estimateWUbmWeight(const Matrix &W) {
  if (W.rows != (size_t)R_ || W.cols != (size_t)R_) LIA_THROW("estimateWUbmWeight: W must be rankT x rankT");
  LIA_CHECK(lr_tv_estimate_w_ubm_weight(tv_, W.data.data()));
}
void TVAcc::estimateWEigenDecomposition(const Matrix &D, const Matrix &Q) {
  if (D.rows != (size_t)world_.C || D.cols != (size_t)R_ || Q.rows != (size_t)R_ || Q.cols != (size_t)R_)
    LIA_THROW("estimateWEigenDecomposition: D must be distribCount x rankT, Q rankT x rankT");
  LIA_CHECK(lr_tv_estimate_w_eigen_decomposition(tv_, D.data.data(), Q.data.data()));
}

void TVAcc::loadMeanEstimate(const std::vector<double> &mean) {
  if (mean.size() != (size_t)world_.C * world_.D) LIA_THROW("Incorrect dimension of meanEstimate vector");
  LIA_CHECK(lr_tv_set_mean(tv_, mean.data()));
}
void TVAcc::reloadStats() { LIA_CHECK(lr_tv_set_stats(tv_, N_.data.data(), F_.data.data())); }
Matrix TVAcc::getUbmMeans() {
  Matrix m(1, (size_t)world_.C * world_.D);
  LIA_CHECK(lr_tv_get_mean(tv_, m.data.data()));
  return m;
}

Matrix TVAcc::getW() {
  Matrix W(lines_.size(), R_);
  LIA_CHECK(lr_tv_get_W(tv_, W.data.data()));
  return W;
}

void TVAcc::saveWbyFile(const Config &c) {
  const std::string path = c.getParam("saveVectorFilesPath"), ext = c.getString("vectorFilesExtension", ".y");
  XList ids(c.getParam("targetIdList"));
  Matrix W = getW();
  size_t session = 0, lineNo = 0;
  for (auto &line : ids.lines()) {
    if (lineNo++ < firstLine_) continue;  // another rank's i-vector (one file per line: nothing to gather)
    if (session >= W.rows) break;
    Matrix y(1, R_);
    for (int i = 0; i < R_; i++) y(0, i) = W(session, i);
    y.save(path + line[0] + ext, c.getString("saveMatrixFormat", "DB"));
    session++;
  }
}

}  // namespace lia
