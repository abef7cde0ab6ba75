// drivers.cpp -- int Foo(Config&) entry points with the reference programs' parameter names and
// output formats.  Errors follow the reference: catch, print e.toString(), return 0.
#include <cmath>
#include <map>
#include <memory>
#include <set>
#include <algorithm>
#include <fstream>
#include <iostream>

#include "lia_host.h"

namespace lia {

static char setDecision(double llr, double threshold) { return llr > threshold ? 1 : 0; }  // GeneralTools.cpp:232

// "gender client decision seg [begin end] LLR" (IOFormat.cpp:112-120)
static void outputResultLine(double llr, const std::string &client, const std::string &seg,
                             const std::string &gender, int decision, std::ostream &os) {
  os << gender << " " << client << " " << decision << " " << seg << " " << llr << std::endl;
}
static void outputResultLine(double llr, const std::string &client, const std::string &seg, double b,
                             double e, const std::string &gender, int decision, std::ostream &os) {
  os << gender << " " << client << " " << decision << " " << seg << " " << b << " " << e << " " << llr
     << std::endl;
}

// ------------------------------------------------------------------ TrainWorld
int TrainWorld(Config &c) {
  try {
    const bool verbose = c.getBool("verbose", false);
    const std::string out = c.getParam("outputWorldFilename");
    const std::string label = c.getParam("labelSelectedFrames");
    TrainCfg cfg(c);
    // one stream (inputFeatureFilename: a feature file or a list of feature files) or several
    // (inputStreamList: one list per stream, weightStreamList: their weights; TrainWorld.cpp:120-141)
    std::vector<std::string> streamNames;
    std::vector<double> weights;
    if (c.existsParam("inputStreamList")) {
      streamNames = XList(c.getParam("inputStreamList")).allElements();
      if (streamNames.empty()) LIA_THROW("TrainWorld error:no input stream");
      weights.assign(streamNames.size(), 1.0 / (double)streamNames.size());
      if (c.existsParam("weightStreamList")) {
        std::vector<std::string> w = XList(c.getParam("weightStreamList")).allElements();
        if (w.size() != streamNames.size())
          LIA_THROW("TrainWorld error: number of weigths differs than number of input streams");
        for (size_t i = 0; i < w.size(); i++) weights[i] = std::stod(w[i]);
      }
    } else {
      streamNames.push_back(c.getParam("inputFeatureFilename"));
      weights.push_back(1.0);
    }
    std::vector<std::unique_ptr<FeatureServer>> servers;
    std::vector<TrainStream> streams;
    long nFrames = 0;
    for (size_t i = 0; i < streamNames.size(); i++) {
      const std::string &in = streamNames[i];
      std::vector<std::string> files;
      if (in.size() > 4 && in.compare(in.size() - 4, 4, ".lst") == 0)
        files = XList(in).allElements();
      else
        files.push_back(in);
      servers.emplace_back(new FeatureServer(c, files));
      SegCluster segs = selectedSegments(c, *servers.back(), label);
      if (segs.empty()) LIA_THROW("TrainWorld error: no frame selected with label " + label + " in stream " + in);
      nFrames += totalFrame(segs);
      streams.push_back(TrainStream{servers.back().get(), segs, weights[i]});
    }
    const int vectSize = servers[0]->getVectSize();
    std::vector<double> gMean, gCov;
    if (c.getBool("use01", false)) {
      gMean.assign(vectSize, 0.0);
      gCov.assign(vectSize, 1.0);
    } else {
      computeMeanCov(streams, gMean, gCov);
    }
    MixtureGD world;
    if (c.existsParam("inputWorldFilename")) {
      world = MixtureGD::loadFromConfig(c.getParam("inputWorldFilename"), c);
    } else {
      world.resize((int)c.getLong("mixtureDistribCount"), vectSize);
      mixtureInit(streams, gCov, c, world);  // (deterministic: every rank builds the same initial model)
      if (c.getBool("saveInitModel", true) && Shard::get().rank == 0) world.saveFromConfig(out + "init", c);
    }
    if (verbose)
      std::cout << "Train world model: " << world.C << " components, " << streams.size() << " stream(s), " << nFrames
                << " frames" << std::endl;
    trainModel(c, streams, gCov, world, cfg);
    if (Shard::get().rank == 0) world.saveFromConfig(out, c);
  } catch (std::exception &e) {
    std::cout << e.what() << std::endl;
  }
  return 0;
}

// ------------------------------------------------------------------ TrainTarget (MAPOccDep)
int TrainTarget(Config &c) {
  try {
    const std::string label = c.getParam("labelSelectedFrames");
    MAPCfg mapCfg(c);
    MixtureGD world = MixtureGD::loadFromConfig(c.getParam("inputWorldFilename"), c);
    XList ids(c.getParam("targetIdList"));  // "id file1 file2 ..." per line (TrainTarget.cpp:100-130)
    // options of the reference this engine does not implement are refused, not ignored
    if (c.getBool("useModelData", false)) LIA_THROW("useModelData (modelBasedadaptModel) is not implemented by this engine");
    if (c.existsParam("mixtureServer")) LIA_THROW("mixtureServer output is not implemented by this engine");
    if (c.getBool("outputAdaptParam", false) || c.existsParam("superVectors"))
      LIA_THROW("outputAdaptParam / superVectors is not implemented by this engine (TrainTarget --channelCompensation JFA writes supervectors)");
    const bool initByClient = c.getBool("initByClient", false);    // EM starts from the client's existing model (:136-139)
    const bool saveEmptyModel = c.getBool("saveEmptyModel", false);
    Matrix channel;  // NAP: the client supervector loses its projection on the channel subspace (:96-102, 154-157)
    const bool nap = c.existsParam("NAP");
    if (nap) {
      channel.load(c.getParam("NAP"), c.getString("loadMatrixFormat", "DB"));
      if (channel.cols != (size_t)world.C * world.D) LIA_THROW("Incorrect dimension of the NAP channel matrix");
    }
    for (auto &line : ids.lines()) {
      std::vector<std::string> files(line.begin() + 1, line.end());
      FeatureServer fs(c, files);
      SegCluster segs = selectedSegments(c, fs, label);
      MixtureGD client = world;  // the client starts from the world model
      if (initByClient) client = MixtureGD::loadFromConfig(line[0], c);
      client.id = line[0];
      if (segs.empty()) {
        std::cout << " WARNING - NO DATA FOR TRAINING [" << line[0] << "]";
        if (saveEmptyModel) {
          std::cout << " World model is returned" << std::endl;
          client.saveFromConfig(line[0], c);
        }
        continue;
      }
      adaptModel(c, fs, segs, world, client, mapCfg);
      if (nap) {  // computeNap (SuperVectors.cpp:128-138): v -= U'(U v) on the mean supervector
        std::vector<double> t(channel.rows, 0.0);
        for (size_t i = 0; i < channel.rows; i++)
          for (size_t k = 0; k < channel.cols; k++) t[i] += channel(i, k) * client.mean[k];
        for (size_t i = 0; i < channel.rows; i++)
          for (size_t k = 0; k < channel.cols; k++) client.mean[k] -= channel(i, k) * t[i];
      }
      client.saveFromConfig(line[0], c);
    }
  } catch (std::exception &e) {
    std::cout << e.what() << std::endl;
  }
  return 0;
}

// ------------------------------------------------------------------ ComputeTest
int ComputeTest(Config &c) {
  try {
    const std::string gender = c.getParam("gender");
    const std::string label = c.getParam("labelSelectedFrames");
    const bool segmental = c.getBool("segmentLLR", false);
    const double threshold = c.getDouble("decisionThreshold", 0.0);
    const double frameLength = c.getDouble("frameLength", 0.01);
    const int K = (int)c.getLong("topDistribsCount", 10);
    const bool complete = c.getString("computeLLKWithTopDistribs", "COMPLETE") == "COMPLETE";
    const double minLLK = c.getDouble("minLLK", -200.0), maxLLK = c.getDouble("maxLLK", 200.0);
    const long worldDecime = c.getLong("worldDecime", 1);  // ComputeTest.cpp:111-113
    if (worldDecime < 1) LIA_THROW("worldDecime must be >= 1");
    // windowLLR (WindowLLR, UnsupervisedTools.cpp:92-150; ComputeTest.cpp:100, 143, 165-178): a score per client for
    // every window of windowLLRSize selected frames, shifted by windowLLRDec, next to the file / segment scores
    const bool windowSet = c.getBool("windowLLR", false);
    const long windowSize = windowSet ? c.getLong("windowLLRSize", 30) : 0;
    const long windowDec = windowSet ? c.getLong("windowLLRDec", windowSize) : 0;
    if (windowSet && (windowSize < 1 || windowDec < 1 || windowDec > windowSize)) LIA_THROW("windowLLRSize / windowLLRDec out of range");
    if (windowSet && worldDecime != 1) LIA_THROW("windowLLR with worldDecime > 1 is not implemented by this engine");
    XList ndx(c.getParam("ndxFilename"));
    MixtureGD worldM = MixtureGD::loadFromConfig(c.getParam("inputWorldFilename"), c);
    Gmm world(worldM, true);
    std::map<std::string, std::unique_ptr<Gmm>> cache;  // client models stay resident (TabClientLine)
    // several ranks: contiguous ranges of NDX lines (independent test files, no collective); every rank
    // writes its part, rank 0 concatenates them in rank order -- the file a single process writes
    const Shard &sh = Shard::get();
    const std::string outName = c.getParam("outputFilename");
    auto partName = [&](int r) { return sh.world > 1 ? outName + ".part" + std::to_string(r) : outName; };
    std::ofstream outNist(partName(sh.rank).c_str(), std::ios::out | std::ios::trunc);
    const auto allLines = ndx.lines();
    const auto myRange = sh.range(allLines.size());
    for (size_t li = myRange.first; li < myRange.second; li++) {
      const auto &line = allLines[li];
      const std::string &test = line[0];
      FeatureServer fs(c, {test});
      SegCluster segs = selectedSegments(c, fs, label);
      if (segs.empty()) {
        std::cout << "ATTENTION, TEST FILE [" << test << "] is empty" << std::endl;
        continue;
      }
      std::vector<lr_gmm *> clients;
      for (size_t i = 1; i < line.size(); i++) {
        auto it = cache.find(line[i]);
        if (it == cache.end())
          it = cache.emplace(line[i], std::unique_ptr<Gmm>(new Gmm(MixtureGD::loadFromConfig(line[i], c), true))).first;
        clients.push_back(it->second->h());
      }
      std::vector<lr_seg> es = toEngineSegs(fs, segs);
      const size_t nOut = segmental ? es.size() : 1;
      std::vector<double> mw(nOut), mc(clients.size() * nOut);
      if (windowSet) {
        // per-frame world / client log-likelihoods of the selected frames (the same DETERMINE_TOP / USE_TOP entry
        // points), then the reference's loop replayed on them: window lines in frame order, the segment / file lines
        // where the reference writes them
        const size_t D = fs.ld(), nC = clients.size();
        size_t nSel = 0;
        for (const lr_seg &sg : es) nSel += (size_t)sg.length;
        std::vector<float> Xs(nSel * D);
        {
          size_t t = 0;
          for (const lr_seg &sg : es) {
            std::copy(fs.data() + (size_t)sg.begin * D, fs.data() + (size_t)(sg.begin + sg.length) * D, Xs.begin() + t * D);
            t += (size_t)sg.length;
          }
        }
        std::vector<double> llkw(nSel), rest(nSel), llkc(nC * nSel);
        std::vector<uint32_t> idx(nSel * (size_t)K);
        LIA_CHECK(lr_gmm_llk_topk(world.h(), Xs.data(), nSel, D, K, complete ? 1 : 0, minLLK, maxLLK, llkw.data(), idx.data(),
                                  nullptr, rest.data(), nullptr));
        for (size_t i = 0; i < nC; i++)
          LIA_CHECK(lr_gmm_llk_use_topk(clients[i], Xs.data(), nSel, D, K, idx.data(), rest.data(), complete ? 1 : 0, minLLK,
                                        maxLLK, &llkc[i * nSel]));
        // WindowLLR state (setNbClient resets it per NDX line)
        std::vector<unsigned long> idxA((size_t)windowSize, 0);
        std::vector<double> accLlr(nC, 0.0), llrM((size_t)windowSize * nC, 0.0);
        long bIdx = 0, count = 0;
        double sumW = 0.0, nAcc = 0.0;
        std::vector<double> sumC(nC, 0.0);
        size_t t = 0;
        for (size_t o = 0; o < es.size(); o++) {
          for (long f = 0; f < es[o].length; f++, t++) {
            const unsigned long frame = (unsigned long)(es[o].begin + f);
            sumW += llkw[t];
            nAcc += 1.0;
            if (count < windowSize) {  // dec(): the window is not full yet
              count++;
              idxA[(size_t)((bIdx + count - 1) % windowSize)] = frame;
            } else {  // full: drop windowDec frames from its head
              for (long w = 0; w < windowDec; w++) {
                for (size_t i = 0; i < nC; i++) accLlr[i] -= llrM[(size_t)bIdx * nC + i];
                bIdx = (bIdx + 1) % windowSize;
              }
              count -= windowDec - 1;
              idxA[(size_t)((bIdx + count - 1) % windowSize)] = frame;
            }
            for (size_t i = 0; i < nC; i++) {
              sumC[i] += llkc[i * nSel + t];
              const double llr = llkc[i * nSel + t] - llkw[t];
              llrM[(size_t)((bIdx + count - 1) % windowSize) * nC + i] = llr;  // accLLR()
              accLlr[i] += llr;
            }
            if (count == windowSize)
              for (size_t i = 0; i < nC; i++) {
                const double llr = accLlr[i] / (double)windowSize;
                outputResultLine(llr, line[i + 1], test, idxA[(size_t)bIdx] * frameLength,
                                 idxA[(size_t)((bIdx + count - 1) % windowSize)] * frameLength, gender,
                                 setDecision(llr, threshold), outNist);
              }
          }
          if (segmental) {
            for (size_t i = 0; i < nC; i++) {
              const double llr = sumC[i] / nAcc - sumW / nAcc;
              outputResultLine(llr, line[i + 1], test, es[o].begin * frameLength, (es[o].begin + es[o].length) * frameLength,
                               gender, setDecision(llr, threshold), outNist);
              sumC[i] = 0.0;
            }
            sumW = 0.0;
            nAcc = 0.0;
          }
        }
        if (!segmental)
          for (size_t i = 0; i < nC; i++) {
            const double llr = sumC[i] / nAcc - sumW / nAcc;
            outputResultLine(llr, line[i + 1], test, gender, setDecision(llr, threshold), outNist);
          }
        continue;
      }
      LIA_CHECK(lr_compute_test_decime(world.h(), clients.data(), (int)clients.size(), fs.data(),
                                       fs.getFeatureCount(), fs.ld(), es.data(), es.size(), K, complete ? 1 : 0,
                                       minLLK, maxLLK, segmental ? 1 : 0, (int)worldDecime, mw.data(), mc.data()));
      for (size_t o = 0; o < nOut; o++)
        for (size_t i = 0; i < clients.size(); i++) {
          double llr = mc[i * nOut + o] - mw[o];  // ComputeTest.cpp:196-199
          if (segmental)
            outputResultLine(llr, line[i + 1], test, es[o].begin * frameLength,
                             (es[o].begin + es[o].length) * frameLength, gender, setDecision(llr, threshold), outNist);
          else
            outputResultLine(llr, line[i + 1], test, gender, setDecision(llr, threshold), outNist);
        }
    }
    outNist.close();
    if (sh.world > 1) {
      sh.barrier();
      if (sh.rank == 0) {
        std::ofstream all(outName.c_str(), std::ios::out | std::ios::trunc);
        for (int r = 0; r < sh.world; r++) {
          std::ifstream part(partName(r).c_str());
          all << part.rdbuf();
          part.close();
          std::remove(partName(r).c_str());
        }
      }
    }
  } catch (std::exception &e) {
    std::cout << e.what() << std::endl;
  }
  return 0;
}

// ------------------------------------------------------------------ IvExtractor (classic mode)
int IvExtractor(Config &c) {
  try {
    // the first element of each targetIdList line is the model name (IvExtractor.cpp:92-99)
    XList ids(c.getParam("targetIdList"));
    std::vector<std::vector<std::string>> files;
    for (auto &l : ids.lines()) files.push_back(std::vector<std::string>(l.begin() + 1, l.end()));
    TVAcc tv(files, c);
    tv.loadT(c.getParam("totalVariabilityMatrix"), c);
    if (c.getBool("loadAccs", false)) {
      tv.loadN(c);
      tv.loadF_X(c);
    } else {
      tv.computeAndAccumulateTVStat(c);
      tv.saveAccs(c);
    }
    if (c.getBool("minDivergence", false)) {
      Matrix m;
      m.load(c.getString("matrixFilesPath", "") + c.getParam("meanEstimate") + c.getString("loadMatrixFilesExtension", ""),
             c.getString("loadMatrixFormat", "DB"));
      tv.loadMeanEstimate(m.data);
    }
    tv.substractM();
    tv.estimateTETt();
    tv.estimateW();
    tv.saveWbyFile(c);
  } catch (std::exception &e) {
    std::cout << e.what() << std::endl;
  }
  return 0;
}

// ------------------------------------------------------------------ ComputeJFAStats
// JFAAcc("Accumulate") + computeAndAccumulateJFAStat + saveAccs (ComputeJFAStats.cpp:71-87,
// AccumulateJFAStat.cpp:520-576, 4779-4800).  Every element of an NDX line is a session file of that
// line's speaker (JFATranslate, AccumulateJFAStat.h:99-111): speaker = line, session = running count; a
// file listed twice keeps the indices of its first occurrence (_idxOfID).
namespace {
// one pass over the files of a JFA NDX: per-session and per-speaker statistics, saved under the reference's
// names (saveAccs :4779-4800)
void jfaStatsToDisk(const Config &c) {
  XList ndx(c.getParam("ndxFilename"));
  std::vector<std::string> files;
  std::vector<int32_t> spkOfSession;
  std::map<std::string, int> sessionOfFile;
  int loc = 0;
  for (auto &l : ndx.lines()) {
    for (auto &f : l) {
      if (!sessionOfFile.count(f)) sessionOfFile[f] = (int)spkOfSession.size();
      files.push_back(f);
      spkOfSession.push_back(loc);
    }
    loc++;
  }
  const size_t nSessions = spkOfSession.size(), nSpeakers = (size_t)loc;
  if (nSessions == 0) LIA_THROW("JFA statistics: empty ndx");
  std::vector<std::string> unique;
  {
    std::set<std::string> seen;
    for (auto &f : files)
      if (seen.insert(f).second) unique.push_back(f);
  }
  MixtureGD world = MixtureGD::loadFromConfig(c.getParam("inputWorldFilename"), c);
  FeatureServer fs(c, unique);
  SegCluster sel = selectedSegments(c, fs, c.getParam("labelSelectedFrames"));
  std::vector<lr_seg> segs;
  for (const Seg &s : sel) {
    lr_seg e;
    e.begin = (int64_t)(fs.getFirstFeatureIndexOfASource(s.source) + s.begin);
    e.length = s.length;
    e.row = (int32_t)sessionOfFile[s.source];
    e.pad_ = 0;
    segs.push_back(e);
  }
  const size_t C = world.C, sv = (size_t)world.C * world.D;
  Matrix Nh(nSessions, C), Fh(nSessions, sv), N(nSpeakers, C), F(nSpeakers, sv);
  Gmm g(world, true);
  LIA_CHECK(lr_jfa_bwstats(g.h(), fs.data(), fs.getFeatureCount(), fs.ld(), segs.data(), segs.size(), nSessions,
                           spkOfSession.data(), nSpeakers, Nh.data.data(), Fh.data.data(), N.data.data(),
                           F.data.data()));
  const std::string path = c.getString("matrixFilesPath", ""), ext = c.getString("saveMatrixFilesExtension", "");
  const std::string fmt = c.getString("saveMatrixFormat", "DB");
  F.save(path + c.getString("firstOrderStatSpeaker", "F_X") + ext, fmt);
  Fh.save(path + c.getString("firstOrderStatSession", "F_X_h") + ext, fmt);
  Nh.save(path + c.getString("nullOrderStatSession", "N_h") + ext, fmt);
  N.save(path + c.getString("nullOrderStatSpeaker", "N") + ext, fmt);
}
}  // namespace

int ComputeJFAStats(Config &c) {
  try {
    // ComputeJFAStatsMain.cpp:108-117: computeStatMode JFA (default) | ivector (ComputeTVStats :89-103)
    const std::string mode = c.getString("computeStatMode", "JFA");
    if (mode == "ivector") {
      TVAcc tv(c.getParam("ndxFilename"), c);
      tv.computeAndAccumulateTVStat(c);
      tv.saveAccs(c);
      return 0;
    }
    if (mode != "JFA") LIA_THROW("computeStatMode must be JFA or ivector");
    jfaStatsToDisk(c);
  } catch (std::exception &e) {
    std::cout << e.what() << std::endl;
  }
  return 0;
}

// ------------------------------------------------------------------ EigenVoice
// EigenVoice.cpp:71-165.  With D, Z, U, X at their initial zeros the JFAAcc steps of this program are the
// TVAcc steps under other names -- estimateVEVT = estimateTETt (:1266-1293 vs AccumulateTVStat.cpp:777-805),
// estimateAndInverseL_EV + estimateYandV = estimateAandC (:2467-2515 vs :1702-1795), substractMplusDZ =
// substractM (Z = 0), substractUX = nothing (U = 0), updateVestimate = updateTestimate (:3597-3618),
// orthonormalizeV = orthonormalizeT, storeAccs / restoreAccs = the reload of N / F_X -- on the per-SPEAKER
// statistics, so the eigenvoice matrix V is trained by the same device path as T.
int EigenVoice(Config &c) {
  try {
    Config c2 = c;
    c2.setParam("totalVariabilityNumber", c.getParam("eigenVoiceNumber"));
    if (!c2.existsParam("nullOrderStatSpeaker")) c2.setParam("nullOrderStatSpeaker", "N");
    if (!c2.existsParam("firstOrderStatSpeaker")) c2.setParam("firstOrderStatSpeaker", "F_X");
    TVAcc tv(c2.getParam("ndxFilename"), c2);
    if (!c.getBool("loadAccs", false)) {
      if (Shard::get().world == 1) {
        jfaStatsToDisk(c2);  // N, F_X, N_h, F_X_h (computeAndAccumulateJFAStat + saveAccs)
        tv.loadN(c2);
        tv.loadF_X(c2);
      } else {
        tv.computeAndAccumulateTVStat(c2);  // the speaker-level pair, sharded by NDX line
        tv.saveAccs(c2);
      }
    } else {
      tv.loadN(c2);
      tv.loadF_X(c2);
    }
    if (c.getBool("loadInitEigenVoiceMatrix", false))
      tv.loadT(c.getParam("initEigenVoiceMatrix"), c2);
    else
      tv.initT(c2);
    const bool root = Shard::get().rank == 0;
    if (c.getBool("saveInitEigenVoiceMatrix", false) && root) tv.saveT(c.getParam("eigenVoiceMatrix") + "_init", c2);
    const long nbIt = c.getLong("nbIt");
    for (long it = 0; it < nbIt; it++) {
      std::cout << "\t(EigenVoices) --------- start iteration " << it << " --------" << std::endl;
      tv.estimateTETt();   // estimateVEVT
      tv.substractM();     // substractMplusDZ with Z = 0 (the order against estimateVEVT does not matter)
      tv.estimateAandC();  // estimateAndInverseL_EV + estimateYandV
      tv.updateTestimate();
      if (c.getBool("orthonormalizeV", false)) tv.orthonormalizeT();
      tv.resetTmpAcc();
      tv.reloadStats();  // restoreAccs
      if (c.getBool("saveAllEVMatrices", false) && root) tv.saveT(c.getParam("eigenVoiceMatrix") + std::to_string(it), c2);
    }
    if (root) tv.saveT(c.getParam("eigenVoiceMatrix"), c2);
  } catch (std::exception &e) {
    std::cout << e.what() << std::endl;
  }
  return 0;
}

// ------------------------------------------------------------------ EigenChannel (JFA mode)
// EigenChannelJFA (EigenChannel.cpp:71-160).  Speaker factors first (estimateVEVT, estimateAndInverseL_EV,
// substractMplusDZ, estimateY = the i-vector solve on the speaker statistics with V), then the channel subspace
// is trained on the SESSION statistics after removing every speaker's own supervector,
//   F_X_h[h] -= N_h[h] o (M + V y_spk(h))                    (substractMplusVYplusDZ :4400-4428, Z = 0)
// and from there estimateUEUT / estimateAndInverseL_EC / estimateXandU / updateUestimate (:3040-3088, :3620-3640)
// are the TVAcc steps on (N_h, F_X_h) with a zero mean -- the same device path as T and V.
namespace {
// What EigenChannelJFA and EstimateDMatrix share: the four statistics, the NDX structure, the speaker factors y
// (estimateVEVT, estimateAndInverseL_EV, substractMplusDZ with Z = 0, estimateY) and the speaker supervectors V y.
struct JfaState {
  Config cs;  // the configuration with the four statistics under explicit names
  std::string path, lext, sext, lfmt, sfmt;
  std::vector<int> spkOfSession;
  std::vector<std::vector<std::string>> sessionLines;
  size_t nSessions = 0, nSpk = 0, C = 0, sv = 0;
  MixtureGD world;
  Matrix N, F, Nh, Fh;  // speaker / session statistics
  Matrix VY;            // [nSpk x sv] V y (zero without an eigenvoice matrix)
};
Matrix loadSubspace(const JfaState &j, const std::string &name) {
  Matrix V;
  V.load(j.path + name + j.lext, j.lfmt);
  if (V.cols < V.rows) {  // stored transposed (loadEV / loadEC, like loadT :636-639)
    Matrix t(V.cols, V.rows);
    for (size_t i = 0; i < V.rows; i++)
      for (size_t k = 0; k < V.cols; k++) t(k, i) = V(i, k);
    V = t;
  }
  if (V.cols != j.sv) LIA_THROW("Incorrect dimension of the subspace matrix " + name);
  return V;
}
// rows[n x R] * M[R x sv] through the digit GEMM (the supervectors V y / U x)
Matrix supervectors(const Matrix &rows, const Matrix &M) {
  Matrix Mt(M.cols, M.rows);
  for (size_t i = 0; i < M.rows; i++)
    for (size_t k = 0; k < M.cols; k++) Mt(k, i) = M(i, k);
  Matrix out(rows.rows, M.cols);
  LIA_CHECK(lr_gemm_digits(rows.rows, M.cols, M.rows, rows.data.data(), Mt.data.data(), out.data.data(), 1.0, 0.0, 0));
  return out;
}
void jfaPrepare(Config &c, JfaState &j) {
  if (Shard::get().world != 1) LIA_THROW("JFA training programs: one process only (the session statistics are not sharded)");
  j.path = c.getString("matrixFilesPath", "");
  j.lext = c.getString("loadMatrixFilesExtension", "");
  j.sext = c.getString("saveMatrixFilesExtension", "");
  j.lfmt = c.getString("loadMatrixFormat", "DB");
  j.sfmt = c.getString("saveMatrixFormat", "DB");
  j.cs = c;
  if (!j.cs.existsParam("nullOrderStatSpeaker")) j.cs.setParam("nullOrderStatSpeaker", "N");
  if (!j.cs.existsParam("firstOrderStatSpeaker")) j.cs.setParam("firstOrderStatSpeaker", "F_X");
  if (!j.cs.existsParam("nullOrderStatSession")) j.cs.setParam("nullOrderStatSession", "N_h");
  if (!j.cs.existsParam("firstOrderStatSession")) j.cs.setParam("firstOrderStatSession", "F_X_h");
  if (!c.getBool("loadAccs", false)) jfaStatsToDisk(j.cs);
  XList ndx(c.getParam("ndxFilename"));
  int loc = 0;
  for (auto &l : ndx.lines()) {
    for (auto &f : l) {
      j.spkOfSession.push_back(loc);
      j.sessionLines.push_back({f});
    }
    loc++;
  }
  j.nSessions = j.spkOfSession.size();
  j.nSpk = (size_t)loc;
  j.world = MixtureGD::loadFromConfig(c.getParam("inputWorldFilename"), c);
  j.C = j.world.C;
  j.sv = (size_t)j.world.C * j.world.D;
  j.N.load(j.path + j.cs.getParam("nullOrderStatSpeaker") + j.lext, j.lfmt);
  j.F.load(j.path + j.cs.getParam("firstOrderStatSpeaker") + j.lext, j.lfmt);
  j.Nh.load(j.path + j.cs.getParam("nullOrderStatSession") + j.lext, j.lfmt);
  j.Fh.load(j.path + j.cs.getParam("firstOrderStatSession") + j.lext, j.lfmt);
  if (j.N.rows != j.nSpk || j.N.cols != j.C || j.F.rows != j.nSpk || j.F.cols != j.sv || j.Nh.rows != j.nSessions ||
      j.Nh.cols != j.C || j.Fh.rows != j.nSessions || j.Fh.cols != j.sv)
    LIA_THROW("Incorrect dimension of the JFA statistics");
  j.VY = Matrix(j.nSpk, j.sv);
  if (c.existsParam("eigenVoiceMatrix")) {
    Config cv = j.cs;
    cv.setParam("totalVariabilityNumber", c.getParam("eigenVoiceNumber"));
    TVAcc tvV(c.getParam("ndxFilename"), cv);
    tvV.loadN(cv);
    tvV.loadF_X(cv);
    tvV.loadT(c.getParam("eigenVoiceMatrix"), cv);
    tvV.substractM();
    tvV.estimateTETt();
    tvV.estimateW();
    j.VY = supervectors(tvV.getW(), loadSubspace(j, c.getParam("eigenVoiceMatrix")));
  }
}
// F_X_h[h] -= N_h[h] o (M + V y_spk(h))   (substractMplusVYplusDZ :4400-4428 with Z = 0)
Matrix centredSessionStats(const JfaState &j) {
  Matrix Fc = j.Fh;
  const int D = j.world.D;
  for (size_t h = 0; h < j.nSessions; h++) {
    const size_t sp = (size_t)j.spkOfSession[h];
    for (size_t k = 0; k < j.C; k++) {
      const double n = j.Nh(h, k);
      for (int i = 0; i < D; i++) {
        const size_t e = k * D + i;
        Fc(h, e) -= n * (j.world.mean[e] + j.VY(sp, e));
      }
    }
  }
  return Fc;
}
// a TVAcc over the sessions (one NDX line per session) holding (N_h, centred F_X_h) with a zero mean
std::unique_ptr<TVAcc> sessionAcc(const JfaState &j, const Matrix &Fc, const std::string &rankKeyValue, Config &cu) {
  Fc.save(j.path + "F_X_h_centered" + j.sext, j.sfmt);
  cu = j.cs;
  cu.setParam("totalVariabilityNumber", rankKeyValue);
  cu.setParam("nullOrderStatSpeaker", j.cs.getParam("nullOrderStatSession"));
  cu.setParam("firstOrderStatSpeaker", "F_X_h_centered");
  std::unique_ptr<TVAcc> tv(new TVAcc(j.sessionLines, cu));
  tv->loadN(cu);
  tv->loadF_X(cu);
  tv->loadMeanEstimate(std::vector<double>(j.sv, 0.0));  // everything was subtracted already
  return tv;
}
}  // namespace

int EigenChannel(Config &c) {
  try {
    JfaState j;
    jfaPrepare(c, j);
    Config cu;
    std::unique_ptr<TVAcc> tvU = sessionAcc(j, centredSessionStats(j), c.getParam("eigenChannelNumber"), cu);
    if (c.getBool("loadInitChannelMatrix", false))
      tvU->loadT(c.getParam("initEigenChannelMatrix"), cu);
    else
      tvU->initT(cu);
    if (c.getBool("saveInitChannelMatrix", false)) tvU->saveT(c.getParam("initEigenChannelMatrix"), cu);
    const long nbIt = c.getLong("nbIt");
    for (long it = 0; it < nbIt; it++) {
      std::cout << "\t(EigenChannel) --------- start iteration " << it << " --------" << std::endl;
      tvU->estimateTETt();   // estimateUEUT
      tvU->substractM();     // zero mean: keeps the row maxima of F for the digit GEMM up to date
      tvU->estimateAandC();  // estimateAndInverseL_EC + estimateXandU
      tvU->updateTestimate();
      tvU->resetTmpAcc();
      tvU->reloadStats();
      if (c.getBool("saveAllECMatrices", false)) tvU->saveT(c.getParam("eigenChannelMatrix") + std::to_string(it), cu);
    }
    tvU->saveT(c.getParam("eigenChannelMatrix"), cu);
  } catch (std::exception &e) {
    std::cout << e.what() << std::endl;
  }
  return 0;
}

// EigenChannel.cpp:178-290.  The LFA variant keeps a MAP-style diagonal term next to U: D = sqrt(Sigma / tau) (initD
// :1176-1222) and a speaker offset D z re-estimated at every iteration.  Per iteration, on the restored statistics:
//   F_X_h -= N_h o (M + V y + D z)            substractMplusVYplusDZ (:4400-4428), z from the previous iteration
//   x = L^-1 U' Sigma^-1 F_X_h                estimateUEUT, estimateAndInverseL_EC, estimateX (:3264-3295)
//   F_X  -= sum_{h of spk} N_h o (M + U x_h)  substractMplusUX (:4336-4362), on the SPEAKER statistics
//   z = tau / (tau + N) D Sigma^-1 F_X        estimateZMAP (:3576-3594), tau = regulationFactor read as an integer
//   A_k += (L^-1 + x x') N_hk, C += x F_X_h'  estimateU (:3424-3476), then updateUestimate
// x, A and C are one device E-step (estimateAandC computes the same x internally), the rest is element-wise.
int EigenChannelLFA(Config &c) {
  try {
    JfaState j;
    jfaPrepare(c, j);
    const int D = j.world.D;
    const std::string type = c.getString("initDType", "MAP");
    if (type != "MAP") LIA_THROW("initDType " + type + " is not implemented by this engine (MAP)");
    const double reg = c.getDouble("regulationFactor");
    const double tau = (double)c.getLong("regulationFactor");  // estimateZMAP takes regulationFactor.toLong()
    std::vector<double> Dm(j.sv);
    for (size_t e = 0; e < j.sv; e++) Dm[e] = std::sqrt(1.0 / (j.world.covinv[e] * reg));
    Matrix Z(j.nSpk, j.sv);
    Config cu = j.cs;
    cu.setParam("totalVariabilityNumber", c.getParam("eigenChannelNumber"));
    TVAcc tvU(j.sessionLines, cu);
    tvU.loadMeanEstimate(std::vector<double>(j.sv, 0.0));  // the statistics arrive centred
    if (c.getBool("loadInitChannelMatrix", false))
      tvU.loadT(c.getParam("initEigenChannelMatrix"), cu);
    else
      tvU.initT(cu);
    if (c.getBool("saveInitChannelMatrix", false)) tvU.saveT(c.getParam("initEigenChannelMatrix"), cu);
    const long nbIt = c.getLong("nbIt");
    Matrix Fc(j.nSessions, j.sv), Fs(j.nSpk, j.sv);
    for (long it = 0; it < nbIt; it++) {
      std::cout << "\t(EigenChannel) --------- start iteration " << it << " --------" << std::endl;
      for (size_t h = 0; h < j.nSessions; h++) {
        const size_t sp = (size_t)j.spkOfSession[h];
        for (size_t k = 0; k < j.C; k++)
          for (int i = 0; i < D; i++) {
            const size_t e = k * D + i;
            Fc(h, e) = j.Fh(h, e) - j.Nh(h, k) * (j.world.mean[e] + j.VY(sp, e) + Dm[e] * Z(sp, e));
          }
      }
      tvU.setStats(j.Nh, Fc);
      tvU.estimateTETt();
      tvU.substractM();
      tvU.estimateAandC();
      const Matrix UX = supervectors(tvU.getW(), tvU.getT());
      Fs = j.F;
      for (size_t h = 0; h < j.nSessions; h++) {
        const size_t sp = (size_t)j.spkOfSession[h];
        for (size_t k = 0; k < j.C; k++)
          for (int i = 0; i < D; i++) {
            const size_t e = k * D + i;
            Fs(sp, e) -= j.Nh(h, k) * (j.world.mean[e] + UX(h, e));
          }
      }
      for (size_t sp = 0; sp < j.nSpk; sp++)
        for (size_t k = 0; k < j.C; k++)
          for (int i = 0; i < D; i++) {
            const size_t e = k * D + i;
            Z(sp, e) = tau / (tau + j.N(sp, k)) * Dm[e] * j.world.covinv[e] * Fs(sp, e);
          }
      tvU.updateTestimate();
      tvU.resetTmpAcc();
      if (c.getBool("saveAllECMatrices", false)) tvU.saveT(c.getParam("eigenChannelMatrix") + std::to_string(it), cu);
    }
    tvU.saveT(c.getParam("eigenChannelMatrix"), cu);
  } catch (std::exception &e) {
    std::cout << e.what() << std::endl;
  }
  return 0;
}

// EigenChannelMain.cpp:139-143
int EigenChannelDispatch(Config &c) {
  const std::string mode = c.getString("eigenChannelMode", "JFA");
  if (mode == "JFA") return EigenChannel(c);
  if (mode == "LFA") return EigenChannelLFA(c);
  std::cout << "Error : wrong eigenChannelMode parameter, please chose JFA or LFA" << std::endl;
  return 1;
}

// ------------------------------------------------------------------ EstimateDMatrix
// EstimateDMatrix.cpp:99-210.  y with V on the speaker statistics, x with U on the session statistics (both the
// i-vector solve of the device path), then per iteration on F' = F_X - N o (M + V y) - sum_{h of spk} N_h o (U x_h)
// (substractMplusVY :3988-4007, substractUX :4152-4178) the diagonal update estimateZandD (:3480-3515):
//   L = 1 + N Sigma^-1 D^2,  z = F' Sigma^-1 D / L,  D <- sum_spk z F' / sum_spk (1 / L + z^2) N.
int EstimateDMatrix(Config &c) {
  try {
    JfaState j;
    jfaPrepare(c, j);
    const int D = j.world.D;
    // x per session with the eigenchannel matrix (U = 0 when none is given: x = 0)
    Matrix UX(j.nSessions, j.sv);
    if (c.existsParam("eigenChannelMatrix")) {
      Config cu;
      std::unique_ptr<TVAcc> tvU = sessionAcc(j, centredSessionStats(j), c.getParam("eigenChannelNumber"), cu);
      tvU->loadT(c.getParam("eigenChannelMatrix"), cu);
      tvU->substractM();
      tvU->estimateTETt();
      tvU->estimateW();
      UX = supervectors(tvU->getW(), loadSubspace(j, c.getParam("eigenChannelMatrix")));
    }
    // F' (the same at every iteration: restoreAccs, then the same two subtractions)
    Matrix Fp = j.F;
    for (size_t sp = 0; sp < j.nSpk; sp++)
      for (size_t k = 0; k < j.C; k++)
        for (int i = 0; i < D; i++) {
          const size_t e = k * D + i;
          Fp(sp, e) -= j.N(sp, k) * (j.world.mean[e] + j.VY(sp, e));
        }
    for (size_t h = 0; h < j.nSessions; h++) {
      const size_t sp = (size_t)j.spkOfSession[h];
      for (size_t k = 0; k < j.C; k++)
        for (int i = 0; i < D; i++) {
          const size_t e = k * D + i;
          Fp(sp, e) -= j.Nh(h, k) * UX(h, e);
        }
    }
    std::vector<double> Dm(j.sv);
    if (c.getBool("loadInitDMatrix", false)) {
      Matrix d0;
      d0.load(j.path + c.getParam("initDMatrix") + j.sext, j.lfmt);  // (loadD reads with the SAVE extension, :998)
      if (d0.rows != 1 || d0.cols != j.sv) LIA_THROW("Incorrect dimension of D Matrix");
      Dm = d0.data;
    } else {
      const std::string type = c.getString("initDType", "MAP");
      if (type != "MAP") LIA_THROW("initDType " + type + " is not implemented by this engine (MAP | loadInitDMatrix)");
      const double reg = c.getDouble("regulationFactor");
      for (size_t e = 0; e < j.sv; e++) Dm[e] = std::sqrt(1.0 / (j.world.covinv[e] * reg));  // initD :1212-1216
    }
    auto saveD = [&](const std::string &name) {
      Matrix d(1, j.sv);
      d.data = Dm;
      d.save(j.path + name + j.sext, j.sfmt);
    };
    if (c.getBool("saveInitD", false)) saveD(c.getParam("DMatrix") + "_init");
    const long nbIt = c.getLong("nbIt");
    std::vector<double> aux1(j.sv), aux2(j.sv);
    for (long it = 0; it < nbIt; it++) {
      std::cout << "\t(EstimateDMatrix) --------- start iteration " << it << " --------" << std::endl;
      std::fill(aux1.begin(), aux1.end(), 0.0);
      std::fill(aux2.begin(), aux2.end(), 0.0);
      for (size_t sp = 0; sp < j.nSpk; sp++)
        for (size_t k = 0; k < j.C; k++)
          for (int i = 0; i < D; i++) {
            const size_t e = k * D + i;
            const double n = j.N(sp, k), iv = j.world.covinv[e];
            const double L = 1.0 + n * iv * Dm[e] * Dm[e];
            const double z = Fp(sp, e) * iv * Dm[e] / L;
            aux1[e] += (1.0 / L + z * z) * n;
            aux2[e] += z * Fp(sp, e);
          }
      for (size_t e = 0; e < j.sv; e++) Dm[e] = aux2[e] / aux1[e];
      if (c.getBool("saveAllDMatrices", false)) saveD(c.getParam("DMatrix") + std::to_string(it));
    }
    saveD(c.getParam("DMatrix"));
  } catch (std::exception &e) {
    std::cout << e.what() << std::endl;
  }
  return 0;
}

// ------------------------------------------------------------------ TrainTarget, JFA
namespace {
// eigenvoice / eigenchannel matrix of the JFA programs: rank x sv, a single zero row when the key is absent
// (TrainTarget.cpp:437-462, ComputeTest.cpp:259-282)
Matrix jfaSubspace(const Config &c, const char *key, size_t sv) {
  Matrix M(1, sv);
  if (!c.existsParam(key)) return M;
  M.load(c.getString("matrixFilesPath", "") + c.getParam(key) + c.getString("loadMatrixFilesExtension", ""),
         c.getString("loadMatrixFormat", "DB"));
  if (M.cols < M.rows) {
    Matrix t(M.cols, M.rows);
    for (size_t i = 0; i < M.rows; i++)
      for (size_t k = 0; k < M.cols; k++) t(k, i) = M(i, k);
    M = t;
  }
  if (M.cols != sv) LIA_THROW(std::string("Incorrect dimension of the matrix given by ") + key);
  return M;
}
}  // namespace

// TrainTarget.cpp:393-617.  Per client (id + files, all files pooled into ONE set of speaker statistics): the joint
// factors [y; x] with the stacked matrix [V; U] (initVU :1227, estimateVUEVUT :1584-1608, estimateAndInverseL_VU
// :2300-2330, substractMplusDZ with z = 0, estimateYX :3518-3547) -- the i-vector solve with [V; U] in the place of T,
// batched over ALL clients of the list in one device pass -- then on the raw statistics minus N o (M + [V; U]'[y; x])
// (substractMplusVUYX :4364-4386) the diagonal factor z = F' Sigma^-1 D / (1 + N Sigma^-1 D^2) (estimateZ :3550-3572).
// Outputs: the client model M + V y + D z (saveMixture), the supervector Sigma^-1 (V y + D z) (saveSuperVector),
// optionally x / y / z.
namespace {
int trainTargetFactorAnalysis(Config &c, bool lfa) {
  try {
    if (Shard::get().world != 1) LIA_THROW("TrainTarget with channelCompensation JFA / LFA: one process only");
    if (c.getBool("useIdForSelectedFrame", false)) LIA_THROW("useIdForSelectedFrame is not implemented by this engine");
    XList ids(c.getParam("targetIdList"));
    const auto lines = ids.lines();
    if (lines.empty()) LIA_THROW("TrainTarget: empty targetIdList");
    MixtureGD world = MixtureGD::loadFromConfig(c.getParam("inputWorldFilename"), c);
    const int D = world.D;
    const size_t C = (size_t)world.C, sv = C * D;
    const Matrix V = jfaSubspace(c, "eigenVoiceMatrix", sv), U = jfaSubspace(c, "eigenChannelMatrix", sv);
    std::vector<double> Dm(sv, 0.0);
    // LFA (TrainTarget.cpp:664-668, 738-739): D = sqrt(Sigma / tau), z by estimateZMAP with tau read as an integer
    const double tau = lfa ? (double)c.getLong("regulationFactor") : 0.0;
    if (lfa) {
      const double reg = c.getDouble("regulationFactor");
      for (size_t e = 0; e < sv; e++) Dm[e] = std::sqrt(1.0 / (world.covinv[e] * reg));
    } else if (c.existsParam("DMatrix")) {
      Matrix d;
      d.load(c.getString("matrixFilesPath", "") + c.getParam("DMatrix") + c.getString("loadMatrixFilesExtension", ""),
             c.getString("loadMatrixFormat", "DB"));
      if (d.rows != 1 || d.cols != sv) LIA_THROW("Incorrect dimension of D Matrix");
      Dm = d.data;
    }
    const size_t Rv = V.rows, Ru = U.rows, R = Rv + Ru;
    Matrix VU(R, sv);
    std::copy(V.data.begin(), V.data.end(), VU.data.begin());
    std::copy(U.data.begin(), U.data.end(), VU.data.begin() + Rv * sv);
    std::vector<std::vector<std::string>> files;
    for (auto &l : lines) {
      if (l.size() < 2) LIA_THROW("TrainTarget: client [" + l[0] + "] has no feature file");
      files.emplace_back(l.begin() + 1, l.end());
    }
    Config ct = c;
    ct.setParam("totalVariabilityNumber", std::to_string(R));
    TVAcc tv(files, ct);
    tv.computeAndAccumulateTVStat(ct);
    const Matrix N = tv.getN(), F = tv.getF_X();
    tv.setT(VU);
    tv.substractM();
    tv.estimateTETt();
    tv.estimateW();
    const Matrix YX = tv.getW();  // [clients x (Rv + Ru)]
    Matrix Y(lines.size(), Rv), X(lines.size(), Ru);
    for (size_t s = 0; s < lines.size(); s++) {
      for (size_t i = 0; i < Rv; i++) Y(s, i) = YX(s, i);
      for (size_t i = 0; i < Ru; i++) X(s, i) = YX(s, Rv + i);
    }
    const Matrix VUYX = supervectors(YX, VU), VY = supervectors(Y, V);
    // (TrainTargetLFA saves the mixture only, :745-748)
    const bool saveMixture = lfa || c.getBool("saveMixture", true), saveSuperVector = !lfa && c.getBool("saveSuperVector", true);
    const bool saveEmpty = c.getBool("saveEmptyModel", false);
    const std::string sfmt = c.getString("saveMatrixFormat", "DB");
    auto saveRow = [&](const std::string &file, const double *v, size_t n) {
      Matrix m(1, n);
      std::copy(v, v + n, m.data.begin());
      m.save(file, sfmt);
    };
    std::vector<double> z(sv), sup(sv);
    for (size_t s = 0; s < lines.size(); s++) {
      const std::string &id = lines[s][0];
      double occ = 0.0;
      for (size_t k = 0; k < C; k++) occ += N(s, k);
      if (occ <= 0.0) {
        std::cout << " WARNING - NO DATA FOR TRAINING [" << id << "]";
        if (saveEmpty) {
          std::cout << " World model is returned" << std::endl;
          world.saveFromConfig(id, c);
        }
        continue;
      }
      MixtureGD client = world;
      client.id = id;
      for (size_t k = 0; k < C; k++)
        for (int i = 0; i < D; i++) {
          const size_t e = k * D + i;
          const double n = N(s, k), iv = world.covinv[e];
          const double fp = F(s, e) - n * (world.mean[e] + VUYX(s, e));
          z[e] = lfa ? tau / (tau + n) * Dm[e] * iv * fp                       // estimateZMAP :3576-3594
                     : fp * iv * Dm[e] / (1.0 + n * iv * Dm[e] * Dm[e]);  // estimateZ :3550-3572
          const double off = VY(s, e) + Dm[e] * z[e];  // getVYplusDZ :1817-1831
          sup[e] = off * iv;
          client.mean[e] = world.mean[e] + off;
        }
      if (saveMixture) client.saveFromConfig(id, c);
      if (saveSuperVector) saveRow(c.getParam("saveVectorFilesPath") + id + c.getParam("vectorFilesExtension"), sup.data(), sv);
      // (the reference builds these three names from a shadowed, empty path variable, :590-603: relative to the cwd)
      if (!lfa && c.getBool("saveX", false)) saveRow(id + c.getString("xExtension", ".x"), &X.data[s * Ru], Ru);
      if (!lfa && c.getBool("saveY", false)) saveRow(id + c.getString("yExtension", ".y"), &Y.data[s * Rv], Rv);
      if (!lfa && c.getBool("saveZ", false)) saveRow(id + c.getString("zExtension", ".z"), z.data(), sv);
    }
  } catch (std::exception &e) {
    std::cout << e.what() << std::endl;
  }
  return 0;
}

}  // namespace
int TrainTargetJFA(Config &c) { return trainTargetFactorAnalysis(c, false); }
int TrainTargetLFA(Config &c) { return trainTargetFactorAnalysis(c, true); }

// TrainTargetMain.cpp:160-171
int TrainTargetDispatch(Config &c) {
  if (c.existsParam("channelCompensation")) {
    const std::string cc = c.getParam("channelCompensation");
    if (cc == "JFA") return TrainTargetJFA(c);
    if (cc == "LFA") return TrainTargetLFA(c);
  }
  return TrainTarget(c);
}

// ------------------------------------------------------------------ ComputeTest, JFA channel compensation
namespace {
// What ComputeTestDotProduct (:228-370) and ComputeTestJFA (:376-572) do per NDX line before scoring, batched over
// the lines: Baum-Welch statistics of the test segment under the world model (computeAndAccumulateJFAStat),
// x = L^-1 U' Sigma^-1 (F - N o M) with L = I + sum_k N_k U_k' Sigma_k^-1 U_k (estimateUEUT, estimateAndInverseL_EC,
// substractMplusVYplusDZ with y = z = 0 for a test segment, estimateX :3264-3295) -- the i-vector solve with U in
// the place of T, one batched device pass for every test segment -- and the supervector offsets U x.
struct JfaTestSide {
  std::vector<std::vector<std::string>> lines;  // NDX lines: test segment + client ids
  MixtureGD world;
  Matrix N, F;  // raw statistics per line
  Matrix UX;    // [lines x sv] (zero without an eigenchannel matrix: x = 0)
};
void jfaTestSide(Config &c, JfaTestSide &j) {
  if (Shard::get().world != 1) LIA_THROW("ComputeTest with channelCompensation JFA: one process only");
  XList ndx(c.getParam("ndxFilename"));
  j.lines = ndx.lines();
  if (j.lines.empty()) LIA_THROW("ComputeTest: empty NDX list");
  j.world = MixtureGD::loadFromConfig(c.getParam("inputWorldFilename"), c);
  const size_t sv = (size_t)j.world.C * j.world.D;
  std::vector<std::vector<std::string>> tests;
  for (auto &l : j.lines) tests.push_back({l[0]});
  Config ct = c;
  long rank = 1;
  Matrix U;
  const bool haveU = c.existsParam("eigenChannelMatrix");
  if (haveU) {
    U.load(c.getString("matrixFilesPath", "") + c.getParam("eigenChannelMatrix") + c.getString("loadMatrixFilesExtension", ""),
           c.getString("loadMatrixFormat", "DB"));
    if (U.cols < U.rows) {  // loadEC transposes a matrix stored the other way round
      Matrix t(U.cols, U.rows);
      for (size_t i = 0; i < U.rows; i++)
        for (size_t k = 0; k < U.cols; k++) t(k, i) = U(i, k);
      U = t;
    }
    if (U.cols != sv) LIA_THROW("Incorrect dimension of the eigenchannel matrix");
    rank = (long)U.rows;
  }
  ct.setParam("totalVariabilityNumber", std::to_string(rank));
  TVAcc tv(tests, ct);
  tv.computeAndAccumulateTVStat(ct);
  j.N = tv.getN();
  j.F = tv.getF_X();
  j.UX = Matrix(j.lines.size(), sv);
  if (haveU) {
    tv.loadT(c.getParam("eigenChannelMatrix"), ct);
    tv.substractM();
    tv.estimateTETt();
    tv.estimateW();
    Matrix Ut(U.cols, U.rows);
    for (size_t i = 0; i < U.rows; i++)
      for (size_t k = 0; k < U.cols; k++) Ut(k, i) = U(i, k);
    Matrix X = tv.getW();
    LIA_CHECK(lr_gemm_digits(X.rows, sv, U.rows, X.data.data(), Ut.data.data(), j.UX.data.data(), 1.0, 0.0, 0));
  }
}
}  // namespace

int ComputeTestDotProduct(Config &c) {
  try {
    const std::string gender = c.getParam("gender");
    const double threshold = c.getDouble("decisionThreshold", 0.0);
    JfaTestSide j;
    jfaTestSide(c, j);
    const int D = j.world.D;
    const size_t C = (size_t)j.world.C, sv = C * D;
    std::ofstream outNist(c.getParam("outputFilename").c_str(), std::ios::out | std::ios::trunc);
    std::vector<double> fx(sv);
    for (size_t li = 0; li < j.lines.size(); li++) {
      // substractMplusUX (:4336-4362) on the raw statistics, then / sum N (:323-331)
      double sumN = 0.0;
      for (size_t k = 0; k < C; k++) sumN += j.N(li, k);
      for (size_t k = 0; k < C; k++)
        for (int i = 0; i < D; i++) {
          const size_t e = k * D + i;
          fx[e] = (j.F(li, e) - j.N(li, k) * (j.world.mean[e] + j.UX(li, e))) / sumN;
        }
      for (size_t i = 1; i < j.lines[li].size(); i++) {
        Matrix sup;
        sup.load(c.getParam("loadVectorFilesPath") + "/" + j.lines[li][i] + c.getParam("vectorFilesExtension"),
                 c.getString("loadMatrixFormat", "DB"));
        if (sup.data.size() < sv) LIA_THROW("client supervector " + j.lines[li][i] + " is shorter than the model");
        double score = 0.0;
        for (size_t e = 0; e < sv; e++) score += sup.data[e] * fx[e];
        outputResultLine(score, j.lines[li][i], j.lines[li][0], gender, setDecision(score, threshold), outNist);
      }
    }
    outNist.close();
  } catch (std::exception &e) {
    std::cout << e.what() << std::endl;
  }
  return 0;
}

namespace {
int computeTestFrameByFrame(Config &c, bool lfa) {
  try {
    const std::string gender = c.getParam("gender");
    const std::string label = c.getParam("labelSelectedFrames");
    const double threshold = c.getDouble("decisionThreshold", 0.0);
    const int K = (int)c.getLong("topDistribsCount", 10);
    const bool complete = c.getString("computeLLKWithTopDistribs", "COMPLETE") == "COMPLETE";
    const double minLLK = c.getDouble("minLLK", -200.0), maxLLK = c.getDouble("maxLLK", 200.0);
    const long worldDecime = c.getLong("worldDecime", 1);
    if (worldDecime < 1) LIA_THROW("worldDecime must be >= 1");
    JfaTestSide j;
    jfaTestSide(c, j);
    const int D = j.world.D;
    const size_t C = (size_t)j.world.C, sv = C * D;
    // LFA (ComputeTest.cpp:640-644, 660-667): D = sqrt(Sigma / tau); with x = z = 0 when they are formed, the channel
    // factor is the same x and z = tau / (tau + N) D Sigma^-1 (F - N o M) (estimateZMAP) joins the session model
    const double tau = lfa ? (double)c.getLong("regulationFactor") : 0.0;
    const double reg = lfa ? c.getDouble("regulationFactor") : 1.0;
    const bool doCms = lfa && c.getBool("cms", false);
    Gmm world(j.world, true);
    std::map<std::string, std::unique_ptr<Gmm>> cache;
    std::ofstream outNist(c.getParam("outputFilename").c_str(), std::ios::out | std::ios::trunc);
    for (size_t li = 0; li < j.lines.size(); li++) {
      const auto &line = j.lines[li];
      FeatureServer fs(c, {line[0]});
      SegCluster segs = selectedSegments(c, fs, label);
      if (segs.empty()) {
        std::cout << "ATTENTION, TEST FILE [" << line[0] << "] is empty" << std::endl;
        continue;
      }
      std::vector<lr_seg> es = toEngineSegs(fs, segs);
      // substractUXfromFeatures (:4689-4698): posteriors under M + U x (getSpeakerModel :4605-4620 with y = z = 0)
      MixtureGD session = j.world;
      for (size_t e = 0; e < sv; e++) session.mean[e] += j.UX(li, e);
      if (lfa)
        for (size_t k = 0; k < C; k++)
          for (int i = 0; i < D; i++) {
            const size_t e = k * D + i;
            const double iv = j.world.covinv[e], dm = std::sqrt(1.0 / (iv * reg));
            const double z = tau / (tau + j.N(li, k)) * dm * iv * (j.F(li, e) - j.N(li, k) * j.world.mean[e]);
            session.mean[e] += dm * z;
          }
      {
        Gmm sessionModel(session, true);
        LIA_CHECK(lr_jfa_normalize_features(sessionModel.h(), &j.UX.data[li * sv], fs.mutableData(), fs.getFeatureCount(),
                                            fs.ld(), es.data(), es.size()));
      }
      if (doCms) {  // cms() (GeneralTools.cpp:713-730): zero mean, unit deviation over the selected frames
        float *Xf = fs.mutableData();
        const size_t ld = fs.ld();
        std::vector<double> m(D, 0.0), m2(D, 0.0);
        double n = 0.0;
        for (const lr_seg &sg : es)
          for (int64_t t = sg.begin; t < sg.begin + sg.length; t++, n += 1.0)
            for (int i = 0; i < D; i++) {
              const double v = Xf[(size_t)t * ld + i];
              m[i] += v;
              m2[i] += v * v;
            }
        for (int i = 0; i < D; i++) {
          m[i] /= n;
          m2[i] = std::sqrt(m2[i] / n - m[i] * m[i]);
        }
        for (const lr_seg &sg : es)
          for (int64_t t = sg.begin; t < sg.begin + sg.length; t++)
            for (int i = 0; i < D; i++) Xf[(size_t)t * ld + i] = (float)(((double)Xf[(size_t)t * ld + i] - m[i]) / m2[i]);
      }
      std::vector<lr_gmm *> clients;
      for (size_t i = 1; i < line.size(); i++) {
        auto it = cache.find(line[i]);
        if (it == cache.end())
          it = cache.emplace(line[i], std::unique_ptr<Gmm>(new Gmm(MixtureGD::loadFromConfig(line[i], c), true))).first;
        clients.push_back(it->second->h());
      }
      std::vector<double> mw(1), mc(clients.size());
      LIA_CHECK(lr_compute_test_decime(world.h(), clients.data(), (int)clients.size(), fs.data(), fs.getFeatureCount(),
                                       fs.ld(), es.data(), es.size(), K, complete ? 1 : 0, minLLK, maxLLK, 0,
                                       (int)worldDecime, mw.data(), mc.data()));
      for (size_t i = 0; i < clients.size(); i++) {
        const double llr = mc[i] - mw[0];
        outputResultLine(llr, line[i + 1], line[0], gender, setDecision(llr, threshold), outNist);
      }
    }
    outNist.close();
  } catch (std::exception &e) {
    std::cout << e.what() << std::endl;
  }
  return 0;
}

}  // namespace
int ComputeTestJFA(Config &c) { return computeTestFrameByFrame(c, false); }
int ComputeTestLFA(Config &c) { return computeTestFrameByFrame(c, true); }

// ComputeTestMain.cpp:137-165
int ComputeTestDispatch(Config &c) {
  if (c.existsParam("byLabelModel") || c.existsParam("histoMode")) {
    std::cout << "(ComputeTest) byLabelModel / histoMode are not implemented by this engine" << std::endl;
    return 1;
  }
  if (c.existsParam("channelCompensation")) {
    const std::string cc = c.getParam("channelCompensation");
    if (cc == "JFA") return c.getString("scoring", "DotProduct") == "FrameByFrame" ? ComputeTestJFA(c) : ComputeTestDotProduct(c);
    if (cc == "LFA") return ComputeTestLFA(c);
    if (cc == "NAP") {
      std::cout << "(ComputeTest) channelCompensation NAP is not implemented by this engine" << std::endl;
      return 1;
    }
    std::cout << "(ComputeTest) No Channel Compensation" << std::endl;
  }
  return ComputeTest(c);
}

// ------------------------------------------------------------------ IvExtractor (approximate modes)
namespace {
// shared head / tail of IvExtractorUbmWeigth and IvExtractorEigenDecomposition
// (IvExtractor.cpp:151-252, 254-363)
std::vector<std::vector<std::string>> ivFileList(const Config &c) {
  XList ids(c.getParam("targetIdList"));
  std::vector<std::vector<std::string>> files;
  for (auto &l : ids.lines()) files.push_back(std::vector<std::string>(l.begin() + 1, l.end()));
  return files;
}
void ivStatsAndMean(TVAcc &tv, const Config &c) {
  if (c.getBool("loadAccs", false)) {
    tv.loadN(c);
    tv.loadF_X(c);
  } else {
    tv.computeAndAccumulateTVStat(c);
    tv.saveAccs(c);
  }
  if (c.getBool("minDivergence", false)) {
    Matrix m;
    m.load(c.getString("matrixFilesPath", "") + c.getParam("meanEstimate") + c.getString("loadMatrixFilesExtension", ""),
           c.getString("loadMatrixFormat", "DB"));
    tv.loadMeanEstimate(m.data);
  }
  tv.normStatistics();  // subtract the mean and normalise by the UBM co-variance (:239, :350)
}
std::string approxName(const Config &c, const char *suffix) {
  return c.getString("matrixFilesPath", "") + c.getParam("totalVariabilityMatrix") + suffix +
         c.getString("loadMatrixFilesExtension", "");
}
}  // namespace

int IvExtractorUbmWeigth(Config &c) {
  try {
    TVAcc tv(ivFileList(c), c);
    Matrix W;
    if (c.getBool("loadUbmWeightParam", false)) {
      tv.loadT(c.getParam("totalVariabilityMatrix") + "_norm", c);
      W.load(approxName(c, "_weightedCov"), c.getString("loadMatrixFormat", "DB"));
    } else {
      tv.loadT(c.getParam("totalVariabilityMatrix"), c);
      tv.normTMatrix();
      W = tv.getWeightedCov(tv.world().w);
    }
    ivStatsAndMean(tv, c);
    tv.estimateWUbmWeight(W);
    tv.saveWbyFile(c);
  } catch (std::exception &e) {
    std::cout << e.what() << std::endl;
  }
  return 0;
}

int IvExtractorEigenDecomposition(Config &c) {
  try {
    TVAcc tv(ivFileList(c), c);
    Matrix Q, D;
    if (c.getBool("loadEigenDecompositionParam", false)) {
      tv.loadT(c.getParam("totalVariabilityMatrix") + "_norm", c);
      D.load(approxName(c, "_EigDec_D"), c.getString("loadMatrixFormat", "DB"));
      Q.load(approxName(c, "_EigDec_Q"), c.getString("loadMatrixFormat", "DB"));
    } else {
      tv.loadT(c.getParam("totalVariabilityMatrix"), c);
      tv.normTMatrix();
      Matrix W = tv.getWeightedCov(tv.world().w);
      TVAcc::computeEigenProblem(W, Q, tv.rank());
      D = tv.approximateTcTc(Q);
    }
    ivStatsAndMean(tv, c);
    tv.estimateWEigenDecomposition(D, Q);
    tv.saveWbyFile(c);
  } catch (std::exception &e) {
    std::cout << e.what() << std::endl;
  }
  return 0;
}

// ------------------------------------------------------------------ TotalVariability
int TotalVariability(Config &c) {
  try {
    TVAcc tv(c.getParam("ndxFilename"), c);
    bool statsOnDisk = c.getBool("loadAccs", false);
    if (statsOnDisk) {
      tv.loadN(c);
      tv.loadF_X(c);
    } else {
      tv.computeAndAccumulateTVStat(c);
      tv.saveAccs(c);
    }
    if (c.getBool("loadInitTotalVariabilityMatrix", false))
      tv.loadT(c.getParam("initTotalVariabilityMatrix"), c);
    else
      tv.initT(c);
    const bool root = Shard::get().rank == 0;  // T, the mean estimate and the approximation matrices are replicated
    if (c.getBool("saveInitTotalVariabilityMatrix", false) && root) tv.saveT(c.getParam("totalVariabilityMatrix") + "_init", c);
    const bool minDiv = c.getBool("minDivergence", false);
    const long nbIt = c.getLong("nbIt");
    for (long it = 0; it < nbIt; it++) {
      std::cout << "\t(TotalVariability) --------- start iteration " << it << " --------" << std::endl;
      tv.substractM();
      tv.estimateTETt();
      tv.estimateAandC();
      tv.updateTestimate();
      if (minDiv) tv.minDivergence();
      if (c.getBool("orthonormalizeT", false)) tv.orthonormalizeT();
      tv.resetTmpAcc();
      tv.reloadStats();  // the reference re-reads N / F_X from disk here (:149-153); we keep a host copy
      if (c.getBool("saveAllTVMatrices", false) && root) tv.saveT(c.getParam("totalVariabilityMatrix") + std::to_string(it), c);
    }
    if (root) tv.saveT(c.getParam("totalVariabilityMatrix"), c);
    if (minDiv && root) {
      Matrix m = tv.getUbmMeans();
      m.save(c.getString("matrixFilesPath", "") + c.getParam("meanEstimate") + c.getString("saveMatrixFilesExtension", ""),
             c.getString("saveMatrixFormat", "DB"));
    }
    // parameters of the approximate i-vector extraction modes (TotalVariability.cpp:181-241)
    if (c.existsParam("approximationMode")) {
      const std::string mode = c.getParam("approximationMode");
      const std::string fmt = c.getString("saveMatrixFormat", "DB");
      if (mode == "ubmWeight" || mode == "eigenDecomposition") {
        tv.normTMatrix();
        if (root) tv.saveT(c.getParam("totalVariabilityMatrix") + "_norm", c);
        Matrix W = tv.getWeightedCov(tv.world().w);
        if (mode == "ubmWeight") {
          if (root) W.save(approxName(c, "_weightedCov"), fmt);
        } else {
          Matrix Q;
          TVAcc::computeEigenProblem(W, Q, tv.rank());
          Matrix D = tv.approximateTcTc(Q);
          if (root) D.save(approxName(c, "_EigDec_D"), fmt);
          if (root) Q.save(approxName(c, "_EigDec_Q"), fmt);
        }
      } else {
        std::cout << "\t(TotalVariability) This approximation mode does not exists" << std::endl;
      }
    }
  } catch (std::exception &e) {
    std::cout << e.what() << std::endl;
  }
  return 0;
}

// ------------------------------------------------------------------ IvTest
// cosine / mahalanobis / 2cov branches and PLDA training live in backend.cpp
void IvTestTrainPlda(Config &c);
void IvTestPldaScoring(Config &c);
void IvTestNonPlda(Config &c, const std::string &scoring);

int IvTest(Config &c) {
  try {
    const std::string scoring = c.getString("scoring", "plda");
    if (scoring == "cosine" || scoring == "mahalanobis" || scoring == "2cov") {
      IvTestNonPlda(c, scoring);
      return 0;
    }
    if (scoring != "plda") {
      std::cout << "Scoring option is invalid, must be: cosine OR mahalanobis OR 2cov OR plda" << std::endl;
      return 0;
    }
    // the model is trained first unless pldaLoadModel is set (IvTest.cpp:255-298; this mirror
    // defaults to loading, the reference requires the parameter)
    if (!c.getBool("pldaLoadModel", true)) IvTestTrainPlda(c);
    IvTestPldaScoring(c);
  } catch (std::exception &e) {
    std::cout << e.what() << std::endl;
  }
  return 0;
}

}  // namespace lia
