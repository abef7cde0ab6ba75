// backend.cpp -- i-vector back-end of the host mirror: PldaDev (development set statistics and
// normalisations), the test-side normalisation chain, the cosine / Mahalanobis / two-covariance
// branches of IvTest and the IvNorm program (LIA_SpkTools/src/PldaTools.cpp,
// LIA_SpkDet/IvTest/src/IvTest.cpp:112-391, LIA_SpkDet/IvNorm/src/IvNorm.cpp:72-128).
// All numerics go through the lr_iv_* entry points of the engine.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "lia_host.h"

namespace lia {

namespace {
Matrix loadVector(const std::string &path, const std::string &name, const Config &c) {
  Matrix v;
  v.load(path + "/" + name + c.getString("loadVectorFilesExtension", ".y"), c.getString("loadMatrixFormat", "DB"));
  return v;
}
}  // namespace

// ------------------------------------------------------------------ PldaDev
PldaDev::PldaDev(const std::string &ndxFilename, const Config &c) {
  XList list(ndxFilename);
  auto lines = list.lines();
  if (lines.empty()) LIA_THROW("PldaDev: empty development list " + ndxFilename);
  // sortByElementNumber("descend") (:278)
  std::stable_sort(lines.begin(), lines.end(),
                   [](const std::vector<std::string> &a, const std::vector<std::string> &b) { return a.size() > b.size(); });
  nSpk_ = lines.size();
  size_t n = 0;
  for (auto &l : lines) n += l.size();
  const std::string path = c.getParam("loadVectorFilesPath");
  const size_t d = loadVector(path, lines[0][0], c).cols;
  data_ = Matrix(d, n);
  class_.resize(n);
  size_t s = 0;
  for (size_t spk = 0; spk < lines.size(); spk++)
    for (auto &f : lines[spk]) {
      Matrix v = loadVector(path, f, c);
      if (v.rows != 1 || v.cols != d) LIA_THROW("Incorrect dimension of vector to load");
      class_[s] = (int32_t)spk;
      for (size_t k = 0; k < d; k++) data_(k, s) = v.data[k];
      s++;
    }
  computeAll();
}

void PldaDev::computeAll() {
  mean_.assign(data_.rows, 0.0);
  LIA_CHECK(lr_iv_cov_mat((int)data_.rows, data_.cols, data_.data.data(), class_.data(), nSpk_, mean_.data(),
                          nullptr, nullptr, nullptr, nullptr));
}
void PldaDev::lengthNorm() {
  Matrix out(data_.rows, data_.cols);
  LIA_CHECK(lr_iv_normalize((int)data_.rows, data_.cols, data_.data.data(), nullptr, nullptr, 0, 1, out.data.data()));
  data_ = out;
  computeAll();
}
void PldaDev::center(const std::vector<double> &mu) {
  if (mu.size() != data_.rows) LIA_THROW("PldaDev::center: mean dimension mismatch");
  Matrix out(data_.rows, data_.cols);
  LIA_CHECK(lr_iv_normalize((int)data_.rows, data_.cols, data_.data.data(), mu.data(), nullptr, 0, 0, out.data.data()));
  data_ = out;
  computeAll();
}
void PldaDev::rotateLeft(const Matrix &M) {
  if (M.cols != data_.rows) LIA_THROW("Rotation dimension mismatch !");
  Matrix out(M.rows, data_.cols);
  LIA_CHECK(lr_iv_normalize((int)data_.rows, data_.cols, data_.data.data(), nullptr, M.data.data(), (int)M.rows, 0,
                            out.data.data()));
  data_ = out;
  computeAll();
}
void PldaDev::computeCovMat(Matrix &Sigma, Matrix &W, Matrix &B) {
  const size_t d = data_.rows;
  Sigma = Matrix(d, d);
  W = Matrix(d, d);
  B = Matrix(d, d);
  LIA_CHECK(lr_iv_cov_mat((int)d, data_.cols, data_.data.data(), class_.data(), nSpk_, nullptr, nullptr,
                          Sigma.data.data(), W.data.data(), B.data.data()));
}
void PldaDev::computeWccnChol(Matrix &WCCN) {
  WCCN = Matrix(data_.rows, data_.rows);
  LIA_CHECK(lr_iv_wccn_chol((int)data_.rows, data_.cols, data_.data.data(), class_.data(), nSpk_, WCCN.data.data()));
}
void PldaDev::computeMahalanobis(Matrix &M) {
  M = Matrix(data_.rows, data_.rows);
  LIA_CHECK(lr_iv_mahalanobis_matrix((int)data_.rows, data_.cols, data_.data.data(), class_.data(), nSpk_,
                                     M.data.data()));
}
// computeScatterMatUnThreaded (PldaTools.cpp:1607-1640), restated AS WRITTEN: SB sums the unnormalised outer
// products of the centred speaker means; SW is ASSIGNED (not accumulated) per speaker, and its inner loop runs over
// the sessions 0 .. n_c - 1 of the WHOLE set for every speaker c -- so what reaches the eigenproblem is the scatter
// of the first n_last sessions around their own speakers' means, divided by n_last (n_last = session count of the
// last speaker).  Kept so that `ldaMode scatterMatrices` gives the reference's matrices; d x d host arithmetic on
// means the engine computed (lr_iv_cov_mat).
void PldaDev::computeScatterMat(Matrix &SB, Matrix &SW) {
  const size_t d = data_.rows, n = data_.cols;
  std::vector<double> mean(d), spk(d * nSpk_);
  LIA_CHECK(lr_iv_cov_mat((int)d, n, data_.data.data(), class_.data(), nSpk_, mean.data(), spk.data(), nullptr, nullptr,
                          nullptr));
  SB = Matrix(d, d);
  SW = Matrix(d, d);
  std::vector<double> cm(d);
  for (size_t cs = 0; cs < nSpk_; cs++) {
    for (size_t i = 0; i < d; i++) cm[i] = spk[i * nSpk_ + cs] - mean[i];
    for (size_t i = 0; i < d; i++)
      for (size_t j = 0; j < d; j++) SB(i, j) += cm[i] * cm[j];
  }
  size_t nLast = 0;
  for (size_t s2 = 0; s2 < n; s2++)
    if ((size_t)class_[s2] == nSpk_ - 1) nLast++;
  if (nLast == 0) LIA_THROW("computeScatterMat: the last speaker has no session");
  std::vector<double> xc(d);
  for (size_t s2 = 0; s2 < std::min(nLast, n); s2++) {
    for (size_t i = 0; i < d; i++) xc[i] = data_(i, s2) - spk[i * nSpk_ + (size_t)class_[s2]];
    for (size_t i = 0; i < d; i++)
      for (size_t j = 0; j < d; j++) SW(i, j) += xc[i] * xc[j];
  }
  for (double &v : SW.data) v /= (double)nLast;
}

void PldaDev::computeLDA(Matrix &ldaMat, long ldaRank, const Config &c) {
  Matrix Sigma, W, B;
  if (c.getString("ldaMode", "covariance") == "scatterMatrices")
    computeScatterMat(B, W);
  else
    computeCovMat(Sigma, W, B);
  ldaMat = Matrix((size_t)ldaRank, data_.rows);
  LIA_CHECK(lr_iv_lda((int)data_.rows, W.data.data(), B.data.data(), (int)ldaRank, ldaMat.data.data()));
}

static std::string efrName(const Config &c, const char *baseKey, const char *baseDefault, unsigned long it, bool forLoad) {
  const std::string mode = c.getString("ivNormEfrMode", "EFR");
  return c.getString("matrixFilesPath", "") + mode + "_" + c.getString(baseKey, baseDefault) + std::to_string(it) +
         c.getString(forLoad ? "loadMatrixFilesExtension" : "saveMatrixFilesExtension", "");
}
std::string efrMatrixFilename(const Config &c, unsigned long it, bool forLoad) {
  return efrName(c, "ivNormEfrMatrixBaseName", "ivNormEfrMatrix_it", it, forLoad);
}
std::string efrMeanFilename(const Config &c, unsigned long it, bool forLoad) {
  return efrName(c, "ivNormEfrMeanBaseName", "ivNormEfrMean_it", it, forLoad);
}

void PldaDev::sphericalNuisanceNormalization(const Config &c) {
  const unsigned long nbIt = (unsigned long)c.getLong("ivNormIterationNb", 1);
  const bool sph = c.getString("ivNormEfrMode", "EFR") == "sphNorm";
  const std::string fmt = c.getString("saveMatrixFormat", "DB");
  for (unsigned long it = 0; it < nbIt; it++) {
    Matrix Sigma, W, B;
    computeCovMat(Sigma, W, B);
    const size_t d = data_.rows;
    Matrix mat(d, d);
    LIA_CHECK(lr_iv_efr_matrix((int)d, (sph ? W : Sigma).data.data(), mat.data.data()));
    mat.save(efrMatrixFilename(c, it, false), fmt);
    Matrix mean(1, d);
    mean.data = mean_;
    mean.save(efrMeanFilename(c, it, false), fmt);
    // center, rotate, length-normalise in one engine call, then refresh the means (:1916-1926)
    Matrix out(d, data_.cols);
    LIA_CHECK(lr_iv_normalize((int)d, data_.cols, data_.data.data(), mean_.data(), mat.data.data(), (int)d, 1,
                              out.data.data()));
    data_ = out;
    computeAll();
  }
}

namespace {
// one applySphericalNuisanceNormalization pass over a [d x n] set (PldaDev :1931-1975, PldaTest :3793-3840)
void applyEfr(const Config &c, Matrix &X) {
  const unsigned long nbIt = (unsigned long)c.getLong("ivNormIterationNb", 1);
  const std::string fmt = c.getString("loadMatrixFormat", "DB");
  for (unsigned long it = 0; it < nbIt; it++) {
    Matrix mat, mean;
    mat.load(efrMatrixFilename(c, it, true), fmt);
    mean.load(efrMeanFilename(c, it, true), fmt);
    if (mat.cols != X.rows || mean.cols != X.rows) LIA_THROW("EFR parameters do not match the vector size");
    Matrix out(mat.rows, X.cols);
    LIA_CHECK(lr_iv_normalize((int)X.rows, X.cols, X.data.data(), mean.data.data(), mat.data.data(), (int)mat.rows, 1,
                              out.data.data()));
    X = out;
  }
}
void rotate(const Matrix &M, Matrix &X) {
  if (M.cols != X.rows) LIA_THROW("Rotation dimension mismatch !");
  Matrix out(M.rows, X.cols);
  LIA_CHECK(lr_iv_normalize((int)X.rows, X.cols, X.data.data(), nullptr, M.data.data(), (int)M.rows, 0, out.data.data()));
  X = out;
}

// PldaTest::load (:3437-3622): trials "segment model1 model2 ...", optional enrolment list
// "model session1 session2 ..."; or inputVectorFilename: one list, every vector against every vector
struct TestData {
  std::vector<std::string> modelIds, enrolSessions, segIds;
  std::vector<int32_t> modelOf;
  std::map<std::string, int> modelIndex, segIndex;
  std::vector<std::vector<std::string>> trialLines;
  Matrix models, segments;  // [d x enrol sessions], [d x segments]
  std::vector<uint8_t> trials;
  bool fromVectorList = false;

  explicit TestData(const Config &c) {
    std::string vpath;
    if (c.existsParam("inputVectorFilename")) {
      fromVectorList = true;
      vpath = c.getParam("loadVectorFilesPath");
      XList all(c.getParam("inputVectorFilename"));
      for (auto &e : all.allElements()) {
        modelIndex[e] = (int)modelIds.size();
        modelIds.push_back(e);
        enrolSessions.push_back(e);
        modelOf.push_back(modelIndex[e]);
        segIndex[e] = (int)segIds.size();
        segIds.push_back(e);
      }
    } else {
      vpath = c.getParam("testVectorFilesPath");
      XList tr(c.getParam("ndxFilename"));
      trialLines = tr.lines();
      if (c.existsParam("targetIdList") && !c.getParam("targetIdList").empty()) {
        XList enrol(c.getParam("targetIdList"));
        auto lines = enrol.lines();
        std::stable_sort(lines.begin(), lines.end(), [](const std::vector<std::string> &a,
                                                        const std::vector<std::string> &b) { return a.size() > b.size(); });
        for (auto &l : lines) {
          if (!modelIndex.count(l[0])) {
            modelIndex[l[0]] = (int)modelIds.size();
            modelIds.push_back(l[0]);
          }
          for (size_t e = 1; e < l.size(); e++) {
            enrolSessions.push_back(l[e]);
            modelOf.push_back(modelIndex[l[0]]);
          }
        }
      }
      for (auto &l : trialLines) {
        if (!segIndex.count(l[0])) {
          segIndex[l[0]] = (int)segIds.size();
          segIds.push_back(l[0]);
        }
        for (size_t e = 1; e < l.size(); e++)
          if (!modelIndex.count(l[e])) {  // a model without enrolment list is its own single session
            modelIndex[l[e]] = (int)modelIds.size();
            modelIds.push_back(l[e]);
            enrolSessions.push_back(l[e]);
            modelOf.push_back(modelIndex[l[e]]);
          }
      }
    }
    if (enrolSessions.empty() || segIds.empty()) LIA_THROW("PldaTest: no trial to score");
    const size_t d = loadVector(vpath, enrolSessions[0], c).cols;
    models = Matrix(d, enrolSessions.size());
    segments = Matrix(d, segIds.size());
    for (size_t j = 0; j < enrolSessions.size(); j++) {
      Matrix v = loadVector(vpath, enrolSessions[j], c);
      for (size_t i = 0; i < d; i++) models(i, j) = v.data[i];
    }
    for (size_t j = 0; j < segIds.size(); j++) {
      Matrix v = loadVector(vpath, segIds[j], c);
      for (size_t i = 0; i < d; i++) segments(i, j) = v.data[i];
    }
    trials.assign(modelIds.size() * segIds.size(), fromVectorList ? 1 : 0);
    for (auto &l : trialLines)
      for (size_t e = 1; e < l.size(); e++) trials[(size_t)modelIndex[l[e]] * segIds.size() + segIndex[l[0]]] = 1;
  }
  // test-side normalisation (IvTest.cpp:301-318, IvNorm.cpp:101-112)
  void normalize(const Config &c) {
    if (!c.getBool("ivNorm", false)) return;
    if (c.getLong("ivNormIterationNb", 1) > 0) {
      applyEfr(c, models);
      applyEfr(c, segments);
    }
    if (c.getBool("LDA", false)) {
      Matrix lda;
      lda.load(c.getString("matrixFilesPath", "") + c.getParam("ldaMatrix") + c.getString("loadMatrixFilesExtension", ""),
               c.getString("loadMatrixFormat", "DB"));
      rotate(lda, models);
      rotate(lda, segments);
    }
  }
};

void saveColumns(const Matrix &X, const std::vector<std::string> &names, const std::string &dir, const Config &c) {
  for (size_t j = 0; j < names.size(); j++) {
    Matrix v(1, X.rows);
    for (size_t i = 0; i < X.rows; i++) v(0, i) = X(i, j);
    v.save(dir + "/" + names[j] + c.getString("vectorFilesExtension", c.getString("loadVectorFilesExtension", ".y")),
           c.getString("saveMatrixFormat", "DB"));
  }
}

// development-side estimation shared by IvTest and IvNorm (IvTest.cpp:131-180, IvNorm.cpp:79-98)
void estimateNormalisation(PldaDev &dev, const Config &c) {
  if (!c.getBool("ivNormLoadParam", false)) {
    if (c.getLong("ivNormIterationNb", 1) > 0) dev.sphericalNuisanceNormalization(c);
    if (c.getBool("LDA", false)) {
      Matrix lda;
      dev.computeLDA(lda, c.getLong("ldaRank"), c);
      dev.rotateLeft(lda);
      lda.save(c.getString("matrixFilesPath", "") + c.getParam("ldaMatrix") + c.getString("loadMatrixFilesExtension", ""),
               c.getString("saveMatrixFormat", "DB"));
    }
  } else {
    if (c.getLong("ivNormIterationNb", 1) > 0) dev.applySphericalNuisanceNormalization(c);
    if (c.getBool("LDA", false)) {
      Matrix lda;
      lda.load(c.getString("matrixFilesPath", "") + c.getParam("ldaMatrix") + c.getString("loadMatrixFilesExtension", ""),
               c.getString("loadMatrixFormat", "DB"));
      dev.rotateLeft(lda);
    }
  }
}
}  // namespace

void PldaDev::applySphericalNuisanceNormalization(const Config &c) {
  applyEfr(c, data_);
  computeAll();
}

void writeIvTestScores(const Config &c, const Matrix &scores, const std::vector<uint8_t> &trials,
                       const std::vector<std::string> &modelIds, const std::vector<std::string> &segIds) {
  const size_t nm = modelIds.size(), nt = segIds.size();
  if (scores.rows != nm || scores.cols != nt || trials.size() != nm * nt) LIA_THROW("writeIvTestScores: dimension mismatch");
  const std::string out = c.getParam("outputFilename");
  const std::string format = c.getString("outputScoreFormat", "ascii");
  if (format == "ascii") {
    const std::string gender = c.getString("gender", "M");
    const double threshold = c.getDouble("decisionThreshold", 0.0);
    std::ofstream os(out.c_str(), std::ios::out | std::ios::trunc);
    for (size_t s = 0; s < nt; s++)
      for (size_t m = 0; m < nm; m++)
        if (trials[m * nt + s]) {
          const double v = scores(m, s);  // "gender client decision seg LLR" (IOFormat.cpp:112-120)
          os << gender << " " << modelIds[m] << " " << (v > threshold ? 1 : 0) << " " << segIds[s] << " " << v << std::endl;
        }
  } else if (format == "binary") {
    std::ofstream om((out + "_model.txt").c_str(), std::ios::out | std::ios::trunc);
    for (auto &m : modelIds) om << m << std::endl;
    std::ofstream osg((out + "_testSeg.txt").c_str(), std::ios::out | std::ios::trunc);
    for (auto &sg : segIds) osg << sg << std::endl;
    scores.save(out + c.getString("saveMatrixFilesExtension", ""), c.getString("saveMatrixFormat", "DB"));
  } else {
    LIA_THROW("outputScoreFormat must be ascii or binary, got " + format);
  }
}

// the cosine / mahalanobis / 2cov branches of IvTest (IvTest.cpp:112-391), output included
void IvTestNonPlda(Config &c, const std::string &scoring) {
  Matrix scores;
  const std::string mpath = c.getString("matrixFilesPath", "");
  const std::string lext = c.getString("loadMatrixFilesExtension", ""), sext = c.getString("saveMatrixFilesExtension", "");
  const std::string lfmt = c.getString("loadMatrixFormat", "DB"), sfmt = c.getString("saveMatrixFormat", "DB");
  const bool wccn = c.getBool("wccn", false);
  const bool computeWccn = wccn && c.existsParam("loadWccnMatrix") && !c.getBool("loadWccnMatrix", false);
  const bool loadMah = scoring == "mahalanobis" && c.getBool("loadMahalanobisMatrix", false);
  const bool loadWccn = scoring == "cosine" && wccn && c.getBool("loadWccnMatrix", false);
  const bool load2cov = scoring == "2cov" && c.getBool("load2covMatrix", false);
  const bool ivNorm = c.getBool("ivNorm", false);
  const std::string wccnFile = mpath + c.getString("wccnMatrix", "WCCN") + lext;
  std::string w2 = "2Cov_W", b2 = "2Cov_B";
  if (c.existsParam("TwoCovFilename")) {
    w2 = c.getParam("TwoCovFilename") + "_W";
    b2 = c.getParam("TwoCovFilename") + "_B";
  }
  auto mahFile = [&](const std::string &ext) {
    return c.existsParam("mahalanobisMatrix") ? mpath + c.getParam("mahalanobisMatrix") + ext : std::string("Mahalanobis");
  };

  TestData test(c);
  if ((ivNorm && !c.getBool("ivNormLoadParam", false)) || (scoring == "mahalanobis" && !loadMah) ||
      (wccn && !loadWccn) || (scoring == "2cov" && !load2cov)) {
    PldaDev dev(c.getParam("backgroundNdxFilename"), c);
    if (ivNorm) estimateNormalisation(dev, c);
    if (computeWccn) {
      Matrix W;
      dev.computeWccnChol(W);
      W.save(wccnFile, sfmt);
    }
    if (scoring == "mahalanobis") {
      Matrix M;
      dev.computeMahalanobis(M);
      M.save(mahFile(sext), sfmt);
    }
    if (scoring == "2cov") {
      Matrix Sigma, W, B;
      dev.computeCovMat(Sigma, W, B);
      W.save(mpath + w2 + sext, sfmt);
      B.save(mpath + b2 + sext, sfmt);
    }
  }
  test.normalize(c);
  const size_t d = test.models.rows, nm = test.modelIds.size(), nt = test.segIds.size();
  if (test.models.cols != nm)
    LIA_THROW("scoring " + scoring + " takes one enrolment vector per model (PldaTools.cpp:3842-3910 index _models by model)");
  scores = Matrix(nm, nt);
  if (scoring == "cosine") {
    if (wccn) {
      Matrix W;
      W.load(wccnFile, lfmt);
      rotate(W, test.models);
      rotate(W, test.segments);
    }
    LIA_CHECK(lr_iv_cosine_scoring((int)test.models.rows, nm, nt, test.models.data.data(), test.segments.data.data(),
                                   test.trials.data(), scores.data.data()));
  } else if (scoring == "mahalanobis") {
    Matrix M;
    M.load(mahFile(lext), lfmt);
    if (M.rows != d || M.cols != d) LIA_THROW("Mahalanobis matrix does not match the vector size");
    LIA_CHECK(lr_iv_mahalanobis_scoring((int)d, nm, nt, test.models.data.data(), test.segments.data.data(),
                                        M.data.data(), test.trials.data(), scores.data.data()));
  } else {
    Matrix W, B;
    W.load(mpath + w2 + lext, lfmt);
    B.load(mpath + b2 + lext, lfmt);
    if (W.rows != d || B.rows != d) LIA_THROW("two-covariance matrices do not match the vector size");
    LIA_CHECK(lr_iv_two_cov_scoring((int)d, nm, nt, test.models.data.data(), test.segments.data.data(), W.data.data(),
                                    B.data.data(), scores.data.data()));
  }
  writeIvTestScores(c, scores, test.trials, test.modelIds, test.segIds);
}

// ------------------------------------------------------------------ PldaModel (training)
PldaModel::PldaModel(const std::string &mode, const Config &c) : dev_(c.getParam("backgroundNdxFilename"), c) {
  if (mode != "train") LIA_THROW("PldaModel: only the training mode is a host object (scoring loads the matrices directly)");
  rankF_ = (size_t)c.getLong("pldaEigenVoiceNumber");
  rankG_ = (size_t)c.getLong("pldaEigenChannelNumber", 0);
  const size_t d = dev_.getVectSize();
  delta_.assign(d, 0.0);
  if (c.getBool("pldaLoadInitMatrices", false)) {  // :2075-2113
    const std::string path = c.getString("matrixFilesPath", ""), ext = c.getString("loadMatrixFilesExtension", "");
    const std::string fmt = c.getString("loadMatrixFormat", "DB");
    F_.load(path + c.getParam("pldaEigenVoiceMatrixInit") + ext, fmt);
    if (rankG_ > 0) G_.load(path + c.getParam("pldaEigenChannelMatrixInit") + ext, fmt);
    Sigma_.load(path + c.getParam("pldaSigmaMatrixInit") + ext, fmt);
    Matrix m;
    m.load(path + c.getParam("pldaMeanVecInit") + ext, fmt);
    originalMean_ = m.data;
    if (F_.rows != d || F_.cols != rankF_ || Sigma_.rows != d || Sigma_.cols != d || originalMean_.size() != d ||
        (rankG_ > 0 && (G_.rows != d || G_.cols != rankG_)))
      LIA_THROW("PldaModel: initial matrices do not match vectSize / ranks");
  } else {
    initModel(c);
  }
}

void PldaModel::initModel(const Config &c) {
  const size_t d = dev_.getVectSize();
  Matrix W, B;
  dev_.computeCovMat(Sigma_, W, B);  // Sigma is initialised from the total covariance (:2179-2186)
  const std::string law = c.getString("pldaRandomInitLaw", "normal");
  auto fill = [&](Matrix &M, size_t cols) {
    M = Matrix(d, cols);
    if (law == "normal") {
      // boxMullerGeneratorInit + boxMullerGenerator(0, 1) on libc rand() (ScoreWarp.cpp:68-79)
      double x1 = rand() / (float)RAND_MAX, x2;
      for (auto &v : M.data) {
        double val;
        do {
          x2 = x1;
          x1 = rand() / (float)RAND_MAX;
          val = std::sqrt(-2.0 * std::log(x1)) * std::cos(2.0 * 3.14159265358979323846 * x2);
        } while (std::isnan(val) || std::isinf(val));
        v = val;
      }
    } else if (law == "uniform") {
      for (auto &v : M.data) v = 2.0 * drand48() - 1.0;  // Eigen's Random(): uniform in [-1, 1]
    } else {
      LIA_THROW("Selected random initialization law does not exist");
    }
  };
  fill(F_, rankF_);
  fill(G_, rankG_);
  originalMean_ = dev_.getMean();
}

void PldaModel::updateModel(const Config &) {  // :2302-2326
  originalMean_ = dev_.getMean();
  delta_.assign(dev_.getVectSize(), 0.0);
}
void PldaModel::updateMean() { originalMean_ = dev_.getMean(); }
void PldaModel::centerData() { dev_.center(originalMean_); }

void PldaModel::em_iteration(const Config &, unsigned long) {
  Matrix X = dev_.getData();
  const size_t d = X.rows;
  if (F_.rows != d || Sigma_.rows != d) LIA_THROW("PldaModel: model dimension does not match the development data");
  LIA_CHECK(lr_plda_em_iteration((int)d, (int)rankF_, (int)rankG_, X.cols, X.data.data(), dev_.getClass().data(),
                                 dev_.getSpeakerNumber(), F_.data.data(), rankG_ ? G_.data.data() : nullptr,
                                 Sigma_.data.data(), delta_.data()));
  dev_.setData(X);  // _Dev.center(_Delta) changed the development data (:2333)
}

void PldaModel::saveModel(const Config &c) {
  const std::string path = c.getString("matrixFilesPath", ""), ext = c.getString("saveMatrixFilesExtension", "");
  const std::string fmt = c.getString("saveMatrixFormat", "DB");
  const size_t d = dev_.getVectSize();
  Matrix mean(d, 1), md(d, 1);  // column vectors like the reference (:2819-2823)
  mean.data = originalMean_;
  md.data = delta_;
  mean.save(path + c.getString("pldaMeanVec", "pldaMeanVec") + ext, fmt);
  F_.save(path + c.getString("pldaEigenVoiceMatrix", "pldaEigenVoiceMatrix") + ext, fmt);
  if (rankG_ > 0) G_.save(path + c.getString("pldaEigenChannelMatrix", "pldaEigenChannelMatrix") + ext, fmt);
  Sigma_.save(path + c.getString("pldaSigmaMatrix", "pldaSigmaMatrix") + ext, fmt);
  md.save(path + c.getString("pldaMinDivMean", "pldaMinDivMean") + ext, fmt);
}

// ------------------------------------------------------------------ PLDA (PLDA.cpp:73-101)
int PLDA(Config &c) {
  try {
    PldaModel plda("train", c);
    plda.updateMean();
    plda.centerData();
    const unsigned long nbIt = (unsigned long)c.getLong("pldaNbIt");
    for (unsigned long it = 0; it < nbIt; it++) plda.em_iteration(c, it);
    plda.saveModel(c);
  } catch (std::exception &e) {
    std::cout << e.what() << std::endl;
  }
  return 0;
}

// training branch of IvTest for scoring = plda (IvTest.cpp:255-298): optional normalisation of the
// development data, EM, saveModel; the scoring branch then loads what was saved
void IvTestTrainPlda(Config &c) {
  PldaModel plda("train", c);
  if (c.getBool("ivNorm", false) && !c.getBool("ivNormLoadParam", false)) {
    if (c.getLong("ivNormIterationNb", 1) > 0) plda.getDev().sphericalNuisanceNormalization(c);
    if (c.getBool("LDA", false)) {
      Matrix lda;
      plda.getDev().computeLDA(lda, c.getLong("ldaRank"), c);
      plda.getDev().rotateLeft(lda);
      lda.save(c.getString("matrixFilesPath", "") + c.getParam("ldaMatrix") + c.getString("loadMatrixFilesExtension", ""),
               c.getString("saveMatrixFormat", "DB"));
    }
  }
  plda.updateModel(c);
  plda.centerData();
  const unsigned long nbIt = (unsigned long)c.getLong("pldaNbIt");
  for (unsigned long it = 0; it < nbIt; it++) plda.em_iteration(c, it);
  plda.saveModel(c);
}


// scoring branch of IvTest for scoring = plda (IvTest.cpp:301-318, 392-410; PldaTools.cpp:4489-4519)
void IvTestPldaScoring(Config &c) {
  const std::string mpath = c.getString("matrixFilesPath", ""), mext = c.getString("loadMatrixFilesExtension", "");
  const std::string mfmt = c.getString("loadMatrixFormat", "DB");
  // trials, enrolment lists and vectors exactly as the other scorings load them (PldaTest::load,
  // PldaTools.cpp:3437-3622), then the test-side normalisation every scoring mode goes through
  // (IvTest.cpp:301-318: EFR / sphNorm, then LDA).  pldaNativeScoring itself never centres the data
  // (PldaTools.cpp:4489-4519); PldaTest::center is only reached through sphericalNuisanceNormalization.
  TestData test(c);
  test.normalize(c);
  const std::vector<std::string> &modelIds = test.modelIds, &segIds = test.segIds;
  const std::vector<int32_t> &modelOf = test.modelOf;
  const std::map<std::string, int> &modelIndex = test.modelIndex, &segIndex = test.segIndex;
  const Matrix &models = test.models, &segments = test.segments;
  const size_t d = models.rows;
  const size_t nEnrol = test.enrolSessions.size(), nTest = segIds.size(), nModels = modelIds.size();
  Matrix F, G, Sigma;
  F.load(mpath + c.getString("pldaEigenVoiceMatrix", "pldaEigenVoiceMatrix") + mext, mfmt);
  Sigma.load(mpath + c.getString("pldaSigmaMatrix", "pldaSigmaMatrix") + mext, mfmt);
  const int rG = (int)c.getLong("pldaEigenChannelNumber", 0);
  if (rG > 0) G.load(mpath + c.getString("pldaEigenChannelMatrix", "pldaEigenChannelMatrix") + mext, mfmt);
  // the model lives in the space of the NORMALISED vectors (after LDA: ldaRank dimensions)
  if (F.rows != d || Sigma.rows != d || Sigma.cols != d || (rG > 0 && G.rows != d))
    LIA_THROW("IvTest: PLDA model dimension (F " + std::to_string(F.rows) + " x " + std::to_string(F.cols) +
              ", Sigma " + std::to_string(Sigma.rows) + " x " + std::to_string(Sigma.cols) +
              ") does not match the (normalised) i-vector size " + std::to_string(d));
  Matrix scores(nModels, nTest);
  // several ranks: contiguous ranges of MODELS (rows of the score matrix; PldaTools.cpp:4302-4412 splits the
  // models over threads the same way), segments replicated, no collective in the scoring; the row blocks are
  // gathered so that rank 0 writes the file a single process writes
  const Shard &sh = Shard::get();
  const auto mr = sh.range(nModels);
  const size_t myModels = mr.second - mr.first;
  // enrolment columns of the rank's models (model_of is non-decreasing)
  size_t e0 = 0, e1 = nEnrol;
  if (sh.world > 1) {
    e0 = std::lower_bound(modelOf.begin(), modelOf.end(), (int32_t)mr.first) - modelOf.begin();
    e1 = std::lower_bound(modelOf.begin(), modelOf.end(), (int32_t)mr.second) - modelOf.begin();
  }
  Matrix mine(std::max<size_t>(myModels, 1), nTest);
  if (myModels > 0) {
    if (c.getString("pldaScoring", "native") == "enrollMean") {
      // PldaTest::pldaMeanScoring (PldaTools.cpp:4612-4709): each model is the MEAN of its enrolment
      // i-vectors scored as one session -- K_two = (2 FTJF + I)^-1 is exactly the native K_{L+1} at
      // L = 1, so this is the native scorer on one averaged column per model.
      Matrix avg(d, myModels);
      std::vector<double> cnt(myModels, 0.0);
      std::vector<int32_t> one(myModels);
      for (size_t j = e0; j < e1; j++) {
        cnt[modelOf[j] - mr.first] += 1.0;
        for (size_t i = 0; i < d; i++) avg(i, modelOf[j] - mr.first) += models(i, j);
      }
      for (size_t m = 0; m < myModels; m++) {
        one[m] = (int32_t)m;
        for (size_t i = 0; i < d; i++) avg(i, m) /= cnt[m];
      }
      LIA_CHECK(lr_plda_native_scoring((int)d, (int)F.cols, rG, F.data.data(), rG ? G.data.data() : nullptr,
                                       Sigma.data.data(), avg.data.data(), myModels, one.data(), myModels,
                                       segments.data.data(), nTest, mine.data.data()));
    } else {
      // the rank's enrolment columns as a compact [d x (e1 - e0)] block
      Matrix part(d, e1 - e0);
      std::vector<int32_t> partOf(e1 - e0);
      for (size_t j = e0; j < e1; j++) {
        partOf[j - e0] = modelOf[j] - (int32_t)mr.first;
        for (size_t i = 0; i < d; i++) part(i, j - e0) = models(i, j);
      }
      LIA_CHECK(lr_plda_native_scoring((int)d, (int)F.cols, rG, F.data.data(), rG ? G.data.data() : nullptr,
                                       Sigma.data.data(), part.data.data(), e1 - e0, partOf.data(), myModels,
                                       segments.data.data(), nTest, mine.data.data()));
    }
  }
  if (sh.world == 1) {
    scores = mine;
  } else {
    const size_t blockRows = (nModels + sh.world - 1) / sh.world;
    std::vector<double> send(blockRows * nTest, 0.0), recv(send.size() * sh.world);
    if (myModels > 0) std::copy(mine.data.begin(), mine.data.begin() + myModels * nTest, send.begin());
    LIA_CHECK(lr_allgather_host(send.data(), send.size(), recv.data()));
    Shard probe = sh;
    for (int r = 0; r < sh.world; r++) {
      probe.rank = r;
      auto rg = probe.range(nModels);
      std::copy(recv.begin() + (size_t)r * send.size(), recv.begin() + (size_t)r * send.size() + (rg.second - rg.first) * nTest,
                scores.data.begin() + rg.first * nTest);
    }
    if (sh.rank != 0) return;
  }
  // output (IvTest.cpp:412-465): the trials listed in the NDX, segment-major in matrix order
  (void)modelIndex;
  (void)segIndex;
  writeIvTestScores(c, scores, test.trials, modelIds, segIds);
}

// ------------------------------------------------------------------ IvNorm (IvNorm.cpp:72-128)
int IvNorm(Config &c) {
  try {
    if (!c.getBool("ivNormLoadParam", false)) {
      PldaDev dev(c.getString("backgroundNdxFilename", ""), c);
      if (c.getLong("ivNormIterationNb", 1) > 0) dev.sphericalNuisanceNormalization(c);
      if (c.getBool("LDA", false)) {
        Matrix lda;
        dev.computeLDA(lda, c.getLong("ldaRank"), c);
        lda.save(c.getString("matrixFilesPath", "") + c.getParam("ldaMatrix") + c.getString("loadMatrixFilesExtension", ""),
                 c.getString("saveMatrixFormat", "DB"));
      }
    }
    if (c.existsParam("inputVectorFilename") || c.existsParam("ndxFilename")) {
      TestData test(c);
      Config apply = c;
      apply.setParam("ivNorm", "true");
      test.normalize(apply);
      const std::string out = c.getParam("saveVectorFilesPath");
      saveColumns(test.segments, test.segIds, out, c);  // saveSegments (:4712-4731)
      if (!(c.existsParam("inputVectorFilename") && !c.existsParam("targetIdList")))
        saveColumns(test.models, test.enrolSessions, out, c);  // saveVectors (:4734-)
    }
  } catch (std::exception &e) {
    std::cout << e.what() << std::endl;
  }
  return 0;
}

}  // namespace lia
