// lia_host.h -- C++ host mirror of the LIA_RAL / ALIZE surface that the hot path touches
// (SURVEY.md Appendix B), written against the C ABI of liblia_ral_b200.so.
//
// This is NOT alize-core: it is the minimum object surface the five hot programs use (Config,
// XList, MixtureGD + RAW/XML files, FeatureServer over SPRO3/SPRO4/RAW/HTK float32 files (bigEndian honoured) with
// featureServerMask, label -> segment selection, Matrix DB/DT files), plus the LIA_SpkTools
// batch functions re-expressed over the engine: accumulateStatEM, trainModel, TVAcc,
// ComputeTest's frame loop, PLDA native scoring.  Names, parameter keys and error behaviour
// follow the reference (file:line cited at each item); all numerics run on the GPU.
#pragma once

#include <cstdint>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/lia_ral_b200.h"

namespace lia {

// alize::Exception(msg, __FILE__, __LINE__) convention (AccumulateTVStat.cpp:481)
class Exception : public std::runtime_error {
 public:
  Exception(const std::string &msg, const char *file, int line);
  std::string toString() const { return what(); }
};
#define LIA_THROW(msg) throw ::lia::Exception((msg), __FILE__, __LINE__)
// rethrows lr_last_error() when an engine call fails
void check(lr_status st, const char *file, int line);
#define LIA_CHECK(expr) ::lia::check((expr), __FILE__, __LINE__)

// ---- Config: flat name -> value map; "name<ws>value" text files (banner lines "***" ignored),
// overridden by "--name value" from the command line (TrainWorldMain.cpp:90-104).
class Config {
 public:
  Config() {}
  explicit Config(const std::string &file) { load(file); }
  void load(const std::string &file);
  void parseCmdLine(int argc, char **argv);  // handles --config <file> first, then overrides
  bool existsParam(const std::string &n) const { return kv_.count(n) != 0; }
  void setParam(const std::string &n, const std::string &v) { kv_[n] = v; }
  const std::string &getParam(const std::string &n) const;  // throws when missing
  std::string getString(const std::string &n, const std::string &def) const;
  long getLong(const std::string &n) const;
  long getLong(const std::string &n, long def) const;
  double getDouble(const std::string &n) const;
  double getDouble(const std::string &n, double def) const;
  bool getBool(const std::string &n, bool def) const;

 private:
  std::map<std::string, std::string> kv_;
};

// ---- XList: whitespace separated tokens per line (NDX files, lists)
class XList {
 public:
  XList() {}
  explicit XList(const std::string &file) { load(file); }
  void load(const std::string &file);
  const std::vector<std::vector<std::string>> &lines() const { return lines_; }
  std::vector<std::string> allElements() const;
  std::vector<std::string> allUniqueElements() const;

 private:
  std::vector<std::vector<std::string>> lines_;
};

// ---- Matrix<double>: row-major; DB = uint32 rows, uint32 cols, double[]; DT = "rows cols\n" text
struct Matrix {
  size_t rows = 0, cols = 0;
  std::vector<double> data;
  Matrix() {}
  Matrix(size_t r, size_t c) : rows(r), cols(c), data(r * c, 0.0) {}
  double &operator()(size_t i, size_t j) { return data[i * cols + j]; }
  double operator()(size_t i, size_t j) const { return data[i * cols + j]; }
  void load(const std::string &file, const std::string &format /*DB|DT*/);
  void save(const std::string &file, const std::string &format) const;
};

// ---- MixtureGD / DistribGD: diagonal GMM with the alize RAW and XML file formats
struct MixtureGD {
  std::string id;
  int C = 0, D = 0;
  std::vector<double> w, mean, cov, covinv, cst, det;
  void resize(int c, int d);
  void computeAll();  // covInv, det, cst (DistribGD::computeAll)
  void load(const std::string &file, const std::string &format /*RAW|XML*/);
  void save(const std::string &file, const std::string &format) const;
  // mixtureFilesPath + name + loadMixtureFileExtension, format from the config
  static MixtureGD loadFromConfig(const std::string &name, const Config &c);
  void saveFromConfig(const std::string &name, const Config &c) const;
};

// ---- segments / labels
struct Seg {
  std::string source;
  long begin = 0, length = 0;  // frames, relative to the source file
  std::string label;
};
typedef std::vector<Seg> SegCluster;
long timeToFrameIdx(double t, double frameLength);  // SegTools.cpp:135-142
long totalFrame(const SegCluster &c);

// ---- FeatureServer: featureServerBufferSize ALL_FEATURES semantics -- every listed file is read
// into one float32 [frames x vectSize] block (after featureServerMask), sources indexed by name.
class FeatureServer {
 public:
  FeatureServer(const Config &c, const std::vector<std::string> &files);
  int getVectSize() const { return D_; }
  size_t getFeatureCount() const { return D_ ? X_.size() / D_ : 0; }
  size_t getSourceCount() const { return names_.size(); }
  const std::string &getNameOfASource(size_t i) const { return names_[i]; }
  size_t getFirstFeatureIndexOfASource(const std::string &name) const;
  size_t getFeatureCountOfASource(const std::string &name) const;
  const float *data() const { return X_.data(); }
  float *mutableData() { return X_.data(); }  // writeFeature: JFA feature compensation works in place
  size_t ld() const { return (size_t)D_; }

 private:
  int D_ = 0;
  std::vector<float> X_;
  std::vector<std::string> names_;
  std::vector<size_t> first_, count_;
};
std::vector<int> parseMask(const std::string &mask);  // "0-15,17-32"
// frames of label `labelSelectedFrames` for every source of fs (label files
// labelFilesPath + source + labelFilesExtension; addDefaultLabel / defaultLabel honoured);
// verifyClusterFile semantics: segments are clipped to the file length (SegTools.cpp:151-169)
SegCluster selectedSegments(const Config &c, const FeatureServer &fs, const std::string &label);
// segments -> engine segments (absolute frame index in the FeatureServer block), one row each
std::vector<lr_seg> toEngineSegs(const FeatureServer &fs, const SegCluster &segs, int row = 0);

// ---- engine RAII
class Gmm {
 public:
  explicit Gmm(const MixtureGD &m, bool use_file_cst = false);
  ~Gmm();
  Gmm(const Gmm &) = delete;
  Gmm &operator=(const Gmm &) = delete;
  lr_gmm *h() const { return h_; }
  void set(const MixtureGD &m);
  void get(MixtureGD &m) const;  // parameters + computeAll products back from the device

 private:
  lr_gmm *h_ = nullptr;
};

// ---- TopGauss (TopGauss.cpp:68-200): for every selected frame of a file, the indices of the retained top components
// (topGauss >= 1: that many; < 1: as many as it takes to exceed that share of the frame likelihood) and the weight /
// likelihood mass outside them.  File (nbGaussianFilesDir + name, native unsigned long = u64):
//   nt, nbgcnt, nbg[nt], idx[nbgcnt], snsw[nt], snsl[nt]
class TopGauss {
 public:
  double compute(const MixtureGD &ubm, const FeatureServer &fs, const std::string &file, const Config &c);  // mean LLK
  void write(const std::string &file, const Config &c) const;
  void read(const std::string &file, const Config &c);
  uint64_t nt = 0, nbgcnt = 0;
  std::vector<uint64_t> nbg, idx;
  std::vector<double> snsw, snsl;
};

// ---- TrainTools
struct TrainCfg {  // TrainTools.cpp:67-93
  double initVarianceFlooring, initVarianceCeiling, finalVarianceFlooring, finalVarianceCeiling;
  long nbTrainIt;
  double baggedFrameProbability;
  bool normalizeModel = false, normalizeModelMeanOnly = false;  // :76-86
  long normalizeModelNbIt = 1;
  bool componentReduction = false;                                // :87-92
  long targetDistribCount = 0;
  explicit TrainCfg(const Config &c);
};
// normalizeMixture (TrainTools.cpp:287-315) towards N(0, 1): every component is expressed relative to the
// moment-matched single Gaussian of the mixture (mixtureFusion :273-283); meanOnly repeats nbIt times
void normalizeMixture(MixtureGD &m, long nbIt, bool meanOnly);
// selectComponent(nbTop) + reduceModel + normalizeWeights (:197-229): the nbTop heaviest components, in index order
void reduceToTopWeights(MixtureGD &m, size_t nbTop);
double setItParameter(double begin, double end, int nbIt, int it);  // TrainTools.cpp:560-564
// bagging of 3..7-frame chunks with libc rand() exactly like GeneralTools.cpp:309-313, 455-500
SegCluster baggedSegments(const SegCluster &in, double p, long minLen, long maxLen);
// E-step over a cluster: returns sum log-likelihood; accumulates occ / m1 / m2 / n
struct EmAcc {
  std::vector<double> occ, m1, m2;
  double n = 0;
  void reset(int C, int D);
};
double accumulateStatEM(const FeatureServer &fs, const Gmm &g, const SegCluster &segs, EmAcc &acc,
                        double weight = 1.0);  // AccumulateStat.cpp:131
// one input stream of TrainWorld (inputStreamList / weightStreamList, TrainWorld.cpp:120-141): its
// feature server, its selected segments and its weight in the final model (default 1 / nbStream)
struct TrainStream {
  const FeatureServer *fs;
  SegCluster segs;
  double weight;
};
void computeMeanCov(const std::vector<TrainStream> &streams, std::vector<double> &mean, std::vector<double> &cov);
void mixtureInit(const std::vector<TrainStream> &streams, const std::vector<double> &globalCov, const Config &c,
                 MixtureGD &world);
void trainModel(const Config &c, const std::vector<TrainStream> &streams, const std::vector<double> &globalCov,
                MixtureGD &world, const TrainCfg &cfg);  // trainModelStream, TrainTools.cpp:1030-1110
void computeMeanCov(const FeatureServer &fs, const SegCluster &segs, std::vector<double> &mean,
                    std::vector<double> &cov);  // TrainTools.cpp:593-602
void mixtureInit(const FeatureServer &fs, const SegCluster &segs, const std::vector<double> &globalCov,
                 const Config &c, MixtureGD &world);  // TrainTools.cpp:674-766
void trainModel(const Config &c, const FeatureServer &fs, const SegCluster &segs,
                const std::vector<double> &globalCov, MixtureGD &world, const TrainCfg &cfg);

// ---- one process per GPU.  Every program accepts --lrWorldSize N --lrRank r --lrCommFile <path on a
// filesystem all ranks see> (or the environment: LR_WORLD_SIZE / LR_RANK / LR_COMM_FILE); the rank's device is
// --device (default: rank % device count).  The data shard the way the reference's threaded variants split
// them: contiguous NDX-line ranges (AccumulateTVStat.cpp:498-507), segment ranges balanced by frames for the
// EM accumulation (AccumulateStat.cpp:183-208 hands segments to threads), model rows for PLDA scoring
// (PldaTools.cpp:4302-4412).
struct Shard {
  int rank = 0, world = 1;
  static Shard &get();
  // [begin, end) of n units for this rank: contiguous, balanced
  std::pair<size_t, size_t> range(size_t n) const;
  // contiguous ranges balanced by weight (frames per unit)
  std::pair<size_t, size_t> rangeByWeight(const std::vector<double> &w) const;
  void barrier() const;
};
// lr_init on the rank's device + the communicator; called by every main before the driver
void initEngine(const Config &c);

// ---- MAP adaptation (TrainTools.cpp:110-147, 445-489, 871-904): the "next" row TrainTarget
struct MAPCfg {
  bool mean = false, var = false, weight = false;
  std::string method;  // MAPOccDep | MAPModelBased (regulation factors), MAPConst | MAPConst2 (a priori weights)
  double r[3] = {0, 0, 0};
  long nbTrainIt = 1;
  double baggedFrameProbability = 1.0;
  bool normalizeModel = false, normalizeModelMeanOnly = false;  // TrainTools.cpp:129-139
  long normalizeModelNbIt = 1;
  explicit MAPCfg(const Config &c);
};
// client holds the ML (EM) estimate on entry, the MAP estimate on return (computeMAPOccDep)
void computeMAPOccDep(const MixtureGD &initModel, MixtureGD &client, const MAPCfg &cfg, double frameCount);
void computeMAPConst(const MixtureGD &initModel, MixtureGD &client, const MAPCfg &cfg);   // TrainTools.cpp:355-382
void computeMAPConst2(const MixtureGD &initModel, MixtureGD &client, const MAPCfg &cfg);  // :388-419
// computeMAP (:547-559): dispatch on MAPAlgo
void computeMAP(const MixtureGD &initModel, MixtureGD &client, const MAPCfg &cfg, double frameCount);
void adaptModel(const Config &c, const FeatureServer &fs, const SegCluster &segs, const MixtureGD &apriori,
                MixtureGD &client, const MAPCfg &cfg);

// ---- TVAcc (AccumulateTVStat.h): same method names, state in HBM
class TVAcc {
 public:
  TVAcc(const std::string &ndxFile, const Config &c);  // :109, _init :129-196
  TVAcc(const std::vector<std::vector<std::string>> &fileLines, const Config &c);  // XList variant
  ~TVAcc();
  void computeAndAccumulateTVStat(const Config &c);  // :268
  void loadT(const std::string &name, const Config &c);  // :632 (transposes when rows > cols)
  void initT(const Config &c);                            // :701 (Box-Muller on libc rand())
  Matrix getT();
  void setStats(const Matrix &N, const Matrix &F);        // statistics held by the caller (JFA: re-centred per iteration)
  void setT(const Matrix &T);                             // [rank x C*D] held by the caller (JFA: V, U or [V; U])
  void saveT(const std::string &name, const Config &c);
  void loadN(const Config &c);
  void loadF_X(const Config &c);
  void saveAccs(const Config &c);  // :1614
  void substractM();
  void estimateTETt();
  void estimateW();
  void estimateAandC();
  void resetTmpAcc();
  void updateTestimate();
  void minDivergence();
  void orthonormalizeT();
  // approximate i-vector modes (IvExtractor.cpp:151-363, TotalVariability.cpp:181-241)
  void normTMatrix();                                        // :1600
  void normStatistics();                                     // :1215
  Matrix getWeightedCov(const std::vector<double> &weight);  // :2826 (W is returned)
  static void computeEigenProblem(const Matrix &EP, Matrix &eigenVect, long rank);  // :2988
  Matrix approximateTcTc(const Matrix &Q);                   // :3106 (D is returned)
  void estimateWUbmWeight(const Matrix &W);                  // :2337
  void estimateWEigenDecomposition(const Matrix &D, const Matrix &Q);  // :2556
  const MixtureGD &world() const { return world_; }
  void saveWbyFile(const Config &c);  // :2799-2822
  void loadMeanEstimate(const std::vector<double> &mean);  // :671
  void reloadStats();      // resend the host copy of N / F_X (TotalVariability.cpp:149-153)
  Matrix getUbmMeans();
  Matrix getW();
  const Matrix &getN() const { return N_; }    // host copies of the raw statistics
  const Matrix &getF_X() const { return F_; }
  size_t nSpeakers() const { return lines_.size(); }
  int rank() const { return R_; }

 private:
  Config cfg_;
  std::vector<std::vector<std::string>> lines_;  // THIS RANK's NDX lines: id + files
  size_t firstLine_ = 0, totalLines_ = 0, totalSessions_ = 0;  // position of the shard in the whole list
  void shardLines();
  MixtureGD world_;
  int R_ = 0;
  lr_tv *tv_ = nullptr;
  Matrix N_, F_;
  void init(const Config &c);
};

// ---- PldaDev (PldaTools.h): development i-vectors, one NDX line per speaker (every element a
// session); statistics and normalisations run on the engine (lr_iv_*), data stays on the host
class PldaDev {
 public:
  PldaDev(const std::string &ndxFilename, const Config &c);  // :97, load :274-350
  size_t getVectSize() const { return data_.rows; }
  size_t getSpeakerNumber() const { return nSpk_; }
  size_t getSessionNumber() const { return data_.cols; }
  const Matrix &getData() const { return data_; }
  const std::vector<double> &getMean() const { return mean_; }
  const std::vector<int32_t> &getClass() const { return class_; }
  void setData(const Matrix &X) { data_ = X; computeAll(); }  // the engine centres _data in place
  void computeAll();                                             // :353-385
  void lengthNorm();                                             // :436-464
  void center(const std::vector<double> &mu);                    // :466-474
  void rotateLeft(const Matrix &M);                              // :498-514
  void computeCovMat(Matrix &Sigma, Matrix &W, Matrix &B);       // :516-571
  void computeWccnChol(Matrix &WCCN);                            // :1113-1175
  void computeMahalanobis(Matrix &M);                            // :1366-1378
  void computeLDA(Matrix &ldaMat, long ldaRank, const Config &c);  // :1381-1415 (ldaMode covariance)
  void computeScatterMat(Matrix &SB, Matrix &SW);  // computeScatterMatUnThreaded :1607-1640, as written
  void sphericalNuisanceNormalization(const Config &c);          // :1822-1929 (estimates, saves, applies)
  void applySphericalNuisanceNormalization(const Config &c);     // :1931-1975

 private:
  Matrix data_;  // [vectSize x sessions]
  std::vector<int32_t> class_;
  size_t nSpk_ = 0;
  std::vector<double> mean_;
};
// ---- PldaModel in training mode (PldaTools.cpp:2028-2122, 2176-2343, 2790-2870): the model
// matrices live on the host, every EM iteration is one lr_plda_em_iteration call
class PldaModel {
 public:
  PldaModel(const std::string &mode /*train*/, const Config &c);  // :2028
  PldaDev &getDev() { return dev_; }
  void updateModel(const Config &c);  // :2302-2326 (after the development data changed dimension)
  void updateMean();                  // :2290
  void centerData();                  // :2297
  void em_iteration(const Config &c, unsigned long it);  // :2329-2343
  void saveModel(const Config &c);    // :2816-2870
  const Matrix &F() const { return F_; }
  const Matrix &G() const { return G_; }
  const Matrix &Sigma() const { return Sigma_; }

 private:
  PldaDev dev_;
  size_t rankF_ = 0, rankG_ = 0;
  Matrix F_, G_, Sigma_;
  std::vector<double> originalMean_, delta_;
  void initModel(const Config &c);  // :2176-2203 (Sigma from the data, Box-Muller F / G)
};

// file names of the EFR / sphNorm parameters of iteration `it` (:1836-1842, :1907-1913)
std::string efrMatrixFilename(const Config &c, unsigned long it, bool forLoad);
std::string efrMeanFilename(const Config &c, unsigned long it, bool forLoad);

// score output of IvTest (IvTest.cpp:412-465): outputScoreFormat ascii = one NIST line per trial,
// segments outer / models inner in matrix order; binary = <out>_model.txt, <out>_testSeg.txt and the
// [models x segments] score matrix saved as <out> + saveMatrixFilesExtension
void writeIvTestScores(const Config &c, const Matrix &scores, const std::vector<uint8_t> &trials,
                       const std::vector<std::string> &modelIds, const std::vector<std::string> &segIds);

// ---- drivers: int Foo(Config&) like the reference programs
int TrainWorld(Config &c);        // LIA_SpkDet/TrainWorld/src/TrainWorld.cpp:101
int ComputeTest(Config &c);       // LIA_SpkDet/ComputeTest/src/ComputeTest.cpp:90
int TrainTargetJFA(Config &c);       // TrainTarget.cpp:393-617 (joint [y; x] with [V; U], z with D, supervector + model)
int TrainTargetLFA(Config &c);       // TrainTarget.cpp:620-760 (the same with D = sqrt(Sigma / tau) and the MAP z)
int TrainTargetDispatch(Config &c);  // TrainTargetMain.cpp:160-171 (channelCompensation)
int ComputeTestDotProduct(Config &c);  // :228-370 (JFA: channel-compensated statistics . client supervector)
int ComputeTestJFA(Config &c);         // :376-572 (JFA: U x removed from the frames, then the top-K LLR)
int ComputeTestLFA(Config &c);         // :574-762 (LFA: the same with the MAP offset D z in the session model, optional cms)
int ComputeTestDispatch(Config &c);    // ComputeTestMain.cpp:137-165 (channelCompensation / scoring)
int IvExtractor(Config &c);       // LIA_SpkDet/IvExtractor/src/IvExtractor.cpp:70 (mode classic)
int ComputeJFAStats(Config &c);  // ComputeJFAStats.cpp:71-87 (per-session + per-speaker BW statistics)
int EigenVoice(Config &c);       // EigenVoice.cpp:71-165 (V trained by the TVAcc device path on speaker statistics)
int EigenChannel(Config &c);     // EigenChannel.cpp:71-160, JFA mode (U trained on the session statistics)
int EigenChannelLFA(Config &c);  // EigenChannel.cpp:178-290 (MAP D, z re-estimated every iteration)
int EigenChannelDispatch(Config &c);  // EigenChannelMain.cpp:139-143 (eigenChannelMode)
int EstimateDMatrix(Config &c);  // EstimateDMatrix.cpp:99-210 (y, x on the device path, diagonal D update)
int IvExtractorUbmWeigth(Config &c);          // IvExtractor.cpp:151 (mode ubmWeight; the reference's spelling)
int IvExtractorEigenDecomposition(Config &c); // IvExtractor.cpp:254 (mode eigenDecomposition)
int TotalVariability(Config &c);  // LIA_SpkDet/TotalVariability/src/TotalVariability.cpp:71
int IvTest(Config &c);            // LIA_SpkDet/IvTest/src/IvTest.cpp:73 (scoring = cosine | mahalanobis | 2cov | plda native)
int IvNorm(Config &c);            // LIA_SpkDet/IvNorm/src/IvNorm.cpp:72
int PLDA(Config &c);              // LIA_SpkDet/PLDA/src/PLDA.cpp:73 (PLDA model training)
int TrainTarget(Config &c);       // LIA_SpkDet/TrainTarget/src/TrainTarget.cpp:75 (MAPOccDep)

}  // namespace lia
