// HostSelfTestMain.cpp -- CPU-only check of the host surface (no engine call): Config, XList,
// MixtureGD RAW <-> XML, Matrix DB <-> DT, FeatureServer with mask, label selection, bagging.
// Prints "key value" lines that tests/test_host_cpu.py compares with numpy.
#include <cstdio>
#include <iostream>

#include "lia_host.h"

int main(int argc, char **argv) {
  try {
    lia::Config c;
    c.parseCmdLine(argc, argv);
    std::cout.precision(17);
    lia::MixtureGD m = lia::MixtureGD::loadFromConfig(c.getParam("inputWorldFilename"), c);
    double sw = 0, sm = 0, sc = 0, scst = 0;
    for (double v : m.w) sw += v;
    for (double v : m.mean) sm += v;
    for (double v : m.cov) sc += v;
    lia::MixtureGD m2 = m;
    m2.computeAll();
    for (int k = 0; k < m.C; k++) scst += m2.cst[k] / m.cst[k];
    std::cout << "gmm " << m.C << " " << m.D << " " << sw << " " << sm << " " << sc << " " << scst / m.C << "\n";
    m.save(c.getParam("tmpPrefix") + ".xml", "XML");
    lia::MixtureGD x;
    x.load(c.getParam("tmpPrefix") + ".xml", "XML");
    double dmax = 0;
    for (size_t i = 0; i < m.mean.size(); i++) dmax = std::max(dmax, std::abs(m.mean[i] - x.mean[i]) + std::abs(m.covinv[i] - x.covinv[i]));
    std::cout << "xml_roundtrip " << dmax << "\n";
    {
      // model surgery of the training loops (no engine call): normalizeMixture, reduceToTopWeights, the MAP variants
      lia::MixtureGD nm = m;
      lia::normalizeMixture(nm, 1, false);
      double gm = 0, gv = 0;  // worst deviation of the normalised mixture's global mean / variance from N(0, 1)
      for (int i = 0; i < nm.D; i++) {
        double a = 0, b = 0;
        for (int k = 0; k < nm.C; k++) {
          a += nm.w[k] * nm.mean[(size_t)k * nm.D + i];
          b += nm.w[k] * (nm.cov[(size_t)k * nm.D + i] + nm.mean[(size_t)k * nm.D + i] * nm.mean[(size_t)k * nm.D + i]);
        }
        gm = std::max(gm, std::abs(a));
        gv = std::max(gv, std::abs(b - a * a - 1.0));
      }
      lia::MixtureGD mo = m;
      lia::normalizeMixture(mo, 2, true);
      double covSame = 0, mo0 = 0;
      for (size_t e = 0; e < m.cov.size(); e++) covSame = std::max(covSame, std::abs(mo.cov[e] - m.cov[e]));
      for (int k = 0; k < mo.C; k++) mo0 += mo.w[k] * mo.mean[(size_t)k * mo.D];
      std::cout << "normalize " << gm << " " << gv << " " << covSame << " " << mo0 << "\n";
      lia::MixtureGD rm = m;
      lia::reduceToTopWeights(rm, 3);
      double sw3 = 0;
      for (double v : rm.w) sw3 += v;
      std::cout << "reduce " << rm.C << " " << sw3 << " " << rm.w[0] << " " << rm.w[1] << " " << rm.w[2] << " " << rm.mean[0] << " "
                << rm.cst[0] << "\n";
      lia::Config mc = c;
      mc.setParam("MAPAlgo", "MAPConst2");
      mc.setParam("meanAdapt", "true");
      mc.setParam("MAPAlphaMean", "0.75");
      lia::MAPCfg cfg(mc);
      lia::MixtureGD cl = m;
      for (double &v : cl.mean) v += 1.0;
      for (int k = 0; k < cl.C; k++) cl.w[k] = 1.0 / cl.C;
      lia::computeMAP(m, cl, cfg, 100.0);
      const double wk = m.w[1], ck = 1.0 / m.C;
      const double expect = (0.75 * wk * m.mean[m.D] + 0.25 * ck * (m.mean[m.D] + 1.0)) / (0.75 * wk + 0.25 * ck);
      std::cout << "map_const2 " << std::abs(cl.mean[m.D] - expect) << " " << std::abs(cl.w[1] - m.w[1]) << "\n";
    }
    lia::XList ndx(c.getParam("ndxFilename"));
    if (c.existsParam("topGauss")) {  // needs the engine: TopGauss::compute -> write -> read for the first test file
      lia::initEngine(c);
      lia::FeatureServer tfs(c, {ndx.lines()[0][0]});
      lia::TopGauss tg, back;
      const double llk = tg.compute(m, tfs, ndx.lines()[0][0], c);
      tg.write(ndx.lines()[0][0] + ".tg", c);
      back.read(ndx.lines()[0][0] + ".tg", c);
      const bool same = back.nt == tg.nt && back.nbgcnt == tg.nbgcnt && back.nbg == tg.nbg && back.idx == tg.idx &&
                        back.snsw == tg.snsw && back.snsl == tg.snsl;
      std::cout << "topgauss " << tg.nt << " " << tg.nbgcnt << " " << llk << " " << same << "\n";
      return 0;
    }
    std::cout << "ndx " << ndx.lines().size() << " " << ndx.allElements().size() << " " << ndx.allUniqueElements().size() << "\n";
    std::vector<std::string> files;
    for (auto &l : ndx.lines()) files.push_back(l[0]);
    lia::FeatureServer fs(c, files);
    double sx = 0;
    for (size_t i = 0; i < fs.getFeatureCount() * fs.ld(); i++) sx += fs.data()[i];
    std::cout << "features " << fs.getFeatureCount() << " " << fs.getVectSize() << " " << sx << "\n";
    lia::SegCluster sel = lia::selectedSegments(c, fs, c.getParam("labelSelectedFrames"));
    double ssel = 0;
    for (auto &s : sel) {
      const float *p = fs.data() + (fs.getFirstFeatureIndexOfASource(s.source) + s.begin) * fs.ld();
      for (long i = 0; i < s.length * (long)fs.ld(); i++) ssel += p[i];
    }
    std::cout << "selected " << sel.size() << " " << lia::totalFrame(sel) << " " << ssel << "\n";
    lia::Matrix a(3, 5);
    for (size_t i = 0; i < a.data.size(); i++) a.data[i] = 0.1 * i - 0.7;
    a.save(c.getParam("tmpPrefix") + ".db", "DB");
    a.save(c.getParam("tmpPrefix") + ".dt", "DT");
    lia::Matrix b, d;
    b.load(c.getParam("tmpPrefix") + ".db", "DB");
    d.load(c.getParam("tmpPrefix") + ".dt", "DT");
    double md = 0;
    for (size_t i = 0; i < a.data.size(); i++) md = std::max(md, std::abs(a.data[i] - b.data[i]) + std::abs(a.data[i] - d.data[i]));
    std::cout << "matrix_roundtrip " << b.rows << " " << d.cols << " " << md << "\n";
    srand(7);
    lia::SegCluster bag = lia::baggedSegments(sel, 0.5, 3, 7);
    long maxlen = 0;
    for (auto &s : bag) maxlen = std::max(maxlen, s.length);
    std::cout << "bagged " << bag.size() << " " << lia::totalFrame(bag) << " " << maxlen << "\n";
    std::cout << "setItParameter " << lia::setItParameter(0.5, 0.1, 5, 2) << " " << lia::timeToFrameIdx(0.299999999, 0.01) << "\n";
    {  // IvTest score output: ascii (segments outer, models inner, trial mask) and binary
      lia::Matrix sc(2, 3);
      for (size_t i = 0; i < 6; i++) sc.data[i] = 0.5 * (double)i - 1.0;
      std::vector<uint8_t> mask = {1, 0, 1, 1, 1, 0};
      lia::Config oc = c;
      oc.setParam("outputFilename", c.getParam("tmpPrefix") + "_scores.res");
      oc.setParam("gender", "F");
      lia::writeIvTestScores(oc, sc, mask, {"m0", "m1"}, {"s0", "s1", "s2"});
      oc.setParam("outputScoreFormat", "binary");
      oc.setParam("outputFilename", c.getParam("tmpPrefix") + "_scores");
      oc.setParam("saveMatrixFilesExtension", ".mat");
      oc.setParam("saveMatrixFormat", "DB");
      lia::writeIvTestScores(oc, sc, mask, {"m0", "m1"}, {"s0", "s1", "s2"});
      lia::Matrix back;
      back.load(c.getParam("tmpPrefix") + "_scores.mat", "DB");
      std::cout << "scores_binary " << back.rows << " " << back.cols << " " << back.data[5] << "\n";
    }
    {  // rank sharding (no device needed): contiguous balanced ranges, by count and by frames
      lia::Shard sh;
      sh.world = 3;
      std::cout << "shard_range";
      for (sh.rank = 0; sh.rank < 3; sh.rank++) {
        auto r = sh.range(10);
        std::cout << " " << r.first << "-" << r.second;
      }
      std::cout << "\nshard_weight";
      const std::vector<double> frames = {100, 100, 100, 900, 100, 100, 100, 300};
      for (sh.rank = 0; sh.rank < 3; sh.rank++) {
        auto r = sh.rangeByWeight(frames);
        std::cout << " " << r.first << "-" << r.second;
      }
      std::cout << "\n";
    }
    try {
      c.getParam("noSuchParameter");
    } catch (lia::Exception &e) {
      std::cout << "exception " << (e.toString().find("noSuchParameter") != std::string::npos) << "\n";
    }
  } catch (std::exception &e) {
    std::cout << "FAILED " << e.what() << std::endl;
    return 1;
  }
  return 0;
}
