// base.cpp -- Config, XList, Matrix and MixtureGD file formats, labels, FeatureServer.
// File-format ground truth: SURVEY.md §8b (probed on the reference's fixtures).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <set>
#include <sstream>

#include "lia_host.h"

namespace lia {

Exception::Exception(const std::string &msg, const char *file, int line)
    : std::runtime_error("Exception: " + msg + " [" + file + ":" + std::to_string(line) + "]") {}

void check(lr_status st, const char *file, int line) {
  if (st != LR_OK) throw Exception(std::string("engine: ") + lr_last_error(), file, line);
}

// ------------------------------------------------------------------ Config
static std::string trim(const std::string &s) {
  size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
  return a == std::string::npos ? "" : s.substr(a, b - a + 1);
}

void Config::load(const std::string &file) {
  std::ifstream in(file.c_str());
  if (!in) LIA_THROW("Config file not found: " + file);
  std::string line;
  while (std::getline(in, line)) {
    line = trim(line);
    if (line.empty() || line[0] == '#' || line[0] == '*' || line[0] == '%') continue;  // banners
    size_t p = line.find_first_of(" \t");
    std::string name = line.substr(0, p), value = p == std::string::npos ? "" : trim(line.substr(p));
    kv_[name] = value;
  }
}

void Config::parseCmdLine(int argc, char **argv) {
  for (int i = 1; i + 1 < argc; i++)
    if (std::string(argv[i]) == "--config") load(argv[i + 1]);
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    if (a.rfind("--", 0) != 0) LIA_THROW("unexpected command line token: " + a);
    std::string name = a.substr(2);
    if (i + 1 < argc && std::string(argv[i + 1]).rfind("--", 0) != 0) {
      if (name != "config") kv_[name] = argv[i + 1];
      i++;
    } else {
      kv_[name] = "true";
    }
  }
}

const std::string &Config::getParam(const std::string &n) const {
  auto it = kv_.find(n);
  if (it == kv_.end()) LIA_THROW("Parameter not found in the configuration: " + n);
  return it->second;
}
std::string Config::getString(const std::string &n, const std::string &def) const {
  return existsParam(n) ? getParam(n) : def;
}
long Config::getLong(const std::string &n) const { return std::stol(getParam(n)); }
long Config::getLong(const std::string &n, long def) const { return existsParam(n) ? getLong(n) : def; }
double Config::getDouble(const std::string &n) const { return std::stod(getParam(n)); }
double Config::getDouble(const std::string &n, double def) const {
  return existsParam(n) ? getDouble(n) : def;
}
bool Config::getBool(const std::string &n, bool def) const {
  if (!existsParam(n)) return def;
  std::string v = getParam(n);
  std::transform(v.begin(), v.end(), v.begin(), ::tolower);
  return v == "true" || v == "1" || v == "yes";
}

// ------------------------------------------------------------------ XList
void XList::load(const std::string &file) {
  std::ifstream in(file.c_str());
  if (!in) LIA_THROW("List file not found: " + file);
  lines_.clear();
  std::string line;
  while (std::getline(in, line)) {
    std::istringstream ss(line);
    std::vector<std::string> tok;
    std::string t;
    while (ss >> t) tok.push_back(t);
    if (!tok.empty()) lines_.push_back(tok);
  }
}
std::vector<std::string> XList::allElements() const {
  std::vector<std::string> out;
  for (auto &l : lines_) out.insert(out.end(), l.begin(), l.end());
  return out;
}
std::vector<std::string> XList::allUniqueElements() const {
  std::vector<std::string> out;
  std::set<std::string> seen;
  for (auto &l : lines_)
    for (auto &e : l)
      if (seen.insert(e).second) out.push_back(e);
  return out;
}

// ------------------------------------------------------------------ Matrix
void Matrix::load(const std::string &file, const std::string &format) {
  if (format == "DB") {
    FILE *f = fopen(file.c_str(), "rb");
    if (!f) LIA_THROW("Matrix file not found: " + file);
    uint32_t r = 0, c = 0;
    bool ok = fread(&r, 4, 1, f) == 1 && fread(&c, 4, 1, f) == 1;
    rows = r;
    cols = c;
    data.assign(rows * cols, 0.0);
    ok = ok && fread(data.data(), sizeof(double), rows * cols, f) == rows * cols;
    fclose(f);
    if (!ok) LIA_THROW("Truncated DB matrix file: " + file);
  } else if (format == "DT") {
    std::ifstream in(file.c_str());
    if (!in) LIA_THROW("Matrix file not found: " + file);
    in >> rows >> cols;
    data.assign(rows * cols, 0.0);
    for (auto &v : data)
      if (!(in >> v)) LIA_THROW("Truncated DT matrix file: " + file);
  } else {
    LIA_THROW("Unknown matrix format: " + format);
  }
}
void Matrix::save(const std::string &file, const std::string &format) const {
  if (format == "DB") {
    FILE *f = fopen(file.c_str(), "wb");
    if (!f) LIA_THROW("Cannot write matrix file: " + file);
    uint32_t r = (uint32_t)rows, c = (uint32_t)cols;
    fwrite(&r, 4, 1, f);
    fwrite(&c, 4, 1, f);
    fwrite(data.data(), sizeof(double), data.size(), f);
    fclose(f);
  } else if (format == "DT") {
    std::ofstream out(file.c_str());
    if (!out) LIA_THROW("Cannot write matrix file: " + file);
    out.precision(17);
    out << rows << " " << cols << "\n";
    for (size_t i = 0; i < rows; i++) {
      for (size_t j = 0; j < cols; j++) out << (*this)(i, j) << " ";
      out << "\n";
    }
  } else {
    LIA_THROW("Unknown matrix format: " + format);
  }
}

// ------------------------------------------------------------------ MixtureGD
void MixtureGD::resize(int c, int d) {
  C = c;
  D = d;
  w.assign(C, 1.0 / C);
  mean.assign((size_t)C * D, 0.0);
  cov.assign((size_t)C * D, 1.0);
  covinv.assign((size_t)C * D, 1.0);
  cst.assign(C, 0.0);
  det.assign(C, 1.0);
}
void MixtureGD::computeAll() {
  const double pi2 = 2.0 * 3.14159265358979323846;
  for (int c = 0; c < C; c++) {
    double dt = 1.0;
    for (int i = 0; i < D; i++) {
      dt *= cov[(size_t)c * D + i];
      covinv[(size_t)c * D + i] = 1.0 / cov[(size_t)c * D + i];
    }
    det[c] = dt;
    cst[c] = 1.0 / (std::pow(pi2, 0.5 * D) * std::sqrt(dt));
  }
}
void MixtureGD::load(const std::string &file, const std::string &format) {
  if (format == "RAW") {
    // uint32 C, uint32 D, double w[C], then per component: cst, det, 1 flag byte, covInv[D], mean[D]
    FILE *f = fopen(file.c_str(), "rb");
    if (!f) LIA_THROW("Mixture file not found: " + file);
    uint32_t c = 0, d = 0;
    bool ok = fread(&c, 4, 1, f) == 1 && fread(&d, 4, 1, f) == 1;
    if (!ok || c == 0 || d == 0 || c > (1u << 20) || d > 4096) {
      fclose(f);
      LIA_THROW("Bad RAW mixture header: " + file);
    }
    resize((int)c, (int)d);
    ok = fread(w.data(), 8, C, f) == (size_t)C;
    for (int k = 0; ok && k < C; k++) {
      unsigned char flag;
      ok = fread(&cst[k], 8, 1, f) == 1 && fread(&det[k], 8, 1, f) == 1 && fread(&flag, 1, 1, f) == 1 &&
           fread(&covinv[(size_t)k * D], 8, D, f) == (size_t)D &&
           fread(&mean[(size_t)k * D], 8, D, f) == (size_t)D;
    }
    fclose(f);
    if (!ok) LIA_THROW("Truncated RAW mixture file: " + file);
    for (size_t i = 0; i < cov.size(); i++) cov[i] = 1.0 / covinv[i];
  } else if (format == "XML") {
    std::ifstream in(file.c_str());
    if (!in) LIA_THROW("Mixture file not found: " + file);
    std::stringstream buf;
    buf << in.rdbuf();
    const std::string t = buf.str();
    auto attr = [&](size_t from, const std::string &name) -> std::string {
      size_t p = t.find(name + "=\"", from);
      if (p == std::string::npos) LIA_THROW("XML mixture: attribute " + name + " missing in " + file);
      p += name.size() + 2;
      return t.substr(p, t.find('"', p) - p);
    };
    size_t h = t.find("<MixtureGD");
    if (h == std::string::npos) LIA_THROW("XML mixture: no <MixtureGD> in " + file);
    id = attr(h, "id");
    resize(std::stoi(attr(h, "distribCount")), std::stoi(attr(h, "vectSize")));
    size_t pos = h;
    for (int k = 0; k < C; k++) {
      pos = t.find("<DistribGD", pos + 1);
      if (pos == std::string::npos) LIA_THROW("XML mixture: missing <DistribGD> in " + file);
      int idx = std::stoi(attr(pos, "i"));
      w[idx] = std::stod(attr(pos, "weight"));
      cst[idx] = std::stod(attr(pos, "cst"));
      det[idx] = std::stod(attr(pos, "det"));
      size_t end = t.find("</DistribGD>", pos);
      for (const char *tag : {"covInv", "mean"}) {
        size_t q = pos;
        std::string open = std::string("<") + tag + " i=\"";
        while ((q = t.find(open, q)) != std::string::npos && q < end) {
          q += open.size();
          int i = std::stoi(t.substr(q, t.find('"', q) - q));
          size_t v0 = t.find('>', q) + 1;
          double v = std::stod(t.substr(v0, t.find('<', v0) - v0));
          (std::string(tag) == "mean" ? mean : covinv)[(size_t)idx * D + i] = v;
        }
      }
    }
    for (size_t i = 0; i < cov.size(); i++) cov[i] = 1.0 / covinv[i];
  } else {
    LIA_THROW("Unknown mixture format: " + format);
  }
}
void MixtureGD::save(const std::string &file, const std::string &format) const {
  if (format == "RAW") {
    FILE *f = fopen(file.c_str(), "wb");
    if (!f) LIA_THROW("Cannot write mixture file: " + file);
    uint32_t c = (uint32_t)C, d = (uint32_t)D;
    fwrite(&c, 4, 1, f);
    fwrite(&d, 4, 1, f);
    fwrite(w.data(), 8, C, f);
    for (int k = 0; k < C; k++) {
      unsigned char flag = 0;
      fwrite(&cst[k], 8, 1, f);
      fwrite(&det[k], 8, 1, f);
      fwrite(&flag, 1, 1, f);
      fwrite(&covinv[(size_t)k * D], 8, D, f);
      fwrite(&mean[(size_t)k * D], 8, D, f);
    }
    fclose(f);
  } else if (format == "XML") {
    std::ofstream out(file.c_str());
    if (!out) LIA_THROW("Cannot write mixture file: " + file);
    out.precision(19);
    out << "<MixtureGD version=\"1\" id=\"" << (id.empty() ? "#1" : id) << "\" distribCount=\"" << C
        << "\" vectSize=\"" << D << "\">\n";
    for (int k = 0; k < C; k++) {
      out << "\t<DistribGD i=\"" << k << "\" weight=\"" << w[k] << "\" cst=\"" << cst[k] << "\" det=\""
          << det[k] << "\">\n";
      for (int i = 0; i < D; i++)
        out << "\t\t<covInv i=\"" << i << "\">" << covinv[(size_t)k * D + i] << "</covInv>\n";
      for (int i = 0; i < D; i++)
        out << "\t\t<mean i=\"" << i << "\">" << mean[(size_t)k * D + i] << "</mean>\n";
      out << "\t</DistribGD>\n";
    }
    out << "</MixtureGD>\n";
  } else {
    LIA_THROW("Unknown mixture format: " + format);
  }
}
MixtureGD MixtureGD::loadFromConfig(const std::string &name, const Config &c) {
  MixtureGD m;
  m.load(c.getString("mixtureFilesPath", "") + name + c.getString("loadMixtureFileExtension", ""),
         c.getString("loadMixtureFileFormat", "RAW"));
  m.id = name;
  return m;
}
void MixtureGD::saveFromConfig(const std::string &name, const Config &c) const {
  save(c.getString("mixtureFilesPath", "") + name + c.getString("saveMixtureFileExtension", ""),
       c.getString("saveMixtureFileFormat", "RAW"));
}

// ------------------------------------------------------------------ segments
long timeToFrameIdx(double t, double frameLength) {
  double whole;
  double frac = std::modf(t / frameLength, &whole);
  return frac > 0.99999 ? (long)whole + 1 : (long)whole;
}
long totalFrame(const SegCluster &c) {
  long n = 0;
  for (auto &s : c) n += s.length;
  return n;
}

std::vector<int> parseMask(const std::string &mask) {
  std::vector<int> out;
  std::stringstream ss(mask);
  std::string part;
  while (std::getline(ss, part, ',')) {
    part = trim(part);
    if (part.empty()) continue;
    size_t dash = part.find('-');
    if (dash == std::string::npos) {
      out.push_back(std::stoi(part));
    } else {
      int a = std::stoi(part.substr(0, dash)), b = std::stoi(part.substr(dash + 1));
      for (int i = a; i <= b; i++) out.push_back(i);
    }
  }
  return out;
}

// ------------------------------------------------------------------ FeatureServer
static uint32_t bswap32(uint32_t v) { return (v >> 24) | ((v >> 8) & 0xFF00u) | ((v << 8) & 0xFF0000u) | (v << 24); }
static uint16_t bswap16(uint16_t v) { return (uint16_t)((v >> 8) | (v << 8)); }

// bigEndian: the payload (and the binary header fields) of SPRO3 / SPRO4 / RAW files are byte
// swapped; HTK files are big-endian by definition.
static void readFeatureFile(const std::string &path, const std::string &format, int vectSizeCfg, bool bigEndian,
                            std::vector<float> &raw, int &dim, size_t &frames) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) LIA_THROW("Feature file not found: " + path);
  fseek(f, 0, SEEK_END);
  long size = ftell(f);
  fseek(f, 0, SEEK_SET);
  long off = 0;
  if (format == "SPRO3") {
    // 16-byte header (4 x uint32, third = frame count), then frames x dim float32
    uint32_t h[4];
    if (fread(h, 4, 4, f) != 4) {
      fclose(f);
      LIA_THROW("Truncated SPRO3 file: " + path);
    }
    off = 16;
    frames = bigEndian ? bswap32(h[2]) : h[2];
    if (frames == 0 || (size - off) % (4 * (long)frames) != 0) {
      fclose(f);
      LIA_THROW("Inconsistent SPRO3 header: " + path);
    }
    dim = (int)((size - off) / 4 / (long)frames);
  } else if (format == "SPRO4") {
    // optional "<header> ... </header>" text, then uint16 dim, uint32 flags, float rate
    char tag[8] = {0};
    if (fread(tag, 1, 8, f) == 8 && std::strncmp(tag, "<header>", 8) == 0) {
      std::string buf;
      int ch;
      while ((ch = fgetc(f)) != EOF) {
        buf.push_back((char)ch);
        if (buf.size() >= 9 && buf.compare(buf.size() - 9, 9, "</header>") == 0) break;
      }
      fgetc(f);  // newline
      off = ftell(f);
    } else {
      fseek(f, 0, SEEK_SET);
    }
    uint16_t d16;
    uint32_t flags;
    float rate;
    if (fread(&d16, 2, 1, f) != 1 || fread(&flags, 4, 1, f) != 1 || fread(&rate, 4, 1, f) != 1) {
      fclose(f);
      LIA_THROW("Truncated SPRO4 file: " + path);
    }
    off += 10;
    dim = bigEndian ? bswap16(d16) : d16;
    if (dim <= 0) {
      fclose(f);
      LIA_THROW("Bad SPRO4 dimension: " + path);
    }
    frames = (size_t)((size - off) / (4 * (long)dim));
  } else if (format == "HTK") {
    // 12-byte big-endian header: int32 nSamples, int32 sampPeriod (100 ns), int16 sampSize (bytes),
    // int16 parmKind; uncompressed float32 parameters, big-endian
    uint32_t ns, period;
    uint16_t ssize, kind;
    if (fread(&ns, 4, 1, f) != 1 || fread(&period, 4, 1, f) != 1 || fread(&ssize, 2, 1, f) != 1 ||
        fread(&kind, 2, 1, f) != 1) {
      fclose(f);
      LIA_THROW("Truncated HTK file: " + path);
    }
    off = 12;
    frames = bswap32(ns);
    dim = bswap16(ssize) / 4;
    kind = bswap16(kind);
    if ((kind & 0x0400) || dim <= 0 || (long)frames * dim * 4 > size - off) {  // _C: compressed
      fclose(f);
      LIA_THROW("Unsupported (compressed) or inconsistent HTK file: " + path);
    }
    bigEndian = true;
  } else if (format == "RAW") {
    if (vectSizeCfg <= 0) {
      fclose(f);
      LIA_THROW("RAW feature files need the vectSize parameter");
    }
    dim = vectSizeCfg;
    frames = (size_t)(size / (4 * (long)dim));
  } else {
    fclose(f);
    LIA_THROW("Unsupported loadFeatureFileFormat: " + format);
  }
  raw.resize(frames * (size_t)dim);
  fseek(f, off, SEEK_SET);
  size_t got = fread(raw.data(), 4, raw.size(), f);
  fclose(f);
  if (got != raw.size()) LIA_THROW("Truncated feature file: " + path);
  if (bigEndian) {
    uint32_t *w = reinterpret_cast<uint32_t *>(raw.data());
    for (size_t i = 0; i < raw.size(); i++) w[i] = bswap32(w[i]);
  }
}

FeatureServer::FeatureServer(const Config &c, const std::vector<std::string> &files) {
  const std::string path = c.getString("featureFilesPath", ""), ext = c.getString("loadFeatureFileExtension", "");
  const std::string format = c.getString("loadFeatureFileFormat", "SPRO4");
  const bool bigEndian = c.getBool("bigEndian", false);
  std::vector<int> mask;
  if (c.existsParam("featureServerMask")) mask = parseMask(c.getParam("featureServerMask"));
  for (auto &name : files) {
    std::vector<float> raw;
    int dim = 0;
    size_t frames = 0;
    readFeatureFile(path + name + ext, format, (int)c.getLong("vectSize", 0), bigEndian, raw, dim, frames);
    std::vector<int> m = mask;
    if (m.empty())
      for (int i = 0; i < dim; i++) m.push_back(i);
    for (int i : m)
      if (i < 0 || i >= dim) LIA_THROW("featureServerMask selects coefficient outside the file vectSize: " + name);
    if (D_ == 0) D_ = (int)m.size();
    if ((int)m.size() != D_) LIA_THROW("Feature files with different vectSize: " + name);
    names_.push_back(name);
    first_.push_back(getFeatureCount());
    count_.push_back(frames);
    size_t base = X_.size();
    X_.resize(base + frames * (size_t)D_);
    for (size_t t = 0; t < frames; t++)
      for (int j = 0; j < D_; j++) X_[base + t * D_ + j] = raw[t * dim + m[j]];
  }
}
size_t FeatureServer::getFirstFeatureIndexOfASource(const std::string &name) const {
  for (size_t i = 0; i < names_.size(); i++)
    if (names_[i] == name) return first_[i];
  LIA_THROW("Unknown feature source: " + name);
}
size_t FeatureServer::getFeatureCountOfASource(const std::string &name) const {
  for (size_t i = 0; i < names_.size(); i++)
    if (names_[i] == name) return count_[i];
  LIA_THROW("Unknown feature source: " + name);
}

SegCluster selectedSegments(const Config &c, const FeatureServer &fs, const std::string &label) {
  SegCluster out;
  const std::string lpath = c.getString("labelFilesPath", ""), lext = c.getString("labelFilesExtension", ".lbl");
  const double frameLength = c.getDouble("frameLength", 0.01);
  const bool addDefault = c.getBool("addDefaultLabel", false);
  const std::string defLabel = c.getString("defaultLabel", "");
  for (size_t s = 0; s < fs.getSourceCount(); s++) {
    const std::string &src = fs.getNameOfASource(s);
    const long n = (long)fs.getFeatureCountOfASource(src);
    std::ifstream in((lpath + src + lext).c_str());
    if (!in) {
      // no label file: the whole file carries the default label when addDefaultLabel is set
      if (addDefault && defLabel == label) out.push_back({src, 0, n, label});
      continue;
    }
    double b, e;
    std::string lab;
    while (in >> b >> e >> lab) {
      if (lab != label) continue;
      long fb = timeToFrameIdx(b, frameLength), fe = timeToFrameIdx(e, frameLength);  // end inclusive
      if (fb >= n) continue;                                                           // verifyClusterFile
      if (fe >= n) fe = n - 1;
      if (fe >= fb) out.push_back({src, fb, fe - fb + 1, lab});
    }
  }
  return out;
}

std::vector<lr_seg> toEngineSegs(const FeatureServer &fs, const SegCluster &segs, int row) {
  std::vector<lr_seg> out;
  out.reserve(segs.size());
  for (auto &s : segs) {
    lr_seg e;
    e.begin = (int64_t)(s.begin + (long)fs.getFirstFeatureIndexOfASource(s.source));
    e.length = s.length;
    e.row = row;
    e.pad_ = 0;
    out.push_back(e);
  }
  return out;
}

}  // namespace lia
