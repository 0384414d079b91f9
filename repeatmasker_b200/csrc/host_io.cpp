// host_io.cpp -- matrix / FASTA readers and the double-precision pieces that stay on the host.
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>

#include "host.hpp"

namespace rmgpu {

static std::vector<std::string> split_ws(const std::string& s) {
  std::vector<std::string> out;
  std::istringstream is(s);
  std::string t;
  while (is >> t) out.push_back(t);
  return out;
}

static bool is_int(const std::string& t) {
  if (t.empty()) return false;
  for (char c : t)
    if (!(std::isdigit((unsigned char)c) || c == '-')) return false;
  return true;
}

// Grammar of Matrix.pm:308-364: every line upper-cased; "FREQS A 0.285 C 0.215 ..." ; an alphabet row of
// single letters; integer rows with an optional leading row letter.
ScoreMatrix ScoreMatrix::from_file(const std::string& path) {
  std::ifstream in(path);
  if (!in) throw Error("matrix (" + path + ") does not exist!");
  ScoreMatrix m;
  size_t slash = path.find_last_of("/\\");
  m.name = slash == std::string::npos ? path : path.substr(slash + 1);
  double fr[256];
  std::fill(fr, fr + 256, -1.0);
  bool have_freqs = false;
  int row = 0;
  std::string line;
  while (std::getline(in, line)) {
    for (auto& c : line) c = (char)std::toupper((unsigned char)c);
    size_t fp = line.find("FREQS");
    if (fp != std::string::npos && fp + 5 < line.size() && std::isspace((unsigned char)line[fp + 5])) {
      auto t = split_ws(line.substr(fp + 5));
      for (size_t i = 0; i + 1 < t.size(); i += 2) fr[(unsigned char)t[i][0]] = std::atof(t[i + 1].c_str());
      have_freqs = true;
      continue;
    }
    auto t = split_ws(line);
    if (t.empty()) continue;
    bool letters = t.size() >= 4;
    for (auto& x : t) letters = letters && x.size() == 1 && std::isupper((unsigned char)x[0]);
    if (letters) {
      if ((int)t.size() > kMaxAlpha) throw Error("matrix alphabet larger than 16 symbols: " + path);
      m.n = (int)t.size();
      m.alphabet.clear();
      for (auto& x : t) m.alphabet.push_back(x[0]);
      continue;
    }
    if (m.n == 0) continue;
    size_t off = (t[0].size() == 1 && !std::isdigit((unsigned char)t[0][0])) ? 1 : 0;
    if (t.size() >= off + 4 && is_int(t[off]) && is_int(t[off + 1]) && is_int(t[off + 2]) && is_int(t[off + 3])) {
      if (row < kMaxAlpha)
        for (size_t i = 0; i + off < t.size() && i < (size_t)kMaxAlpha; i++) m.score[row][i] = std::atoi(t[i + off].c_str());
      row++;
    }
  }
  if (m.n == 0 || row < m.n) throw Error("could not parse matrix " + path);
  for (int i = 0; i < m.n; i++) {
    unsigned char c = (unsigned char)m.alphabet[i];
    if (have_freqs)
      m.freq[i] = fr[c] > 0 ? fr[c] : 0.0;
    else
      m.freq[i] = (c == 'A' || c == 'C' || c == 'G' || c == 'T') ? 0.25 : 0.0;
  }
  m.finish();
  return m;
}

CanonTables::CanonTables() {
  for (int i = 0; i < 256; i++) {
    of_letter[i] = kCanonN;
    exact[i] = false;
  }
  for (int k = 0; k < 16; k++) {
    const unsigned char up = (unsigned char)kCanonLetters[k], lo = (unsigned char)std::tolower(up);
    of_letter[up] = of_letter[lo] = (uint8_t)k;
    exact[up] = exact[lo] = true;
  }
  for (int k = 0; k < 16; k++) comp[k] = of_letter[(unsigned char)complement_letter(kCanonLetters[k])];
}

const CanonTables& canon() {
  static const CanonTables t;
  return t;
}

void SeqSet::mark_exotic() {
  const CanonTables& ct = canon();
  exotic.assign(seqs.size(), 0);
  parallel_for((int)seqs.size(), [&](int i) {
    for (unsigned char c : seqs[i])
      if (!ct.exact[c]) {
        exotic[i] = 1;
        break;
      }
  }, 64);
}

char complement_letter(char c) {
  // table of CrossmatchSearchEngine.pm:833 (tr/ACGTYRMKHBVD/TGCARYKMDVBH/), everything else maps to itself
  static const struct Table {
    char t[256];
    Table() {
      for (int i = 0; i < 256; i++) t[i] = (char)i;
      const char* from = "ACGTYRMKHBVD";
      const char* to = "TGCARYKMDVBH";
      for (int i = 0; from[i]; i++) t[(unsigned char)from[i]] = to[i];
    }
  } table;
  return table.t[(unsigned char)c];
}

std::string revcomp_letters(const std::string& s) {
  std::string r(s.size(), 'N');
  for (size_t i = 0; i < s.size(); i++) r[i] = complement_letter(s[s.size() - 1 - i]);
  return r;
}

void ScoreMatrix::finish() {
  // lambda: Matrix.pm:397-464 -- keep the doubling + bisection loop exactly (stops at 1e-5)
  auto sum_at = [&](double lam) {
    double s = 0;
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++)
        if (freq[i] != 0 && freq[j] != 0) s += freq[i] * freq[j] * std::exp(lam * score[i][j]);
    return s;
  };
  double lo = 0, lam = 0.5, s;
  do {
    s = sum_at(lam);
    if (s < 1.0) {
      lo = lam;
      lam *= 2.0;
    }
  } while (s < 1.0);
  double hi = lam;
  while (hi - lo > .00001) {
    lam = (lo + hi) / 2.0;
    s = sum_at(lam);
    if (s >= 1.0)
      hi = lam;
    else
      lo = lam;
  }
  lambda = lam;
  // letters missing from the matrix (B D H V ...) score as N (else last column); 'x' (RepeatMasker:6059) is X
  size_t pn = alphabet.find('N');
  uint8_t fallback = (uint8_t)(pn == std::string::npos ? n - 1 : pn);
  std::fill(code_of, code_of + 256, fallback);
  for (int i = 0; i < n; i++) {
    code_of[(unsigned char)alphabet[i]] = (uint8_t)i;
    code_of[(unsigned char)std::tolower((unsigned char)alphabet[i])] = (uint8_t)i;
  }
  for (int i = 0; i < n; i++) comp_code[i] = code_of[(unsigned char)complement_letter(alphabet[i])];
  for (int k = 0; k < kMaxAlpha; k++) col_of_canon[k] = code_of[(unsigned char)kCanonLetters[k]];
}

int ScoreMatrix::min_score() const {
  int v = score[0][0];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) v = std::min(v, score[i][j]);
  return v;
}
int ScoreMatrix::max_score() const {
  int v = score[0][0];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) v = std::max(v, score[i][j]);
  return v;
}

SeqSet SeqSet::from_fasta(const std::string& path) {
  std::ifstream in(path, std::ios::binary);
  if (!in) throw Error("sequence file (" + path + ") does not exist!");
  SeqSet s;
  std::string line;
  while (std::getline(in, line)) {
    if (!line.empty() && line[0] == '>') {
      size_t b = 1;
      while (b < line.size() && (line[b] == ' ' || line[b] == '\t')) b++;
      size_t e = b;
      while (e < line.size() && !std::isspace((unsigned char)line[e])) e++;
      s.names.push_back(line.substr(b, e - b));
      s.seqs.emplace_back();
    } else if (!s.seqs.empty()) {
      auto& q = s.seqs.back();
      for (char c : line)
        if (std::isalpha((unsigned char)c)) q.push_back((char)std::toupper((unsigned char)c));
    }
  }
  return s;
}

SeqSet SeqSet::from_arrays(int32_t n, const char* const* names, const char* const* seqs, const int32_t* lens) {
  SeqSet s;
  s.seqs.resize(n);
  int64_t total = 0;
  for (int i = 0; i < n; i++) {
    s.names.emplace_back(names && names[i] ? names[i] : ("seq" + std::to_string(i)));
    total += lens[i];
  }
  static const struct Upper {
    char t[256];
    Upper() {
      for (int i = 0; i < 256; i++) t[i] = (char)std::toupper(i);
    }
  } upper;
  parallel_for(n, [&](int i) {
    std::string& q = s.seqs[i];
    q.resize((size_t)lens[i]);
    const char* src = seqs[i];
    for (int32_t k = 0; k < lens[i]; k++) q[k] = upper.t[(unsigned char)src[k]];
  }, total > (1 << 18) ? 1 : n + 1);
  return s;
}

KimuraResult kimura_divergence(const char* qaln, const char* saln, size_t len, bool strandC) {
  // mutation classes of SearchResult.pm:179-217: only pairs of A/C/G/T are "well characterised"; CT TC AG GA are
  // transitions, every other unequal pair a transversion
  auto base = [](char c) { return c == 'A' ? 0 : c == 'G' ? 1 : c == 'C' ? 2 : c == 'T' ? 3 : -1; };
  double transitions = 0, prevTrans = 0;
  int transversions = 0, cpg = 0, wcb = 0;
  char pS = 0;
  for (size_t i = 0; i < len; i++) {
    const size_t at = strandC ? len - 1 - i : i;
    const char sb = strandC ? complement_letter(saln[at]) : saln[at];
    const char qb = strandC ? complement_letter(qaln[at]) : qaln[at];
    if (sb == '-') continue;
    const int bq = base(qb), bs = base(sb);
    const bool wc = bq >= 0 && bs >= 0;
    if (wc) wcb++;
    const int mt = (!wc || bq == bs) ? 0 : ((bq >> 1) == (bs >> 1) ? 1 : 2);  // purine<->purine, pyrimidine<->pyrimidine
    if (pS == 'C' && sb == 'G') {
      cpg++;
      if (mt == 1)
        prevTrans += 1;
      else if (mt == 2)
        transversions++;
      if (prevTrans == 2)
        prevTrans = 1;  // two transitions in a CpG count as one
      else if (prevTrans == 1)
        prevTrans = 1.0 / 10;  // a single one as a tenth
    } else {
      transitions += prevTrans;
      prevTrans = 0;
      if (mt == 1)
        prevTrans = 1;  // held back for the CpG accounting
      else if (mt == 2)
        transversions++;
    }
    pS = sb;
  }
  transitions += prevTrans;
  return {kimura_from_counts(transitions, transversions, wcb), transitions, transversions, wcb, cpg};
}

double kimura_from_counts(double transitions, int transversions, int wcb) {
  double kimura = 100.00;
  if (wcb >= 1) {
    const double p = transitions / wcb, q = (double)transversions / wcb;
    const double logOperand = (1 - (2 * p) - q) * std::pow(1 - (2 * q), 0.5);
    if (logOperand > 0) kimura = std::fabs(-0.5 * std::log(logOperand)) * 100;
  }
  return kimura;
}

int complexity_adjust(int raw, const int64_t counts[kMaxAlpha], const ScoreMatrix& m) {
  double t_factor = 0, t_sum = 0, t_counts = 0;
  for (int c = 0; c < m.n; c++) {
    if (counts[c] > 0 && m.freq[c] > 0 && std::log(m.freq[c]) != 0) {
      double k = (double)counts[c];
      t_factor += k * std::log(k);
      t_sum += k * std::log(m.freq[c]);
      t_counts += k;
    }
  }
  if (t_counts != 0) t_factor -= t_counts * std::log(t_counts);
  t_sum -= t_factor;
  double adj = (double)raw + t_sum / m.lambda + .999;
  int v = (int)adj;
  return v < 0 ? 0 : v;
}

}  // namespace rmgpu
