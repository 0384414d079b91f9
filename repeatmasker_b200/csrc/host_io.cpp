// host_io.cpp -- matrix / FASTA readers and the double-precision pieces that stay on the host.
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>

#include "../../include/rmgpu.h"
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


// ---------------------------------------------------------------------------------------------------------------
// The alignment block of a .cat file entry.  What SearchResult::_toCrossMatchFormat
// (SearchResult.pm:2486-2834) appends to the annotation line in mode AlignWithQuerySeq: 50-column chunks of
// "  <name 13> <start 10> <row> <end>", the line of difference marks between them (i transition, v transversion,
// - gap, ? a pair with an ambiguity code), then the Matrix / Kimura / transition / gap statistics.  The reference walks
// every column in Perl; the batch children of a run spend more time here than the GPU spends searching.
// Rows are the ones the SearchResult object holds (query forward; for orientation C the subject coordinates descend).
// ---------------------------------------------------------------------------------------------------------------
namespace {

inline bool ambiguity_code(int c) {  // the reference's /[BDHVRYKMSWNX]/i
  switch (c) {
    case 'B': case 'D': case 'H': case 'V': case 'R': case 'Y': case 'K': case 'M': case 'S': case 'W': case 'N': case 'X':
      return true;
  }
  return false;
}

// the reference's %mutChar (SearchResult.pm:2438-2484) as a rule: purine<->purine / pyrimidine<->pyrimidine 'i', other
// A/C/G/T pairs 'v', one of the 16 nucleotide codes against '-' a gap; 0 = not in the table
inline char difference_mark(int q, int s) {
  auto base = [](int c) { return c == 'A' ? 0 : c == 'G' ? 1 : c == 'C' ? 2 : c == 'T' ? 3 : -1; };
  const int bq = base(q), bs = base(s);
  if (bq >= 0 && bs >= 0) return (bq >> 1) == (bs >> 1) ? 'i' : 'v';
  auto code = [&](int c) { return base(c) >= 0 || ambiguity_code(c); };
  if ((q == '-' && code(s)) || (s == '-' && code(q))) return '-';
  return 0;
}

void name_field(std::string& out, const char* name) {  // up to 13 characters, padded, and the separating blank
  const size_t n = std::min<size_t>(strlen(name), 13);
  out.append(name, n);
  out.append(13 - n + 1, ' ');
}

void number_field(std::string& out, long long v) {  // right-aligned in 10 columns (wider numbers are not cut)
  const std::string t = std::to_string(v);
  if (t.size() < 10) out.append(10 - t.size(), ' ');
  out += t;
}

}  // namespace
}  // namespace rmgpu

extern "C" int64_t rmgpu_format_alignment(const char* query_name, const char* subject_name, const char* qaln, const char* saln,
                                          int64_t qstart, int64_t qend, int64_t sstart, int64_t send, int strand,
                                          const char* matrix_name, const char* pct_kimura, const char* cpg_sites,
                                          const char* pct_raw_kimura, char* out, int64_t cap) {
  using namespace rmgpu;
  const size_t len = strlen(qaln);
  if (strlen(saln) != len) return -1;
  std::string r;
  r.reserve(len * 5 + 512);
  r += '\n';
  long long qStart = qstart, qEnd = 0, sStart = strand ? send : sstart, sEnd = 0;
  long long transitions = 0, transversions = 0, gapChars = 0, gaps = 0;
  for (size_t off = 0; off < len; off += 50) {
    const size_t n = std::min<size_t>(50, len - off);
    const char *q = qaln + off, *s = saln + off;
    const long long insertions = std::count(q, q + n, '-'), deletions = std::count(s, s + n, '-');
    if (sEnd > 0) {  // every chunk but the first (the reference tests the previous subject end, :2680)
      const long long qIncr = (long long)n > insertions, sIncr = (long long)n > deletions;
      qStart = qEnd + qIncr;
      sStart = strand ? sEnd - sIncr : sEnd + sIncr;
    }
    qEnd = (long long)n == insertions ? qStart : qStart + (long long)n - 1 - insertions;
    r += "  ";
    name_field(r, query_name);
    number_field(r, qStart);
    r += ' ';
    r.append(q, n);
    r += ' ';
    r += std::to_string(qEnd);
    r += '\n';
    r.append(27, ' ');
    for (size_t j = 0; j < n; j++) {
      const int a = std::toupper((unsigned char)q[j]), b = std::toupper((unsigned char)s[j]);
      char mark = ' ';
      if (a != b) {
        const char mc = difference_mark(a, b);
        if (mc) {
          mark = mc;
          if (mc == 'i') transitions++;
          if (mc == 'v') transversions++;
          if (mc == '-') {
            gapChars++;
            // a gap "opens" where the previous QUERY column of the chunk is no gap; for the chunk's first column the
            // reference's `$qChars[ ( $j - 1 ) | 0 ]` (:2723) is index -1, the chunk's LAST column
            if (q[j ? j - 1 : n - 1] != '-') gaps++;
          }
        } else if (ambiguity_code(a) || ambiguity_code(b)) {
          mark = '?';
        }
      }
      r += mark;
    }
    r += '\n';
    if (strand) {
      r += "C ";
      sEnd = sStart - (long long)n + 1 + deletions;
    } else {
      r += "  ";
      sEnd = sStart + (long long)n - 1 - deletions;
    }
    name_field(r, subject_name);
    number_field(r, sStart);
    r += ' ';
    r.append(s, n);
    r += ' ';
    r += std::to_string(sEnd);
    r += "\n\n";
  }
  r += "Matrix = ";
  r += matrix_name ? matrix_name : "Unknown";
  r += '\n';
  if (pct_kimura) {
    r += "Kimura (with divCpGMod) = ";
    r += pct_kimura;
    r += '\n';
  }
  if (cpg_sites && pct_raw_kimura) {
    r += "CpG sites = ";
    r += cpg_sites;
    r += ", Kimura (unadjusted) = ";
    r += pct_raw_kimura;
    r += '\n';
  }
  char buf[64];
  r += "Transitions / transversions = ";
  if (transversions) {
    snprintf(buf, sizeof buf, "%0.2f", (double)transitions / (double)transversions);
    r += buf;
  } else {
    r += "1.00";
  }
  r += " (" + std::to_string(transitions) + "/" + std::to_string(transversions) + ")\n";
  const long long span = qend - qstart;
  if (span > 0) {
    snprintf(buf, sizeof buf, "%0.2f", (double)gaps / (double)span);
    r += "Gap_init rate = ";
    r += buf;
    r += " (" + std::to_string(gaps) + " / " + std::to_string(span) + ")";
  } else {
    r += "Gap_init rate = 0.0 ( " + std::to_string(gaps) + " / 0 )";
  }
  if (gaps > 0) {
    snprintf(buf, sizeof buf, "%0.2f", (double)gapChars / (double)gaps);
    r += ", avg. gap size = ";
    r += buf;
    r += " (" + std::to_string(gapChars) + " / " + std::to_string(gaps) + ")\n\n";
  } else {
    r += ", avg. gap size = 0.0 (0 / 0)\n\n";
  }
  if (out && cap > 0) {
    const size_t n = std::min<size_t>(r.size(), (size_t)cap - 1);
    memcpy(out, r.data(), n);
    out[n] = 0;
  }
  return (int64_t)r.size();
}
