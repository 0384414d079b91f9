// host.hpp -- host-side data model of librmgpu (C++17).  No oracle code is used anywhere in this directory.
#pragma once
#include <algorithm>
#include <cstdint>
#include <exception>
#include <thread>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace rmgpu {

// Static split of [0, n) over a few host threads (independent records: packing on the way in, post-processing on the
// way out); at least `grain` items per thread.
template <typename F>
void parallel_for(int n, F&& f, int grain = 8) {
  static const int hw = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  const int nt = std::min(hw, std::max(1, n / std::max(grain, 1)));
  if (nt <= 1) {
    for (int i = 0; i < n; i++) f(i);
    return;
  }
  std::vector<std::thread> th;
  std::vector<std::exception_ptr> err(nt);
  for (int t = 0; t < nt; t++)
    th.emplace_back([&, t] {
      try {
        for (int i = t; i < n; i += nt) f(i);
      } catch (...) {
        err[t] = std::current_exception();
      }
    });
  for (auto& x : th) x.join();
  for (auto& e : err)
    if (e) std::rethrow_exception(e);
}

constexpr int kMaxAlpha = 16;  // 4-bit residue codes on the device

// Residue codes on the device are matrix-independent ("canonical"): the 15 IUPAC nucleotide letters and X, so that the
// packed library and query records stay valid when only the scoring matrix changes (RepeatMasker switches GC-bin
// matrices per batch, RepeatMasker:4784-4821) and the device can spell alignment rows itself.  Any other letter is
// scored as N like the oracle does and marks its record "exotic" (its rows are then spelled on the host).
constexpr char kCanonLetters[17] = "ACGTRYKMSWBDHVNX";
constexpr uint8_t kCanonN = 14;
struct CanonTables {
  uint8_t of_letter[256];  // upper- or lower-case letter -> canonical code
  bool exact[256];         // the letter is one of the 16 (else it was folded into N)
  uint8_t comp[16];        // canonical code of the complementary letter (table of CrossmatchSearchEngine.pm:833)
  CanonTables();
};
const CanonTables& canon();

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// Scoring matrix in the grammar of Matrix.pm:308-364; score[subject][query] (SearchResult.pm:1894-1917).
struct ScoreMatrix {
  int n = 0;
  std::string alphabet;
  int score[kMaxAlpha][kMaxAlpha] = {};
  double freq[kMaxAlpha] = {};
  double lambda = 0;
  uint8_t code_of[256] = {};  // letter -> matrix column (fallback: N, else last column)
  uint8_t comp_code[kMaxAlpha] = {};  // column of the complementary letter
  uint8_t col_of_canon[kMaxAlpha] = {};  // canonical device code -> matrix column (= code_of[kCanonLetters[c]])
  std::string name;           // basename, reported as matrixName (CrossmatchSearchEngine.pm:398-409)

  static ScoreMatrix from_file(const std::string& path);
  void finish();  // freqs -> lambda (Matrix.pm:397-464), letter maps
  int min_score() const;
  int max_score() const;
};

struct SeqSet {
  std::vector<std::string> names;
  std::vector<std::string> seqs;  // upper-case letters (FastaDB.pm:1110-1118 upper-cases on load)
  std::vector<uint8_t> exotic;    // per record: holds a letter outside the canonical 16 (filled by mark_exotic)
  void mark_exotic();
  int64_t total_bases() const {
    int64_t t = 0;
    for (auto& s : seqs) t += (int64_t)s.size();
    return t;
  }
  static SeqSet from_fasta(const std::string& path);
  static SeqSet from_arrays(int32_t n, const char* const* names, const char* const* seqs, const int32_t* lens);
};

char complement_letter(char c);  // table of CrossmatchSearchEngine.pm:833
std::string revcomp_letters(const std::string& s);

// SearchResult.pm:1607-1683 (calcKimuraDivergence with divCpGMod): rows as the engine reports them; for strand C they
// are read back to front through the complement table, i.e. as CrossmatchSearchEngine.pm:824-837 stores them.
struct KimuraResult {
  double kimura, transitions;
  int transversions, wellCharacterized, cpgSites;
};
KimuraResult kimura_divergence(const char* qaln, const char* saln, size_t len, bool strandC);
// the final formula of the same routine from its three counts: |-1/2 ln((1 - 2p - q) sqrt(1 - 2q))| * 100
double kimura_from_counts(double transitions, int transversions, int wellCharacterized);

// SearchResult.pm:2111-2150 in double precision, perl int() truncation.
int complexity_adjust(int raw, const int64_t counts[kMaxAlpha], const ScoreMatrix& m);

// Owner of the strings an rmgpu_result points to (rmgpu_result::opaque).
struct ResultArena {
  std::vector<std::unique_ptr<char[]>> blocks;
  std::vector<std::unique_ptr<ResultArena>> children;  // arenas of merged partial results
  const char* dup(const char* p, size_t n) {
    std::unique_ptr<char[]> b(new char[n + 1]);
    if (n) memcpy(b.get(), p, n);
    b[n] = 0;
    blocks.push_back(std::move(b));
    return blocks.back().get();
  }
  const char* dup(const std::string& s) { return dup(s.data(), s.size()); }
  std::vector<std::vector<char>> buffers;  // whole buffers handed over without a copy (the packed alignment rows)
  const char* keep_buffer(std::vector<char>&& b) {
    buffers.emplace_back(std::move(b));
    return buffers.back().data();
  }
  std::vector<std::unique_ptr<std::string>> strings;  // strings handed over without a copy
  const char* keep(std::string&& s) {
    strings.emplace_back(new std::string(std::move(s)));
    return strings.back()->c_str();
  }
};

}  // namespace rmgpu
