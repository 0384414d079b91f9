// engine.cu -- session, device layout, kernel orchestration and the C ABI of librmgpu.so (include/rmgpu.h).
//
// Data layout in HBM:
//   library   rowcode[nGroups*R] (1 B/row, 0xFF = pad), prof[nGroups*R] (8 B/row, rebuilt per matrix / class map),
//             gflags/gsubject per 32-lane group, sflags per stripe, scodes (1 B/base) for K2/K3
//   queries   qpacked (4 bit/base, what crosses PCIe), sel (2 B/column, both strands), codes_fwd/codes_rc (1 B/base)
//   scratch   stripe boundary ring (K1 8 B/column, K2 16 B/column), candidate table, traceback direction bytes
#include <algorithm>
#include <array>
#include <chrono>
#include <cstring>
#include <map>
#include <memory>
#include <numeric>
#include <set>
#include <atomic>
#include <exception>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rmgpu.h"
#include "host.hpp"
#include "kernels.cuh"
#include "remote.hpp"

namespace rmgpu {

#define CUDA_CHECK(x)                                                                                         \
  do {                                                                                                        \
    cudaError_t e_ = (x);                                                                                     \
    if (e_ != cudaSuccess)                                                                                    \
      throw Error(std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " + __FILE__ + ":" +          \
                  std::to_string(__LINE__) + " (" #x ")");                                                    \
  } while (0)

static thread_local std::string g_last_error;

// bytes held by the DevBufs of this process, per device: lets a session track free HBM without asking the driver on
// every call
static std::atomic<long long> g_devbuf_bytes[64];

template <typename T>
struct DevBuf {  // grow-only device buffer
  T* p = nullptr;
  size_t cap = 0;
  int dev = 0;  // device the current allocation lives on
  ~DevBuf() {
    if (p) {
      cudaFree(p);
      g_devbuf_bytes[dev & 63] -= (long long)(cap * sizeof(T));
    }
  }
  T* ensure(size_t n) {
    if (n > cap) {
      const size_t oldCap = cap;
      if (p) {
        cudaFree(p);
        g_devbuf_bytes[dev & 63] -= (long long)(cap * sizeof(T));
      }
      p = nullptr;
      cap = 0;
      // geometric growth for buffers that creep up, but never more than 256 MB beyond what was asked for
      const size_t want = std::max(n, std::min(oldCap + oldCap / 2, n + (size_t)(256u << 20) / sizeof(T)));
      cudaGetDevice(&dev);
      CUDA_CHECK(cudaMalloc(&p, want * sizeof(T)));
      cap = want;
      g_devbuf_bytes[dev & 63] += (long long)(cap * sizeof(T));
    }
    return p;
  }
  void upload(const std::vector<T>& v, cudaStream_t st) {
    ensure(std::max<size_t>(v.size(), 1));
    if (!v.empty()) CUDA_CHECK(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
  }
};

struct EventTimer {
  cudaEvent_t a, b;
  EventTimer() {
    cudaEventCreate(&a);
    cudaEventCreate(&b);
  }
  ~EventTimer() {
    cudaEventDestroy(a);
    cudaEventDestroy(b);
  }
  void start(cudaStream_t s) { cudaEventRecord(a, s); }
  void stop(cudaStream_t s) { cudaEventRecord(b, s); }
  double ms() {
    float f = 0;
    cudaEventSynchronize(b);
    cudaEventElapsedTime(&f, a, b);
    return f;
  }
};

constexpr int kTraceAttempts = 2;  // band widths K3 tries per launch (W, 2W) before the host reschedules a candidate
constexpr int kPackRunStripes = 4;  // library packing: subjects are laid end to end in runs of at most this many stripes
constexpr int64_t kCheckpointBudget = 48ll << 30;  // K1 column checkpoints: at most this many bytes per pass
constexpr int kSegWindowRounds = 4;                 // K1 column segments: window of the claim order, in rounds of the resident warps
constexpr int64_t kSegRingExtra = 4ll << 30;       // K1 column segments: extra boundary-ring bytes they may cost
constexpr int64_t kSegStateBudget = 2ll << 30;     // K1 column segments: registers of every ticket between two segments (5 KB each)

static double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// -------------------------------------------------------------------------------------------------------------
struct Candidate {
  int q, strand, subject;
  int i0, j0, i1, j1, raw;
};

// Host view of a hot mask (kHotBlocks bits, bit c = column block c).
struct HotMask {
  unsigned long long w[kHotWords] = {};
  static HotMask all() {
    HotMask m;
    for (auto& x : m.w) x = ~0ull;
    return m;
  }
  bool none() const {
    for (auto x : w)
      if (x) return false;
    return true;
  }
  bool test(int c) const { return (w[c >> 6] >> (c & 63)) & 1ull; }
  HotMask operator|(const HotMask& o) const {
    HotMask m;
    for (int k = 0; k < kHotWords; k++) m.w[k] = w[k] | o.w[k];
    return m;
  }
  HotMask operator&(const HotMask& o) const {
    HotMask m;
    for (int k = 0; k < kHotWords; k++) m.w[k] = w[k] & o.w[k];
    return m;
  }
  static HotMask range(int lo, int hi) {  // bits lo..hi
    HotMask m;
    for (int c = std::max(lo, 0); c <= hi && c < kHotBlocks; c++) m.w[c >> 6] |= 1ull << (c & 63);
    return m;
  }
  HotMask shr1() const {  // bit c -> c - 1
    HotMask m;
    for (int k = 0; k < kHotWords; k++) m.w[k] = (w[k] >> 1) | (k + 1 < kHotWords ? w[k + 1] << 63 : 0ull);
    return m;
  }
  HotMask shl1() const {  // bit c -> c + 1
    HotMask m;
    for (int k = 0; k < kHotWords; k++) m.w[k] = (w[k] << 1) | (k > 0 ? w[k - 1] >> 63 : 0ull);
    return m;
  }
  int lowest() const {
    for (int k = 0; k < kHotWords; k++)
      if (w[k]) return 64 * k + __builtin_ctzll(w[k]);
    return -1;
  }
  int highest() const {
    for (int k = kHotWords - 1; k >= 0; k--)
      if (w[k]) return 64 * k + 63 - __builtin_clzll(w[k]);
    return -1;
  }
};

}  // namespace rmgpu

using namespace rmgpu;

struct rmgpu_session {
  int device = 0;
  int numSMs = 0;
  cudaStream_t stream = nullptr;
  static constexpr int R = 16;  // subject rows per lane in K1

  ScoreMatrix matrix;
  bool haveMatrix = false;
  SeqSet lib, queries;
  bool haveLib = false, haveQueries = false;

  // ---- library layout
  int nGroups = 0, nStripes = 0;
  std::vector<int32_t> h_gsubject, h_subjFirstGroup;
  DevBuf<uint2> d_ck;
  bool ckEnabled = false;
  DevBuf<uint8_t> d_rowcode, d_gflags, d_sflags, d_scodes;
  DevBuf<uint2> d_prof;
  DevBuf<int32_t> d_gsubject, d_slen;
  std::vector<uint8_t> h_sflags;
  // One K1 launch: its ticket order (depends on the number of query records), boundary ring, counters and stream.
  struct ScoreSlot {
    DevBuf<int32_t> d_order, d_prodPos, d_consPos;
    int orderGroup = 0;   // max position distance between a stripe and the stripe below it in the ticket order
    int orderForNQ = -1;  // batch size the order was built for
    DevBuf<unsigned long long> d_bnd, d_dbg, d_trace;
    size_t traceWords = 0;
    DevBuf<int> d_finished, d_ticket, d_abort, d_segDone;
    DevBuf<uint16_t> d_level;
    DevBuf<uint32_t> d_segState;
    int maxLevel = 0;     // deepest stripe of a chain
    cudaStream_t stream = nullptr;
    cudaEvent_t evStart = nullptr, evStop = nullptr;
    int q0 = 0, nQ = 0, grid = 0, nslot = 0;
    bool dbg = false;
    // the ring must be all zero (tag 0 = never written) when a launch starts: it is cleared right after a launch has been
    // collected, on the slot's stream, i.e. behind the origin pass and the traceback of the same step
    const void* cleanPtr = nullptr;
    size_t cleanBytes = 0, usedBytes = 0;
  } slot[1];
  void build_order(ScoreSlot& sl, int nQ, int warps);
  void launch_score(const rmgpu_params& p, ScoreSlot& sl, int q0, int nQ, bool exact);
  void collect_score(ScoreSlot& sl, std::vector<int32_t>& pairbest, std::vector<unsigned long long>& hotmask);
  void prepare_score(const rmgpu_params& p, bool exact);
  DevBuf<int64_t> d_soff;
  std::vector<int64_t> h_soff;
  bool profileValid = false;

  // ---- query layout
  std::vector<int64_t> h_qoff, h_selOff;
  std::vector<int32_t> h_qlen, h_selLen;
  int maxQ = 0;
  int64_t selTotal = 0;
  DevBuf<uint8_t> d_qpacked, d_codes_fwd, d_codes_rc;
  DevBuf<int64_t> d_qoff, d_selOff;
  DevBuf<int32_t> d_qlen, d_selLen;
  DevBuf<uint16_t> d_sel;
  uint8_t class_of_code[16] = {};
  int8_t class_score[16][8] = {};
  bool classesExact = true;
  std::vector<uint8_t> classKey;  // to detect class-map changes
  int64_t canonCounts[16] = {};   // residues of the batch (both strands) per canonical code
  bool streamsValid = false;      // sel / code streams and class map match the current matrix and batch
  void prepare_streams();
  int64_t h2d_bytes_library = 0;
  int64_t h2d_bytes_queries = 0;
  double ms_h2d_queries = 0;

  // ---- scratch
  DevBuf<int32_t> d_pairbest;
  DevBuf<unsigned long long> d_hotmask;
  DevBuf<int32_t> d_hotShift;
  DevBuf<unsigned long long> d_dbgC;
  // second stream + scratch of the stripe-pipeline fallback, which runs concurrently with the block-parallel path
  cudaStream_t stream2 = nullptr;
  DevBuf<OriginTask> d_tasksC;
  DevBuf<int32_t> d_ticketTaskC;
  DevBuf<OriginBoundary> d_bndC;
  DevBuf<int> d_progressC, d_finishedC, d_ticketC, d_overflowC, d_abortC, d_leftFlagC;
  DevBuf<unsigned long long> d_keysC, d_valsC;
  // Scratch of one block-parallel origin launch (K2b + resolve).  Two sets: the common triples run from the calling
  // thread on `stream`, the triples whose int16 values wrapped run -- behind their int32 score pass -- from the second
  // host thread on `stream2`, and the two launches share the machine.
  struct BlockBufs {
    DevBuf<BlockTask> d_btasks;
    DevBuf<uint32_t> d_snap;
    DevBuf<int32_t> d_ticketTask, d_ticketStripe, d_prodTicket, d_consTicket;
    DevBuf<OriginBoundary> d_bnd2;
    DevBuf<int> d_progress, d_finished, d_ticket, d_overflow, d_abort, d_redo;
    DevBuf<unsigned long long> d_keys, d_vals, d_pbest, d_dbg, d_fkeys, d_fvals;
  } bb[2];
  // int32 score pass of the wrapped triples (swg_score32_kernel)
  DevBuf<WideTask> d_wtasks;
  DevBuf<int32_t> d_ticketTaskW;
  DevBuf<int2> d_bndW, d_ck32;
  DevBuf<int> d_progressW, d_finishedW, d_ticketW, d_abortW, d_edgeW;
  DevBuf<unsigned long long> d_hot32;
  int32_t h_mat[16][16];
  void check_abort_on(const char* what, int* flag, cudaStream_t st);
  std::vector<int32_t> h_hotShift;
  DevBuf<TraceCand> d_cand;
  DevBuf<uint8_t> d_dir, d_ops;
  DevBuf<int32_t> d_rows, d_opsLen, d_status;

  void set_library(SeqSet s);
  void set_queries(SeqSet s);
  void mask_queries(int32_t n, const int32_t* qi, const int32_t* qstart, const int32_t* qend);
  DevBuf<int32_t> d_maskRanges;
  void build_classes_and_profile();
  void run_score(const rmgpu_params& p, std::vector<int32_t>& pairbest, std::vector<unsigned long long>& hotmask,
                 double& ms, int& launches, bool exact);  // the whole batch in one launch, waited for
  int ovfThresh = 32000;  // K1 reports 32767 for any best at or above this
  int scoreSlack = 0;  // how much below the true best the sampled K1 best can be (0 in exact mode)
  void run_origin(const rmgpu_params& p, const std::vector<int32_t>& pairbest, const std::vector<unsigned long long>& hotmask,
                  int q0, int q1, std::vector<Candidate>& cands, rmgpu_stats& st);  // query records [q0, q1)
  void run_traceback(const rmgpu_params& p, const std::vector<Candidate>& cands, rmgpu_stats& st);
  void run_finalize(const rmgpu_params& p, const std::vector<Candidate>& cands, std::vector<HitDev>& hits,
                    std::vector<char>& rows, rmgpu_stats& st);
  int respell_exotic(const HitDev& h, char* qrow, char* srow) const;
  std::vector<TraceCand> h_candAll;
  DevBuf<TraceCand> d_candAll;
  int64_t opsBytes = 0;
  std::vector<double> h_logtab;
  DevBuf<double> d_logtab;
  DevBuf<char> d_rowsOut, d_rowsPacked;
  DevBuf<HitDev> d_hits;
  DevBuf<int32_t> d_qFirst, d_qCount, d_rankOrder, d_nKept;
  DevBuf<int64_t> d_coverOff;
  DevBuf<uint32_t> d_cover;
  DevBuf<long long> d_rowsTotal;
  double ms_finalize = 0;
  rmgpu_result* search(const rmgpu_params& p);        // whole resident batch, in memory-bounded chunks of queries
  rmgpu_result* search_batch(const rmgpu_params& p);  // one chunk: K1 -> K2 -> K3 -> host post
  int query_capacity() const;
  long long freeAtOpen = 0, devbufAtOpen = 0;  // free HBM and this process's DevBuf bytes when the session opened
};

// -------------------------------------------------------------------------------------------------------------
// Library packing: subjects -> groups of R rows -> stripes of 32 groups.  Subjects are laid end to end in "runs": a run
// starts at lane 0 of a fresh stripe and every subject in it starts at the group after the previous one, so a stripe
// may hold the tail of one subject and the head of the next (the kernel resets the column state at a group that
// starts a subject; the boundary it reads at lane 0 and the one it writes at lane 31 then belong to different
// subjects).  Runs are filled best-fit-decreasing up to kPackRunStripes stripes, which leaves < 1 % of padded rows
// instead of the ~8 % of one-subject-per-chain packing, while the dependency chains (consecutive stripes that
// share a subject) stay short.  A subject longer than a run is a run of its own; the free lanes of its last
// stripe are filled with subjects that fit entirely.
// -------------------------------------------------------------------------------------------------------------
void rmgpu_session::set_library(SeqSet s) {
  lib = std::move(s);
  lib.mark_exotic();
  const CanonTables& ct = canon();
  const int n = (int)lib.seqs.size();
  std::vector<int> groups(n);
  std::multimap<int, int> pool;  // groups -> subject, the subjects not yet placed
  for (int i = 0; i < n; i++) {
    if (lib.seqs[i].size() >= (1u << 20)) throw Error("subject longer than 1,048,575 bases: " + lib.names[i]);
    groups[i] = (int)((lib.seqs[i].size() + R - 1) / R);
    if (groups[i] > 0) pool.insert({groups[i], -i});  // -i: among equal sizes the lowest index is the last element
  }
  const int cap = kPackRunStripes * 32;
  struct Place {
    int subject, firstGroup;
  };
  std::vector<Place> places;
  int stripes = 0;
  while (!pool.empty()) {
    auto last = std::prev(pool.end());  // largest subject left opens a run
    int total = last->first;
    places.push_back({-last->second, stripes * 32});
    pool.erase(last);
    for (;;) {
      // room: up to the run capacity, or (oversize subject) up to the end of its last stripe
      const int room = total <= cap ? cap - total : (32 - total % 32) % 32;
      auto it = pool.upper_bound(room);  // first subject that does NOT fit
      if (it == pool.begin()) break;
      --it;
      places.push_back({-it->second, stripes * 32 + total});
      total += it->first;
      pool.erase(it);
    }
    stripes += (total + 31) / 32;
  }
  nStripes = std::max(stripes, 1);
  nGroups = nStripes * 32;
  std::vector<uint8_t> rowcode((size_t)nGroups * R, 0xFF), gflags(nGroups, 1), sflags(nStripes, 0);
  h_gsubject.assign(nGroups, -1);
  h_subjFirstGroup.assign(n, -1);
  parallel_for((int)places.size(), [&](int pi) {  // subjects own disjoint groups
    const Place& pl = places[pi];
    h_subjFirstGroup[pl.subject] = pl.firstGroup;
    const std::string& sq = lib.seqs[pl.subject];
    for (int gi = 0; gi < groups[pl.subject]; gi++) {
      int g = pl.firstGroup + gi;
      h_gsubject[g] = pl.subject;
      gflags[g] = gi == 0 ? 1 : 0;
      for (int r = 0; r < R; r++) {
        size_t pos = (size_t)gi * R + r;
        if (pos < sq.size()) rowcode[(size_t)g * R + r] = ct.of_letter[(unsigned char)sq[pos]];
      }
    }
  }, 32);
  for (int k = 0; k < nStripes; k++) {
    bool prev = !(gflags[k * 32] & 1);
    if (prev) {
      sflags[k] |= 1;
      sflags[k - 1] |= 2;
    }
  }
  h_sflags = sflags;
  slot[0].orderForNQ = -1;
  // unpacked subject codes for K2 / K3
  h_soff.assign(n + 1, 0);
  std::vector<int32_t> slen(n);
  for (int i = 0; i < n; i++) {
    slen[i] = (int32_t)lib.seqs[i].size();
    h_soff[i + 1] = h_soff[i] + slen[i];
  }
  std::vector<uint8_t> scodes((size_t)std::max<int64_t>(h_soff[n], 1));
  parallel_for(n, [&](int i) {
    for (size_t k = 0; k < lib.seqs[i].size(); k++) scodes[h_soff[i] + k] = ct.of_letter[(unsigned char)lib.seqs[i][k]];
  }, 32);
  h2d_bytes_library = (int64_t)rowcode.size() + gflags.size() + sflags.size() + h_gsubject.size() * 4 + scodes.size() +
                      h_soff.size() * 8 + slen.size() * 4;
  d_rowcode.upload(rowcode, stream);
  d_gflags.upload(gflags, stream);
  d_sflags.upload(sflags, stream);
  d_gsubject.upload(h_gsubject, stream);
  d_scodes.upload(scodes, stream);
  d_soff.upload(h_soff, stream);
  d_slen.upload(slen, stream);
  CUDA_CHECK(cudaStreamSynchronize(stream));
  haveLib = true;
  profileValid = false;
}

void rmgpu_session::set_queries(SeqSet s) {
  queries = std::move(s);
  queries.mark_exotic();
  const CanonTables& ct = canon();
  const int n = (int)queries.seqs.size();
  h_qoff.assign(n + 1, 0);
  h_selOff.assign(n + 1, 0);
  h_qlen.resize(n);
  h_selLen.resize(n);
  h_hotShift.clear();
  maxQ = 0;
  for (int i = 0; i < n; i++) {
    int64_t Q = (int64_t)queries.seqs[i].size();
    if (Q >= (1 << 20)) throw Error("query longer than 1,048,575 bases (fragment it: SimpleBatcher.pm:592-800): " + queries.names[i]);
    h_qlen[i] = (int32_t)Q;
    h_selLen[i] = (int32_t)((((Q + 31) >> 5) + 1) * 32);
    h_qoff[i + 1] = h_qoff[i] + ((Q + 1) & ~1ll);  // even nibble offsets
    h_selOff[i + 1] = h_selOff[i] + h_selLen[i];
    maxQ = std::max<int>(maxQ, (int)Q);
    int sh = 0;
    while ((((Q + 31) >> 5) + 1) >> sh > kHotBlocks - 1) sh++;  // at most kHotBlocks column blocks per query (hot mask bits)
    h_hotShift.push_back(sh);
  }
  selTotal = h_selOff[n];
  std::vector<uint8_t> packed((size_t)(((h_qoff[n] / 2 + 1) + 7) & ~3ll), 0);  // whole 32-bit words (mask_queries works on words)
  for (auto& c : canonCounts) c = 0;
  {
    // every record starts on a byte of `packed` (even nibble offsets), so records are packed independently
    std::vector<std::array<int64_t, 16>> cnt(n);
    parallel_for(n, [&](int i) {
      const std::string& q = queries.seqs[i];
      std::array<int64_t, 16>& cc = cnt[i];
      cc.fill(0);
      for (size_t k = 0; k < q.size(); k++) {
        uint8_t c = ct.of_letter[(unsigned char)q[k]];
        cc[c]++;
        int64_t nib = h_qoff[i] + (int64_t)k;
        packed[nib >> 1] |= (uint8_t)(c << ((nib & 1) * 4));
      }
    }, 1);
    for (int i = 0; i < n; i++)
      for (int c = 0; c < 16; c++) {
        canonCounts[c] += cnt[i][c];
        canonCounts[ct.comp[c]] += cnt[i][c];
      }
  }
  double t0 = now_ms();
  d_qpacked.upload(packed, stream);
  d_qoff.upload(h_qoff, stream);
  d_qlen.upload(h_qlen, stream);
  d_selOff.upload(h_selOff, stream);
  d_selLen.upload(h_selLen, stream);
  d_hotShift.upload(h_hotShift, stream);
  d_sel.ensure((size_t)std::max<int64_t>(selTotal, 1));
  d_codes_fwd.ensure((size_t)std::max<int64_t>(h_qoff[n], 1));
  d_codes_rc.ensure((size_t)std::max<int64_t>(h_qoff[n], 1));
  streamsValid = false;
  haveQueries = true;
  if (haveMatrix) prepare_streams();  // K0 now, so that the copy and the expansion are both inside the upload's time
  CUDA_CHECK(cudaStreamSynchronize(stream));
  ms_h2d_queries = now_ms() - t0;
  h2d_bytes_queries = (int64_t)packed.size() + (int64_t)(n + 1) * 24;
}

// Masks ranges of the resident batch with X on the device and in the host copy (the rows of later hits are spelled from
// it): the engine's half of RepeatMasker's stage chain, which masks what a stage found before the next stage searches
// the same batch against another library (RepeatMasker:4115-4654, :6036-6090).  1-based inclusive coordinates, as in
// rmgpu_hit; ranges may overlap.  The residue statistics, class map and K0 streams are rebuilt before the next search.
void rmgpu_session::mask_queries(int32_t n, const int32_t* qi, const int32_t* qstart, const int32_t* qend) {
  if (!haveQueries) throw Error("mask_queries: no query batch is resident");
  const CanonTables& ct = canon();
  const uint8_t X = ct.of_letter[(unsigned char)'X'];
  std::vector<int32_t> ranges;
  for (int32_t r = 0; r < n; r++) {
    if (qi[r] < 0 || qi[r] >= (int32_t)queries.seqs.size()) throw Error("mask_queries: query index out of range");
    std::string& q = queries.seqs[qi[r]];
    const int64_t p0 = std::max<int64_t>((int64_t)qstart[r] - 1, 0), p1 = std::min<int64_t>((int64_t)qend[r], (int64_t)q.size());
    if (p0 >= p1) continue;
    for (int64_t k = p0; k < p1; k++) {
      const uint8_t c = ct.of_letter[(unsigned char)q[k]];
      if (c != X || q[k] != 'X') {
        canonCounts[c]--;
        canonCounts[ct.comp[c]]--;
        canonCounts[X]++;
        canonCounts[ct.comp[X]]++;
        q[k] = 'X';
      }
    }
    ranges.push_back(qi[r]);
    ranges.push_back((int32_t)p0);
    ranges.push_back((int32_t)p1);
  }
  if (ranges.empty()) return;
  d_maskRanges.upload(ranges, stream);
  MaskArgs a;
  a.qpacked = reinterpret_cast<uint32_t*>(d_qpacked.p);
  a.qoff = d_qoff.p;
  a.ranges = d_maskRanges.p;
  a.n = (int)(ranges.size() / 3);
  a.code = X;
  mask_queries_kernel<<<std::min(a.n, numSMs * 8), 256, 0, stream>>>(a);
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaStreamSynchronize(stream));
  queries.mark_exotic();
  streamsValid = false;
}

// Residue classes of the score pass and the per-column selector / code streams (K0).  Depends on the matrix (which letters
// share a column) and on the batch (which columns occur): rebuilt when either changes.
// Up to 7 exact classes; beyond that the rarest columns share class 6 with row-wise max scores, which makes K1 an upper
// bound (still an exact filter: K2 recomputes every passing pair with the full alphabet).
void rmgpu_session::prepare_streams() {
  if (streamsValid) return;
  if (!haveMatrix || !haveQueries) throw Error("matrix and queries must be set");
  const int n = (int)queries.seqs.size();
  int64_t counts[16] = {};
  for (int c = 0; c < 16; c++) counts[matrix.col_of_canon[c]] += canonCounts[c];
  std::vector<int> present;
  for (int c = 0; c < matrix.n; c++)
    if (counts[c] > 0) present.push_back(c);
  std::stable_sort(present.begin(), present.end(), [&](int a, int b) { return counts[a] > counts[b]; });
  std::vector<uint8_t> key;
  int nExact = present.size() <= 7 ? (int)present.size() : 6;
  classesExact = present.size() <= 7;
  uint8_t class_of_col[16];
  for (int c = 0; c < 16; c++) class_of_col[c] = 6;
  for (int k = 0; k < nExact; k++) class_of_col[present[k]] = (uint8_t)k;
  for (int c = 0; c < 16; c++) class_of_code[c] = class_of_col[matrix.col_of_canon[c]];
  for (int sc = 0; sc < 16; sc++) {  // sc: canonical subject code
    for (int k = 0; k < 8; k++) class_score[sc][k] = -128;
    const int scol = matrix.col_of_canon[sc];
    auto clampv = [&](int v) {
      if (v > 127) throw Error("matrix score > 127 is not supported by the int8 profile");
      if (v < -128) classesExact = false;  // clamped: K1 becomes an upper bound, its checkpoints are not exact
      return (int8_t)std::max(v, -128);
    };
    for (int k = 0; k < nExact; k++) class_score[sc][k] = clampv(matrix.score[scol][present[k]]);
    if (!classesExact && present.size() > 7) {
      int mx = -128;
      for (size_t k = 6; k < present.size(); k++) mx = std::max(mx, matrix.score[scol][present[k]]);
      class_score[sc][6] = clampv(mx);
    }
  }
  for (int c = 0; c < 16; c++) key.push_back(class_of_code[c]);
  for (int sc = 0; sc < 16; sc++)
    for (int k = 0; k < 8; k++) key.push_back((uint8_t)class_score[sc][k]);
  if (key != classKey) {
    classKey = key;
    profileValid = false;
  }
  if (n > 0) {
    StreamArgs a;
    a.qpacked = d_qpacked.p;
    a.qoff = d_qoff.p;
    a.qlen = d_qlen.p;
    a.selOff = d_selOff.p;
    a.selLen = d_selLen.p;
    a.sel = d_sel.p;
    a.codes_fwd = d_codes_fwd.p;
    a.codes_rc = d_codes_rc.p;
    a.nQueries = n;
    memcpy(a.class_of_code, class_of_code, 16);
    memcpy(a.comp_code, canon().comp, 16);
    dim3 grid((unsigned)std::min<int64_t>(((int64_t)maxQ + 64 + 255) / 256, 1024), (unsigned)n);
    build_streams_kernel<<<grid, 256, 0, stream>>>(a);
    CUDA_CHECK(cudaGetLastError());
  }
  for (int i = 0; i < 16; i++)
    for (int j = 0; j < 16; j++) h_mat[i][j] = matrix.score[matrix.col_of_canon[i]][matrix.col_of_canon[j]];
  streamsValid = true;
}

void rmgpu_session::build_classes_and_profile() {
  if (profileValid) return;
  ProfileArgs a;
  a.rowcode = d_rowcode.p;
  a.prof = d_prof.ensure((size_t)nGroups * R);
  a.nRows = (int64_t)nGroups * R;
  memcpy(a.class_score, class_score, sizeof(class_score));
  build_profile_kernel<<<std::min<int64_t>((a.nRows + 255) / 256, 4096), 256, 0, stream>>>(a);
  CUDA_CHECK(cudaGetLastError());
  profileValid = true;
}

static uint32_t pack2(int v) { return ((uint32_t)(uint16_t)(int16_t)v) | ((uint32_t)(uint16_t)(int16_t)v << 16); }

// Ticket order of the stripes.  Ticket t = position * nQ + query: a "level" of the order spans all queries of the
// batch.  Chains (a subject spanning several stripes) are dealt round-robin into groups; inside a group positions go
// level by level (all first stripes, then all second stripes, ...).  With group sizes chosen so that one level holds
// a few times more tickets than there are resident warps, a consumer stripe is claimed long after its producer
// finished (no polling in steady state), while the producer->consumer distance stays bounded by the group size,
// which bounds the boundary ring.
void rmgpu_session::build_order(ScoreSlot& sl, int nQ, int warps) {
  if (sl.orderForNQ == nQ) return;
  struct Chain {
    int first, len;
  };
  std::vector<Chain> chains;
  for (int k = 0; k < nStripes; k++) {
    if (h_sflags[k] & 1)
      chains.back().len++;
    else
      chains.push_back({k, 1});
  }
  std::stable_sort(chains.begin(), chains.end(), [](const Chain& a, const Chain& b) { return a.len > b.len; });
  // aim at ~3 x warps tickets per level: chainsPerGroup * nQ ~ 3 * warps
  const int chainsPerGroup = std::max(8, (3 * warps + nQ - 1) / std::max(nQ, 1));
  const int nGroupsT = std::max(1, ((int)chains.size() + chainsPerGroup - 1) / chainsPerGroup);
  std::vector<std::vector<Chain>> grp(nGroupsT);
  for (size_t c = 0; c < chains.size(); c++) grp[c % nGroupsT].push_back(chains[c]);
  std::vector<int32_t> order, posOf(nStripes, -1);
  std::vector<uint16_t> levelOf;
  sl.orderGroup = 0;
  sl.maxLevel = 0;
  for (auto& gchains : grp) {
    int tot = 0, maxLen = 0;
    for (auto& c : gchains) {
      tot += c.len;
      maxLen = std::max(maxLen, c.len);
    }
    sl.orderGroup = std::max(sl.orderGroup, tot);
    for (int level = 0; level < maxLen; level++)
      for (auto& c : gchains)
        if (level < c.len) {
          posOf[c.first + level] = (int32_t)order.size();
          order.push_back(c.first + level);
          levelOf.push_back((uint16_t)std::min(level, 65535));
          sl.maxLevel = std::max(sl.maxLevel, level);
        }
  }
  std::vector<int32_t> prodPos(nStripes, -1), consPos(nStripes, -1);
  for (int pos = 0; pos < nStripes; pos++) {
    int k = order[pos];
    if (h_sflags[k] & 1) prodPos[pos] = posOf[k - 1];
    if (h_sflags[k] & 2) consPos[pos] = posOf[k + 1];
  }
  sl.d_order.upload(order, sl.stream);
  sl.d_prodPos.upload(prodPos, sl.stream);
  sl.d_consPos.upload(consPos, sl.stream);
  sl.d_level.upload(levelOf, sl.stream);
  sl.orderForNQ = nQ;
}

void rmgpu_session::check_abort_on(const char* what, int* flag, cudaStream_t stq) {
  int f = 0;
  CUDA_CHECK(cudaMemcpyAsync(&f, flag, sizeof(int), cudaMemcpyDeviceToHost, stq));
  CUDA_CHECK(cudaStreamSynchronize(stq));
  if (f) throw Error(std::string(what) + ": stripe pipeline watchdog fired (a producer/consumer wait exceeded its bound)");
}

// K1 over the resident batch x library.
// prepare_score: what every launch of the batch shares (class map / streams / profile, thresholds, result and checkpoint
// buffers); launch_score: one launch over the query records [q0, q0 + nQ) on the slot's own stream; collect_score: wait for
// it and bring its best scores and hot masks to the host.
void rmgpu_session::prepare_score(const rmgpu_params& p, bool exact) {
  const int nQ = (int)queries.seqs.size(), nS = (int)lib.seqs.size();
  if (p.gap_init >= 0 || p.ins_gap_ext > 0 || p.del_gap_ext > 0 || p.gap_init < -20000)
    throw Error("gap penalties must be negative");
  prepare_streams();
  build_classes_and_profile();
  scoreSlack = exact ? 0 : 7 * std::abs(p.ins_gap_ext);
  // between two samples (8 columns) a score can rise by at most 8 * max matrix score
  ovfThresh = 32767 - 9 * std::max(matrix.max_score(), 1) - scoreSlack;
  if (ovfThresh < 16384) throw Error("matrix scores too large for the int16 score pass");
  const size_t nPair = (size_t)nQ * nS * 2;
  d_pairbest.ensure(std::max<size_t>(nPair, 1));
  d_hotmask.ensure(std::max<size_t>(nPair * kHotWords, 1));
  // column checkpoints for the origin pass (kHotBlocks per query): only when they fit the budget
  static const bool noCk = getenv("RMGPU_NO_CHECKPOINTS") != nullptr;
  const int64_t ckElems = (int64_t)nQ * kHotBlocks * nGroups * R;
  ckEnabled = !noCk && ckElems * (int64_t)sizeof(uint2) <= kCheckpointBudget;
  if (ckEnabled) d_ck.ensure((size_t)ckElems);
  CUDA_CHECK(cudaMemsetAsync(d_pairbest.p, 0, nPair * sizeof(int32_t), stream));
  CUDA_CHECK(cudaMemsetAsync(d_hotmask.p, 0, nPair * kHotWords * sizeof(unsigned long long), stream));
  CUDA_CHECK(cudaStreamSynchronize(stream));  // K0 and the cleared buffers are in place before any slot's stream starts
}

void rmgpu_session::launch_score(const rmgpu_params& p, ScoreSlot& sl, int q0, int nQ, bool exact) {
  const int nS = (int)lib.seqs.size();
  // the instantiation: gap penalties of RepeatMasker's crossmatch recipes as immediates, anything else from registers
  using ScoreKernel = void (*)(ScoreArgs);
  const int gi = p.gap_init, ie = p.ins_gap_ext, df = std::max(p.del_gap_ext, p.gap_init);
  static const bool noImm = getenv("RMGPU_K1_NO_IMMEDIATES") != nullptr;
  ScoreKernel kern = exact ? (ScoreKernel)swg_score_kernel<R, true> : (ScoreKernel)swg_score_kernel<R, false>;
  if (!exact && !noImm) {
    // the (gap_init, ins_gap_ext, del_gap_ext) of the crossmatch recipes (RepeatMasker:2127-2480; all but the low-complexity
    // stage's -10/-3/-3, which runs the generic instantiation)
#define RMGPU_K1_IMM(G, I, D) \
  if (gi == (G) && ie == (I) && df == (D)) kern = swg_score_kernel<R, false, (G), (I), (D)>;
    RMGPU_K1_IMM(-30, -6, -5)
    RMGPU_K1_IMM(-27, -6, -5)
    RMGPU_K1_IMM(-35, -7, -6)
    RMGPU_K1_IMM(-33, -5, -4)
    RMGPU_K1_IMM(-27, -5, -5)
    RMGPU_K1_IMM(-40, -15, -15)
    RMGPU_K1_IMM(-35, -10, -10)
#undef RMGPU_K1_IMM
  }
  int blocksPerSM = 0;
  CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSM, kern, kScoreWarps * 32, 0));
  if (blocksPerSM < 1) throw Error("K1 does not fit on an SM");
  const int64_t nTickets = (int64_t)nQ * nStripes;
  if (nTickets >= (1ll << 31)) throw Error("too many stripe tasks in one batch");
  const int resident = numSMs * blocksPerSM;
  // ticket order and ring are sized for the warps that can run at once
  const int warps = (int)std::min<int64_t>((int64_t)resident * kScoreWarps, nTickets);
  build_order(sl, nQ, warps);
  // ring slots: a slot is rewritten nslot tickets later; its reader is at most orderGroup * nQ tickets after its writer
  // ... and a boundary's 4-bit tag (generation = ticket / nslot, 15 values) must be unique within a launch: a slot keeps
  // the boundary of its last WRITER, which may be many generations old when the tickets in between had no stripe below
  // column segments (kernels.cuh, swg_score_kernel): tickets dealt in quarters when there are several waves of them, so
  // that the last wave does not run half empty; RMGPU_K1_SEGMENTS=1 switches them off, =n sets another count
  static const char* segEnv = getenv("RMGPU_K1_SEGMENTS");
  int nSeg = segEnv ? std::max(1, atoi(segEnv)) : 4;
  const int segMinBlocks = nSeg * (sl.maxLevel + 8);  // every segment of a split record keeps >= 8 blocks at any chain depth
  if ((!segEnv && nTickets <= warps) || nTickets * nSeg >= (1ll << 31) || sl.maxLevel >= 65535 ||
      (((int64_t)maxQ + 31) >> 5) + 1 < segMinBlocks ||
      nTickets * (int64_t)(kSegStateWords * 32 * sizeof(uint32_t)) > kSegStateBudget)
    nSeg = 1;
  // windows of the claim order: as equal as possible, each at least kSegWindowRounds rounds of the resident warps.  The
  // warps of a sub-partition do not progress evenly (a quarter ticket takes 3 ms on one warp and 15-25 ms on a starved
  // neighbour, profiles/r2_k1_segments.txt): a segment's predecessor must have that long before the segment is claimed
  static const char* winEnv = getenv("RMGPU_K1_SEGWIN");
  const int64_t winRounds = winEnv ? std::max(1, atoi(winEnv)) : kSegWindowRounds;
  const int64_t nWindows = std::max<int64_t>(1, nTickets / (winRounds * (int64_t)warps));
  const int segWindow = (int)((nTickets + nWindows - 1) / nWindows);
  // ring slots: with segments the tickets of a window run interleaved, so the old reader of a slot must lie a whole
  // window back for it to have been claimed first (2 windows + the writer-reader distance)
  const int64_t bndStride = (((int64_t)maxQ + 31) >> 5) * 32 + 32;
  const int64_t wholeSlots = std::min<int64_t>(
      nTickets, std::max<int64_t>(std::max<int64_t>(2 * (int64_t)warps + 2, 2 * (int64_t)sl.orderGroup * nQ + 2), nTickets / 14 + 1));
  int64_t segSlots = nSeg > 1 ? std::min<int64_t>(nTickets, 2 * (int64_t)segWindow + (int64_t)sl.orderGroup * nQ + 2) : 0;
  // ... unless that ring would be much larger than the one the pass was planned with (query_capacity): then whole tickets
  if (segSlots > wholeSlots && (segSlots - wholeSlots) * bndStride * 8 > kSegRingExtra) {
    nSeg = 1;
    segSlots = 0;
  }
  const int nslot = (int)std::max<int64_t>(wholeSlots, segSlots);
  if ((nTickets + nslot - 1) / nslot > 15) throw Error("K1 boundary ring: more than 15 tag generations in one launch");
  ScoreArgs a;
  a.prof = d_prof.p;
  a.gflags = d_gflags.p;
  a.gsubject = d_gsubject.p;
  a.sflags = d_sflags.p;
  a.order = sl.d_order.p;
  a.prodPos = sl.d_prodPos.p;
  a.consPos = sl.d_consPos.p;
  a.sel = d_sel.p;
  a.selOff = d_selOff.p + q0;  // offsets into sel are absolute
  a.qlen = d_qlen.p + q0;
  a.hotShift = d_hotShift.p + q0;
  a.nQueries = nQ;
  a.nStripes = nStripes;
  a.nSubjects = nS;
  a.pairbest = d_pairbest.p + (size_t)q0 * nS * 2;
  a.hotmask = d_hotmask.p + (size_t)q0 * nS * 2 * kHotWords;
  a.bnd = sl.d_bnd.ensure((size_t)nslot * bndStride);
  a.bndStride = bndStride;
  a.nslot = nslot;
  a.finished = sl.d_finished.ensure((size_t)nTickets);
  a.ticket = sl.d_ticket.ensure(1);
  a.gi2 = pack2(p.gap_init);
  a.ie2 = pack2(p.ins_gap_ext);
  a.df2 = pack2(std::max(p.del_gap_ext, p.gap_init));
  a.minscore = std::max(1, std::min(p.minscore, ovfThresh) - scoreSlack);
  a.ovfThresh = ovfThresh;
  a.abort_flag = sl.d_abort.ensure(1);
  a.nSeg = nSeg;
  a.segWindow = segWindow;
  a.segMinBlocks = segMinBlocks;
  a.level = sl.d_level.p;
  a.segState = nullptr;
  a.segDone = nullptr;
  if (nSeg > 1) {
    a.segState = sl.d_segState.ensure((size_t)nTickets * kSegStateWords * 32);
    a.segDone = sl.d_segDone.ensure((size_t)nTickets);
    CUDA_CHECK(cudaMemsetAsync(a.segDone, 0, (size_t)nTickets * sizeof(int), sl.stream));
  }
  static const char* traceEnv = getenv("RMGPU_K1_TRACE");
  a.trace = nullptr;
  sl.traceWords = 0;
  if (traceEnv) {
    sl.traceWords = (size_t)nTickets * nSeg * 4;
    a.trace = sl.d_trace.ensure(sl.traceWords);
    CUDA_CHECK(cudaMemsetAsync(a.trace, 0, sl.traceWords * 8, sl.stream));
  }
  a.ck = ckEnabled ? d_ck.p + (size_t)q0 * kHotBlocks * ((size_t)nGroups * R) : nullptr;
  a.ckRows = (int64_t)nGroups * R;
  static const bool dbgOn = getenv("RMGPU_DEBUG_K1") != nullptr;
  sl.dbg = dbgOn;
  a.dbg = nullptr;
  if (dbgOn) {
    a.dbg = sl.d_dbg.ensure(8);
    CUDA_CHECK(cudaMemsetAsync(a.dbg, 0, 64, sl.stream));
  }
  const int grid = (int)std::min<int64_t>(resident, (nTickets * nSeg + kScoreWarps - 1) / kScoreWarps);  // one persistent wave
  sl.q0 = q0;
  sl.nQ = nQ;
  sl.grid = grid;
  sl.nslot = nslot;
  CUDA_CHECK(cudaMemsetAsync(a.abort_flag, 0, sizeof(int), sl.stream));
  CUDA_CHECK(cudaMemsetAsync(a.finished, 0, (size_t)nTickets * sizeof(int), sl.stream));
  CUDA_CHECK(cudaMemsetAsync(a.ticket, 0, sizeof(int), sl.stream));
  // tag 0 = "never written": the boundary ring must not hold tags of an earlier launch
  sl.usedBytes = (size_t)nslot * bndStride * sizeof(unsigned long long);
  if (sl.cleanPtr != (const void*)a.bnd || sl.cleanBytes < sl.usedBytes)
    CUDA_CHECK(cudaMemsetAsync(a.bnd, 0, sl.usedBytes, sl.stream));
  sl.cleanPtr = nullptr;
  sl.cleanBytes = 0;
  CUDA_CHECK(cudaEventRecord(sl.evStart, sl.stream));
  kern<<<grid, kScoreWarps * 32, 0, sl.stream>>>(a);
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaEventRecord(sl.evStop, sl.stream));
}

void rmgpu_session::collect_score(ScoreSlot& sl, std::vector<int32_t>& pairbest, std::vector<unsigned long long>& hotmask) {
  const int nS = (int)lib.seqs.size();
  const size_t off = (size_t)sl.q0 * nS * 2, n = (size_t)sl.nQ * nS * 2;
  CUDA_CHECK(cudaMemcpyAsync(pairbest.data() + off, d_pairbest.p + off, n * sizeof(int32_t), cudaMemcpyDeviceToHost, sl.stream));
  CUDA_CHECK(cudaMemcpyAsync(hotmask.data() + off * kHotWords, d_hotmask.p + off * kHotWords, n * kHotWords * sizeof(unsigned long long),
                             cudaMemcpyDeviceToHost, sl.stream));
  CUDA_CHECK(cudaStreamSynchronize(sl.stream));
  check_abort_on("K1", sl.d_abort.p, sl.stream);
  // clear the ring for the next launch now: the copy runs beside the rest of this step
  CUDA_CHECK(cudaMemsetAsync(sl.d_bnd.p, 0, sl.usedBytes, sl.stream));
  sl.cleanPtr = sl.d_bnd.p;
  sl.cleanBytes = sl.usedBytes;
  if (sl.traceWords) {  // RMGPU_K1_TRACE=<file>: per sub-ticket (ticket * 16 + segment, claim, start, end in ns), last launch wins
    std::vector<unsigned long long> tr(sl.traceWords);
    CUDA_CHECK(cudaMemcpy(tr.data(), sl.d_trace.p, sl.traceWords * 8, cudaMemcpyDeviceToHost));
    if (FILE* f = fopen(getenv("RMGPU_K1_TRACE"), "wb")) {
      fwrite(tr.data(), 8, tr.size(), f);
      fclose(f);
    }
  }
  if (sl.dbg) {
    unsigned long long d[8];
    float ms = 0;
    cudaEventElapsedTime(&ms, sl.evStart, sl.evStop);
    CUDA_CHECK(cudaMemcpy(d, sl.d_dbg.p, 64, cudaMemcpyDeviceToHost));
    int64_t rows = 0;
    for (auto& sq : lib.seqs) rows += (int64_t)sq.size();
    const int warps = sl.grid * kScoreWarps;
    fprintf(stderr, "[K1 dbg] records %d..%d tasks %llu  producer-wait %.1f%% of task cycles  slot-wait %.2e cyc  segment-wait %.2e cyc  mean task %.0f cyc  grid %d nslot %d\n", sl.q0,
            sl.q0 + sl.nQ - 1, d[2], 100.0 * (double)d[0] / (double)std::max(d[3], 1ull), (double)d[1], (double)d[6], (double)d[3] / (double)std::max(d[2], 1ull), sl.grid, sl.nslot);
    fprintf(stderr, "[K1 dbg] stripes %d  padded rows %.2f%%  warp busy until exit: mean %.1f%% of the last warp's (%.2f ms)  launch %.2f ms\n", nStripes,
            100.0 * ((double)nStripes * 32 * R - (double)rows) / (double)std::max<int64_t>(rows, 1),
            100.0 * (double)d[4] / ((double)warps * (double)std::max(d[5], 1ull)), (double)d[5] * 1e-6, ms);
  }
}

void rmgpu_session::run_score(const rmgpu_params& p, std::vector<int32_t>& pairbest,
                              std::vector<unsigned long long>& hotmask, double& ms, int& launches, bool exact) {
  const int nQ = (int)queries.seqs.size(), nS = (int)lib.seqs.size();
  pairbest.assign((size_t)nQ * nS * 2, 0);
  hotmask.assign((size_t)nQ * nS * 2 * kHotWords, 0);
  ms = 0;
  if (nQ == 0 || nS == 0) return;
  prepare_score(p, exact);
  launch_score(p, slot[0], 0, nQ, exact);
  collect_score(slot[0], pairbest, hotmask);
  float f = 0;
  cudaEventElapsedTime(&f, slot[0].evStart, slot[0].evStop);
  ms = f;
  launches += 2;
}

// K2 over every (query, strand, subject) triple whose K1 best reached minscore.
//   common path  (run_blocks): one task per hot column block, started from K1's column checkpoint, all blocks in
//                parallel, pseudo origins resolved through per-block snapshots (swg_origin_block_kernel);
//   fallback     (run_chain):  one sweep with the stripe pipeline (swg_origin_kernel) for triples whose K1 values
//                wrapped (int16: swept from the checkpoint before their first hot block to their last one), that are
//                too large for the 31-bit relative origin or hold merged residue classes, or when checkpoints are
//                disabled (those from column 0); runs concurrently on a second stream.
void rmgpu_session::run_origin(const rmgpu_params& p, const std::vector<int32_t>& pairbest,
                               const std::vector<unsigned long long>& hotmask, int q0, int q1, std::vector<Candidate>& cands,
                               rmgpu_stats& st) {
  const int nS = (int)lib.seqs.size();
  static const bool fullMode = getenv("RMGPU_K2_FULL") != nullptr;
  static const bool noWide = getenv("RMGPU_NO_WIDE") != nullptr;  // wrapped triples take the one-sweep fallback (the round-1 path)
  struct Triple {
    int q, strand, subject;
    HotMask dil;
    bool chain;
    // int16 wrap in K1: an int32 score pass (K2w) over the window [lo32, hi32] of column blocks supplies exact
    // checkpoints and the exact hot mask, then the triple's hot blocks are K2b tasks like everyone else's
    bool wide = false;
    int lo32 = 0, hi32 = -1;
    int64_t ck32Base = -1;
    HotMask dil32;  // exact mask of the window (dilated like dil)
  };
  std::vector<Triple> triples;
  for (int q = q0; q < q1; q++)
    for (int s = 0; s < nS; s++)
      for (int strand = 0; strand < 2; strand++) {
        const size_t idx = ((size_t)q * nS + s) * 2 + strand;
        if (pairbest[idx] < std::max(1, std::min(p.minscore, ovfThresh) - scoreSlack)) continue;
        const int64_t Q = h_qlen[q], L = (int64_t)lib.seqs[s].size();
        HotMask m;
        for (int k = 0; k < kHotWords; k++) m.w[k] = hotmask[idx * kHotWords + k];
        Triple t;
        t.q = q;
        t.strand = strand;
        t.subject = s;
        // K1 files a cell under the block its warp is in, which for lanes lagging behind is the block AFTER the cell's
        // own, and an alignment may start before the block in which it first reaches minscore: add the block before
        t.dil = m | m.shr1();
        const int64_t CB = 32ll << h_hotShift[q];
        // checkpoints are exact only when K1 is (no merged residue class, no int16 wrap)
        t.chain = pairbest[idx] >= 32767 || fullMode || !ckEnabled || !classesExact || t.dil.none() ||
                  (CB + 2) * (L + 1) >= (1ll << 31);
        t.wide = !noWide && pairbest[idx] >= 32767 && !fullMode && ckEnabled && classesExact && !t.dil.none() &&
                 (CB + 2) * (L + 1) < (1ll << 31) && (Q + CB - 1) / CB <= kHotBlocks;
        if (t.wide) t.chain = false;
        triples.push_back(t);
      }
  st.n_pairs_passed += (int)triples.size();
  cands.clear();
  if (triples.empty()) return;
  // the candidate tables carry the triple and the raw score in 24-bit fields (cand_val, resolve_candidates_kernel)
  if (triples.size() >= (1u << 24)) throw Error("origin pass: more than 2^24 passing (query, strand, subject) triples in one pass");
  {
    int64_t maxL = 0;
    for (auto& sq : lib.seqs) maxL = std::max<int64_t>(maxL, (int64_t)sq.size());
    if (std::min<int64_t>(maxQ, maxL) * std::max(matrix.max_score(), 1) >= (1ll << 24))
      throw Error("origin pass: raw scores of 2^24 or more are possible with these sequence lengths");
  }
  std::vector<std::pair<int, Candidate>> found;  // (triple, candidate)
  auto decode_val = [](unsigned long long v, Candidate& c) {
    c.raw = (int)(v >> 40);
    c.i1 = (int)(0xFFFFFu - (uint32_t)((v >> 20) & 0xFFFFFu));
    c.j1 = (int)(0xFFFFFu - (uint32_t)(v & 0xFFFFFu));
  };
  static const bool dbgK2 = getenv("RMGPU_DEBUG_K2") != nullptr;
  const double k2t0 = now_ms();
  std::atomic<int> chainLaunched{0};  // set once the fallback sweep's kernel is in its stream (or there is none)
  // ------------------------------------------------------------------ block-parallel path
  struct PathStats {  // what one host thread of the origin pass adds to rmgpu_stats (merged at the end)
    int launches = 0, redo = 0;
    int64_t d2h = 0;
    double cells = 0;
  };
  // One block-parallel launch over the triples trIdx with the scratch set B on stream sq; candidates go to `out`.
  auto run_blocks = [&](const std::vector<int>& trIdx, bool allBlocks, std::vector<int>& redoOut, BlockBufs& B, cudaStream_t sq,
                        std::vector<std::pair<int, Candidate>>& out, PathStats& ps, bool afterChain) {
    std::vector<BlockTask> tasks;
    int64_t snapTot = 0;
    for (int tr : trIdx) {
      const Triple& T = triples[tr];
      const int64_t Q = h_qlen[T.q], L = (int64_t)lib.seqs[T.subject].size();
      const int64_t CB = 32ll << h_hotShift[T.q];
      const int nblk = (int)((Q + CB - 1) / CB);
      int prev = -1;
      // a wrapped triple only has tasks inside the window of its int32 score pass (T.dil is that pass's exact mask)
      const int cFirst = T.wide ? T.lo32 : 0, cLast = T.wide ? T.hi32 : std::min(nblk, kHotBlocks) - 1;
      for (int c = cFirst; c <= cLast; c++) {
        // the last bit stands for every block from there on
        const bool on = allBlocks || (T.wide ? T.dil32 : T.dil).test(std::min(c, kHotBlocks - 1));
        if (!on) {
          prev = -1;
          continue;
        }
        BlockTask t;
        memset(&t, 0, sizeof(t));
        t.q = T.q;
        t.strand = T.strand;
        t.subject = T.subject;
        t.triple = tr;
        t.c0 = (int)(c * CB);
        t.c1 = (c == kHotBlocks - 1) ? (int)Q : (int)std::min<int64_t>(Q, (c + 1) * CB);
        t.prevTask = prev;
        t.ckOff = c > 0 ? ((int64_t)T.q * kHotBlocks + c) * ((int64_t)nGroups * R) + (int64_t)h_subjFirstGroup[T.subject] * R : -1;
        // K1's checkpoint is exact at the left edge of the window (no cell has wrapped before it), K2w wrote the others
        t.ck32Off = (T.wide && c > T.lo32) ? T.ck32Base + (int64_t)(c - T.lo32) * L : -1;
        t.snapOff = snapTot;
        snapTot += 2 * L;
        prev = (int)tasks.size();
        tasks.push_back(t);
        ps.cells += (double)(t.c1 - t.c0) * (double)L;
      }
    }
    if (tasks.empty()) return;
    // longest subjects first: remap prevTask through the permutation
    std::vector<int> perm(tasks.size()), inv(tasks.size());
    std::iota(perm.begin(), perm.end(), 0);
    std::stable_sort(perm.begin(), perm.end(), [&](int x, int y) {
      return lib.seqs[tasks[x].subject].size() > lib.seqs[tasks[y].subject].size();
    });
    for (size_t i = 0; i < perm.size(); i++) inv[perm[i]] = (int)i;
    std::vector<BlockTask> sorted(tasks.size());
    for (size_t i = 0; i < perm.size(); i++) {
      sorted[i] = tasks[perm[i]];
      if (sorted[i].prevTask >= 0) sorted[i].prevTask = inv[sorted[i].prevTask];
    }
    tasks.swap(sorted);
    // tickets go stripe level by stripe level over all tasks, the levels aligned at the LAST stripe: a task with n
    // stripes has its stripe k in level (maxStripes - n) + k.  The long tasks -- the critical path, one dependent
    // ticket per stripe -- start first and run as systolic pipelines while the machine is still filling; the wide
    // levels, where a consumer stripe is claimed (nearly) a whole level after its producer and never polls, come last,
    // so the kernel ends on plenty of parallel work instead of a thin tail of long tasks
    std::vector<int32_t> ticketTask, ticketStripe, prodTicket, consTicket;
    int maxStripes = 0, blockLevelWidth = 0;
    for (auto& t : tasks) {
      t.nStripes = (int)((lib.seqs[t.subject].size() + kR2 * 32 - 1) / (kR2 * 32));
      maxStripes = std::max(maxStripes, t.nStripes);
    }
    {
      std::vector<int32_t> prevOf(tasks.size(), -1);
      size_t levelWidthMax = 0;
      for (int key = 0; key < maxStripes; key++) {
        size_t before = ticketTask.size();
        for (size_t i = 0; i < tasks.size(); i++) {
          const int level = key - (maxStripes - tasks[i].nStripes);
          if (level < 0) continue;
          const int32_t tkt = (int32_t)ticketTask.size();
          ticketTask.push_back((int32_t)i);
          ticketStripe.push_back(level);
          prodTicket.push_back(prevOf[i]);
          consTicket.push_back(-1);
          if (prevOf[i] >= 0) consTicket[prevOf[i]] = tkt;
          prevOf[i] = tkt;
        }
        levelWidthMax = std::max(levelWidthMax, ticketTask.size() - before);
      }
      blockLevelWidth = (int)levelWidthMax;
    }
    const int nTickets = (int)ticketTask.size();
    int blocksPerSM = 0;
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSM, swg_origin_block_kernel, 128, 0));
    if (blocksPerSM < 1) throw Error("K2b does not fit on an SM");
    // the fallback sweep is launched first (high-priority stream) and takes the SM slots it needs; this grid is sized
    // for the whole machine: CTAs that do not fit next to the sweep start when it ends (tickets are claimed
    // dynamically and a ticket's producer is always an older ticket, so CTAs waiting to become resident are harmless)
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)numSMs * blocksPerSM, ((int64_t)nTickets + 3) / 4));
    const int warps = grid * 4;
    // a slot's reader is at most one level (blockLevelWidth tickets) after its writer
    int64_t maxW = 0;
    for (auto& t : tasks) maxW = std::max<int64_t>(maxW, t.c1 - t.c0);
    const int64_t bndStride = ((maxW + 31) >> 5) * 32 + 32;
    // ... but a boundary is only one column block wide (9 KB): up to 2 GB of ring every ticket simply keeps its own slot,
    // and no ticket ever waits for the reader of an older one
    const int64_t slotBytes = bndStride * (int64_t)sizeof(OriginBoundary);
    const int nslot = (int)std::min<int64_t>(
        nTickets, std::max<int64_t>(std::max<int64_t>(2 * (int64_t)warps + 2, 2 * (int64_t)blockLevelWidth + 2), (2ll << 30) / slotBytes));
    size_t tableSize = 1u << 23;  // first-level table (task, relative or pseudo origin): kept very sparse, probing is atomics
    for (;;) {
      const size_t ftableSize = tableSize >> 6;  // final table (triple, origin): the only one downloaded
      BlockArgs a;
      B.d_btasks.upload(tasks, sq);
      B.d_ticketTask.upload(ticketTask, sq);
      B.d_ticketStripe.upload(ticketStripe, sq);
      B.d_prodTicket.upload(prodTicket, sq);
      B.d_consTicket.upload(consTicket, sq);
      a.tasks = B.d_btasks.p;
      a.ticketTask = B.d_ticketTask.p;
      a.ticketStripe = B.d_ticketStripe.p;
      a.prodTicket = B.d_prodTicket.p;
      a.consTicket = B.d_consTicket.p;
      a.nTickets = nTickets;
      a.codes_fwd = d_codes_fwd.p;
      a.codes_rc = d_codes_rc.p;
      a.qoff = d_qoff.p;
      a.scodes = d_scodes.p;
      a.soff = d_soff.p;
      a.slen = d_slen.p;
      memcpy(a.mat, h_mat, sizeof(h_mat));
      a.gap_init = p.gap_init;
      a.ins_ext = p.ins_gap_ext;
      a.del_ext = p.del_gap_ext;
      a.minscore = p.minscore;
      a.ck = d_ck.p;
      a.ck32 = (&B == &bb[1]) ? d_ck32.p : nullptr;  // only the wrapped triples' launch has int32 checkpoints
      a.bnd = B.d_bnd2.ensure((size_t)nslot * bndStride);
      a.bndStride = bndStride;
      a.nslot = nslot;
      a.progress = B.d_progress.ensure((size_t)nTickets);
      a.finished = B.d_finished.ensure((size_t)nTickets);
      a.abort_flag = B.d_abort.ensure(1);
      a.snap = B.d_snap.ensure((size_t)std::max<int64_t>(snapTot, 1));
      a.ticket = B.d_ticket.ensure(1);
      a.keys = B.d_keys.ensure(tableSize);
      a.vals = B.d_vals.ensure(tableSize);
      a.tableMask = (uint32_t)tableSize - 1;
      a.overflow = B.d_overflow.ensure(1);
      a.pbest = B.d_pbest.ensure((size_t)std::max<int64_t>(snapTot, 1));
      CUDA_CHECK(cudaMemsetAsync(a.pbest, 0, (size_t)std::max<int64_t>(snapTot, 1) * 8, sq));
      a.dbg = nullptr;
      if (dbgK2) {
        a.dbg = B.d_dbg.ensure(72);
        CUDA_CHECK(cudaMemsetAsync(a.dbg, 0, 72 * 8, sq));
      }
      CUDA_CHECK(cudaMemsetAsync(a.abort_flag, 0, sizeof(int), sq));
      CUDA_CHECK(cudaMemsetAsync(a.progress, 0, (size_t)nTickets * sizeof(int), sq));
      CUDA_CHECK(cudaMemsetAsync(a.finished, 0, (size_t)nTickets * sizeof(int), sq));
      CUDA_CHECK(cudaMemsetAsync(a.keys, 0xFF, tableSize * 8, sq));
      CUDA_CHECK(cudaMemsetAsync(a.vals, 0, tableSize * 8, sq));
      CUDA_CHECK(cudaMemsetAsync(a.ticket, 0, sizeof(int), sq));
      CUDA_CHECK(cudaMemsetAsync(a.overflow, 0, sizeof(int), sq));
      ResolveArgs r;
      r.tasks = B.d_btasks.p;
      r.nTasks = (int)tasks.size();
      r.pbest = a.pbest;
      r.slen = d_slen.p;
      r.snap = a.snap;
      r.keys = a.keys;
      r.vals = a.vals;
      r.tableSize = (uint32_t)tableSize;
      r.fkeys = B.d_fkeys.ensure(ftableSize);
      r.fvals = B.d_fvals.ensure(ftableSize);
      r.fmask = (uint32_t)ftableSize - 1;
      r.overflow = a.overflow;
      r.redo = B.d_redo.ensure(triples.size());
      r.verbose = allBlocks ? 1 : 0;
      CUDA_CHECK(cudaMemsetAsync(r.fkeys, 0xFF, ftableSize * 8, sq));
      CUDA_CHECK(cudaMemsetAsync(r.fvals, 0, ftableSize * 8, sq));
      CUDA_CHECK(cudaMemsetAsync(r.redo, 0, triples.size() * sizeof(int), sq));
      // let the (small, high-priority) fallback sweep take its SM slots first, then fill the rest of the machine
      for (double w0 = now_ms(); afterChain && !chainLaunched.load() && now_ms() - w0 < 100.0;) std::this_thread::yield();
      const double tLaunch = now_ms() - k2t0;
      EventTimer tm;
      tm.start(sq);
      swg_origin_block_kernel<<<grid, 128, 0, sq>>>(a);
      resolve_candidates_kernel<<<numSMs * 2, 256, 0, sq>>>(r);
      tm.stop(sq);
      CUDA_CHECK(cudaGetLastError());
      int overflow = 0;
      std::vector<int> redoFlag(triples.size(), 0);
      std::vector<unsigned long long> keys(ftableSize), vals(ftableSize);
      CUDA_CHECK(cudaStreamSynchronize(sq));  // as in run_chain: never block inside a pageable copy while the other path launches
      const double tSync = now_ms() - k2t0;
      CUDA_CHECK(cudaMemcpyAsync(&overflow, a.overflow, sizeof(int), cudaMemcpyDeviceToHost, sq));
      CUDA_CHECK(cudaMemcpyAsync(redoFlag.data(), r.redo, triples.size() * sizeof(int), cudaMemcpyDeviceToHost, sq));
      CUDA_CHECK(cudaMemcpyAsync(keys.data(), r.fkeys, ftableSize * 8, cudaMemcpyDeviceToHost, sq));
      CUDA_CHECK(cudaMemcpyAsync(vals.data(), r.fvals, ftableSize * 8, cudaMemcpyDeviceToHost, sq));
      CUDA_CHECK(cudaStreamSynchronize(sq));
      check_abort_on("K2b", B.d_abort.p, sq);
      if (dbgK2) {
        fprintf(stderr, "[K2 dbg] blocks: %zu triples %zu tasks %d tickets grid %d  kernels %.2f ms (launched at +%.2f, done at +%.2f, results on the host at +%.2f ms)\n",
                trIdx.size(), tasks.size(), nTickets, grid, tm.ms(), tLaunch, tSync, now_ms() - k2t0);
        unsigned long long d[72];
        CUDA_CHECK(cudaMemcpy(d, a.dbg, 72 * 8, cudaMemcpyDeviceToHost));
        fprintf(stderr, "[K2 dbg] blocks: last ticket of stripe level k ended at (ms):");
        for (int k = 0; k < std::min(maxStripes, 64); k++) fprintf(stderr, " %.1f", (double)d[8 + k] * 1e-6);
        fprintf(stderr, "\n");
        fprintf(stderr, "[K2 dbg] blocks: mean ticket %.0f cyc, producer-wait %.1f%% of ticket cycles, slot-wait %.2e cyc, warps busy until exit: mean %.1f%% of the last warp's (%.2f ms), levels %d, widest level %d\n",
                (double)d[3] / (double)std::max(d[2], 1ull), 100.0 * (double)d[0] / (double)std::max(d[3], 1ull), (double)d[1],
                100.0 * (double)d[4] / ((double)warps * (double)std::max(d[5], 1ull)), (double)d[5] * 1e-6, maxStripes, blockLevelWidth);
        {
          const int tk = (int)(d[6] & 0xFFFFFF);
          const BlockTask& bt = tasks[ticketTask[tk]];
          int tasksOfTriple = 0;
          for (auto& t2 : tasks) tasksOfTriple += t2.triple == bt.triple;
          fprintf(stderr, "[K2 dbg] blocks: slowest ticket %d of %d: %.0f kcyc, %.0f kcyc waiting for its producer; stripe %d of %d, columns [%d, %d), subject %d (%zu rows), "
                  "triple with %d tasks, K1 best %d\n", tk, nTickets, (double)(d[6] >> 44) * 1.024, (double)((d[6] >> 24) & 0xFFFFF) * 1.024, ticketStripe[tk], bt.nStripes,
                  bt.c0, bt.c1, bt.subject, lib.seqs[bt.subject].size(), tasksOfTriple,
                  pairbest[((size_t)bt.q * nS + bt.subject) * 2 + bt.strand]);
        }
      }
      ps.launches += 2;
      ps.d2h += (int64_t)ftableSize * 16;
      if (overflow) {  // a probe sequence got long in one of the two tables
        tableSize *= 4;
        if (tableSize > (1u << 28)) throw Error("candidate table overflow");
        continue;
      }
      for (size_t tr = 0; tr < triples.size(); tr++)
        if (redoFlag[tr]) {
          if (allBlocks) throw Error("origin pass: a pseudo origin could not be resolved with every block swept");
          redoOut.push_back((int)tr);
        }
      for (size_t i = 0; i < keys.size(); i++) {
        if (keys[i] == kEmptyKey) continue;
        const int tr = (int)(keys[i] >> 40);
        if (redoFlag[tr]) continue;
        Candidate c;
        c.q = triples[tr].q;
        c.strand = triples[tr].strand;
        c.subject = triples[tr].subject;
        c.i0 = (int)((keys[i] >> 20) & 0xFFFFFu);
        c.j0 = (int)(keys[i] & 0xFFFFFu);
        decode_val(vals[i], c);
        out.push_back({tr, c});
      }
      break;
    }
  };

  // ------------------------------------------------------------------ stripe-pipeline path (full sweeps)
  std::vector<std::pair<int, Candidate>> foundChain;
  double chainMs = 0, chainCells = 0;
  int chainLaunches = 0, chainRedo = 0;
  int64_t chainD2H = 0;
  auto run_chain = [&](const std::vector<int>& trIdx) {
    if (trIdx.empty()) return;
    // A sweep is one long dependent chain of columns: when there are too few of them to fill the machine the time is
    // the per-column latency of one warp, which falls with the rows a lane owns -> 2 rows per lane instead of 8.
    int64_t stripes8 = 0;
    for (int tr : trIdx) stripes8 += (int64_t)(lib.seqs[triples[tr].subject].size() + 255) / 256;
    const int chainR = (stripes8 * 4 <= (int64_t)numSMs * 16) ? 2 : 8;
    auto origin_range_check = [&](const OriginTask& t) {
      if (((int64_t)(t.c1 - t.c0) + 2) * ((int64_t)lib.seqs[t.subject].size() + 1) >= (1ll << 32) - 1)
        throw Error("query x subject too large for one origin sweep (" + queries.names[t.q] + " x " + lib.names[t.subject] +
                    "): fragment the query as SimpleBatcher does");
    };
    std::vector<OriginTask> tasks;
    std::vector<char> taskFull;  // swept with origins over the whole query
    std::vector<int32_t> ticketTask;
    for (int tr : trIdx) {
      const Triple& T = triples[tr];
      OriginTask t;
      memset(&t, 0, sizeof(t));
      t.q = T.q;
      t.strand = T.strand;
      t.subject = T.subject;
      t.nStripes = (int)((lib.seqs[T.subject].size() + chainR * 32 - 1) / (chainR * 32));
      // without checkpoints the cold column blocks are swept score-only; after an int16 wrap the mask is unreliable
      const size_t idx = ((size_t)T.q * nS + T.subject) * 2 + T.strand;
      // after an int16 wrap K1 under-reports the blocks that follow (true scores decay by only ~5-6 per column along
      // a gap), so the mask is stretched 8 blocks forward; a cell >= minscore met in a score-only block still forces
      // a full redo of the task, so exactness never depends on this guess
      HotMask fwd = T.dil;
      const int64_t CBq = 32ll << h_hotShift[T.q];
      if (pairbest[idx] >= 32767) {
        const int stretch = (int)std::max<int64_t>(8, (8192 + CBq - 1) / CBq);  // at least 8 blocks and 8192 columns
        HotMask sh = T.dil;
        for (int k = 1; k <= stretch; k++) {
          sh = sh.shl1();
          fwd = fwd | sh;
        }
      }
      const bool fullSweep = fullMode || T.dil.none();
      const HotMask modeMask = fullSweep ? HotMask::all() : fwd;
      for (int k = 0; k < kHotWords; k++) t.mode[k] = modeMask.w[k];
      taskFull.push_back(fullSweep ? 1 : 0);
      t.hotShift = h_hotShift[T.q];
      t.c0 = 0;
      t.c1 = h_qlen[T.q];
      t.ckOff = -1;
      if (pairbest[idx] >= 32767 && !fullSweep && ckEnabled && classesExact) {
        // int16 wrap: K1 is exact up to the first hot block (a score passes minscore long before it wraps), so the
        // sweep starts at the checkpoint before the first block of the mask; to the right of the last one K1 saw
        // nothing >= minscore, and since a wrap only ever LOWERS K1's values (the wrapped sum is negative and loses
        // every max) a path it missed there descends from the wrapped alignment and scores below its maximum --
        // unless it leaves the window at >= minscore, which the kernel reports (then the task is redone in full)
        const int64_t CB = CBq;
        const int lo = fwd.lowest(), hi = fwd.highest();
        t.c0 = (int)std::min<int64_t>(lo * CB, h_qlen[T.q]);
        if (lo > 0) t.ckOff = ((int64_t)T.q * kHotBlocks + lo) * ((int64_t)nGroups * R) + (int64_t)h_subjFirstGroup[T.subject] * R;
        if (hi < kHotBlocks - 1) t.c1 = (int)std::min<int64_t>(h_qlen[T.q], (hi + 1) * CB);
        if (t.c0 >= t.c1) t.c0 = 0, t.c1 = h_qlen[T.q], t.ckOff = -1;
      }
      // a sweep carries origins as (column in window + 1) * (L + 1) + row in 32 bits (all ones = "from the left")
      origin_range_check(t);
      t.firstTicket = (int)ticketTask.size();
      for (int k = 0; k < t.nStripes; k++) ticketTask.push_back((int32_t)tasks.size());
      tasks.push_back(t);
      chainCells += (double)(t.c1 - t.c0) * (double)lib.seqs[T.subject].size();
    }
    int blocksPerSM = 0;
    if (chainR == 2)
      CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSM, swg_origin_kernel<2>, 128, 0));
    else
      CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSM, swg_origin_kernel<8>, 128, 0));
    if (blocksPerSM < 1) throw Error("K2 does not fit on an SM");
    blocksPerSM = std::min(blocksPerSM, 4);
    const int64_t bndStride = (((int64_t)maxQ + 31) >> 5) * 32 + 32;
    const int nTickets = (int)ticketTask.size();
    const int grid = (int)std::min<int64_t>((int64_t)numSMs * blocksPerSM, ((int64_t)nTickets + 3) / 4);
    const int warps = grid * 4;
    const int nslot = std::min(nTickets, 2 * warps + 2);
    for (int round = 0; round < 2; round++) {
      std::vector<int> leftFlag(tasks.size(), 0);
      size_t tableSize = 1u << 18;  // grows x4 on overflow
      for (;;) {
        OriginArgs a;
        d_tasksC.upload(tasks, stream2);
        d_ticketTaskC.upload(ticketTask, stream2);
        a.tasks = d_tasksC.p;
        a.ticketTask = d_ticketTaskC.p;
        a.nTickets = nTickets;
        a.codes_fwd = d_codes_fwd.p;
        a.codes_rc = d_codes_rc.p;
        a.qoff = d_qoff.p;
        a.qlen = d_qlen.p;
        a.scodes = d_scodes.p;
        a.soff = d_soff.p;
        a.slen = d_slen.p;
        memcpy(a.mat, h_mat, sizeof(h_mat));
        a.gap_init = p.gap_init;
        a.ins_ext = p.ins_gap_ext;
        a.del_ext = p.del_gap_ext;
        a.minscore = p.minscore;
        a.bnd = d_bndC.ensure((size_t)nslot * bndStride);
        a.bndStride = bndStride;
        a.nslot = nslot;
        a.progress = d_progressC.ensure((size_t)nTickets);
        a.finished = d_finishedC.ensure((size_t)nTickets);
        a.ticket = d_ticketC.ensure(1);
        a.keys = d_keysC.ensure(tableSize);
        a.vals = d_valsC.ensure(tableSize);
        a.tableMask = (uint32_t)tableSize - 1;
        a.overflow = d_overflowC.ensure(1);
        a.abort_flag = d_abortC.ensure(1);
        a.dbg = nullptr;
        if (dbgK2) {
          a.dbg = d_dbgC.ensure(8);
          CUDA_CHECK(cudaMemsetAsync(a.dbg, 0, 64, stream2));
        }
        a.ck = d_ck.p;
        a.leftFlag = d_leftFlagC.ensure(tasks.size());
        CUDA_CHECK(cudaMemsetAsync(a.abort_flag, 0, sizeof(int), stream2));
        CUDA_CHECK(cudaMemsetAsync(a.leftFlag, 0, tasks.size() * sizeof(int), stream2));
        CUDA_CHECK(cudaMemsetAsync(a.keys, 0xFF, tableSize * 8, stream2));
        CUDA_CHECK(cudaMemsetAsync(a.vals, 0, tableSize * 8, stream2));
        CUDA_CHECK(cudaMemsetAsync(a.progress, 0, (size_t)nTickets * sizeof(int), stream2));
        CUDA_CHECK(cudaMemsetAsync(a.finished, 0, (size_t)nTickets * sizeof(int), stream2));
        CUDA_CHECK(cudaMemsetAsync(a.ticket, 0, sizeof(int), stream2));
        CUDA_CHECK(cudaMemsetAsync(a.overflow, 0, sizeof(int), stream2));
        EventTimer tm;
        tm.start(stream2);
        if (chainR == 2)
          swg_origin_kernel<2><<<grid, 128, 0, stream2>>>(a);
        else
          swg_origin_kernel<8><<<grid, 128, 0, stream2>>>(a);
        tm.stop(stream2);
        chainLaunched.store(1);
        CUDA_CHECK(cudaGetLastError());
        std::vector<unsigned long long> keys(tableSize), vals(tableSize);
        int overflow = 0;
        // wait for the sweep BEFORE issuing the downloads: a copy into pageable memory blocks inside the driver until
        // the stream reaches it, and while it does the other thread's launches are held back (the two paths serialise)
        CUDA_CHECK(cudaStreamSynchronize(stream2));
        CUDA_CHECK(cudaMemcpyAsync(&overflow, a.overflow, sizeof(int), cudaMemcpyDeviceToHost, stream2));
        CUDA_CHECK(cudaMemcpyAsync(leftFlag.data(), a.leftFlag, tasks.size() * sizeof(int), cudaMemcpyDeviceToHost, stream2));
        CUDA_CHECK(cudaMemcpyAsync(keys.data(), a.keys, tableSize * 8, cudaMemcpyDeviceToHost, stream2));
        CUDA_CHECK(cudaMemcpyAsync(vals.data(), a.vals, tableSize * 8, cudaMemcpyDeviceToHost, stream2));
        CUDA_CHECK(cudaStreamSynchronize(stream2));
        check_abort_on("K2", a.abort_flag, stream2);
        chainMs += tm.ms();
        if (dbgK2) {
          unsigned long long d[8];
          CUDA_CHECK(cudaMemcpy(d, a.dbg, 64, cudaMemcpyDeviceToHost));
          int64_t cols = 0, maxCols = 0, maxStr = 0;
          for (auto& t : tasks) cols += t.c1 - t.c0, maxCols = std::max<int64_t>(maxCols, t.c1 - t.c0), maxStr = std::max<int64_t>(maxStr, t.nStripes);
          fprintf(stderr, "[K2 dbg] chain: %zu triples %d tickets R=%d grid %d round %d  kernel %.2f ms (done at +%.2f ms); window columns mean %.0f max %lld, "
                  "most stripes %lld; mean ticket %.0f cyc, producer-wait %.1f%%, slot-wait %.1f%% of ticket cycles\n", trIdx.size(),
                  nTickets, chainR, grid, round, tm.ms(), now_ms() - k2t0, (double)cols / std::max<size_t>(tasks.size(), 1), (long long)maxCols, (long long)maxStr,
                  (double)d[3] / std::max<double>((double)d[2], 1), 100.0 * d[1] / std::max<double>((double)d[3], 1), 100.0 * d[0] / std::max<double>((double)d[3], 1));
        }
        chainLaunches += 1;
        chainD2H += (int64_t)tableSize * 16;
        if (overflow) {
          tableSize *= 4;
          if (tableSize > (1u << 28)) throw Error("candidate table overflow");
          continue;
        }
        for (size_t i = 0; i < keys.size(); i++) {
          if (keys[i] == kEmptyKey) continue;
          const int ti = (int)(keys[i] >> 32);
          if (leftFlag[ti]) continue;
          const uint32_t origin = (uint32_t)keys[i];
          const OriginTask& t = tasks[ti];
          Candidate c;
          c.q = t.q;
          c.strand = t.strand;
          c.subject = t.subject;
          const uint32_t Lp1 = (uint32_t)lib.seqs[t.subject].size() + 1u;
          c.i0 = (int)(origin / Lp1) + t.c0;  // origins are carried relative to the task's window
          c.j0 = (int)(origin % Lp1);
          decode_val(vals[i], c);
          foundChain.push_back({trIdx[ti], c});
        }
        break;
      }
      // tasks that met a hit needing origins from a score-only region: once more with origins everywhere
      bool any = false;
      for (size_t i = 0; i < tasks.size(); i++) {
        if (leftFlag[i]) {
          if (taskFull[i]) throw Error("origin pass: a fully swept task still reports a left origin");
          taskFull[i] = 1;
          for (int k = 0; k < kHotWords; k++) tasks[i].mode[k] = ~0ull;
          tasks[i].c0 = 0;
          tasks[i].c1 = h_qlen[tasks[i].q];
          tasks[i].ckOff = -1;
          origin_range_check(tasks[i]);
          any = true;
          chainRedo++;
        } else {
          for (int k = 0; k < kHotWords; k++) tasks[i].mode[k] = 0;  // done: nothing to sweep with origins, and nothing >= minscore can appear
          tasks[i].c0 = 0;
          tasks[i].c1 = 0;    // (an empty window makes the ticket a no-op)
          tasks[i].ckOff = -1;
        }
      }
      if (!any) break;
    }
  };

  // ------------------------------------------------------------------ wrapped triples: int32 score pass, then K2b
  std::vector<std::pair<int, Candidate>> foundWide;
  PathStats psWide, psMain;
  auto run_wide = [&](const std::vector<int>& trIdx, std::vector<int>& fallback) {
    if (trIdx.empty()) return;
    std::vector<WideTask> tasks;
    std::vector<int32_t> ticketTask;
    std::vector<int> taskTr;
    int64_t ckTot = 0, maxW = 0;
    for (int tr : trIdx) {
      Triple& T = triples[tr];
      const int64_t Q = h_qlen[T.q], L = (int64_t)lib.seqs[T.subject].size();
      const int64_t CB = 32ll << h_hotShift[T.q];
      const int nblk = (int)((Q + CB - 1) / CB);
      // the window of the one-sweep fallback: from the first block of K1's (there still exact) mask to its last one,
      // stretched because a wrap only ever LOWERS K1's values: to the right of the wrapped alignment K1 under-reports
      // the decaying fan (true scores fall by ~5-6 per column along a gap).  Whatever is >= minscore when it leaves
      // the window raises the edge flag, so exactness never depends on the stretch
      HotMask fwd = T.dil;
      const int stretch = (int)std::max<int64_t>(8, (8192 + CB - 1) / CB);
      HotMask sh = T.dil;
      for (int k = 1; k <= stretch; k++) {
        sh = sh.shl1();
        fwd = fwd | sh;
      }
      const int lo = fwd.lowest(), hi = std::min(fwd.highest(), nblk - 1);
      const int64_t need = (int64_t)(hi - lo + 1) * L;
      if (lo < 0 || hi < lo || ckTot + need > (1ll << 29)) {  // (4 GB of int32 checkpoints per pass at most)
        fallback.push_back(tr);
        continue;
      }
      WideTask t;
      memset(&t, 0, sizeof(t));
      t.q = T.q;
      t.strand = T.strand;
      t.subject = T.subject;
      t.hotShift = h_hotShift[T.q];
      t.c0 = (int)(lo * CB);
      t.c1 = (hi >= nblk - 1) ? (int)Q : (int)((hi + 1) * CB);
      t.ckOff = lo > 0 ? ((int64_t)T.q * kHotBlocks + lo) * ((int64_t)nGroups * R) + (int64_t)h_subjFirstGroup[T.subject] * R : -1;
      t.ck32Off = ckTot;
      ckTot += need;
      T.lo32 = lo;
      T.hi32 = hi;
      T.ck32Base = t.ck32Off;
      t.nStripes = (int)((L + kRW * 32 - 1) / (kRW * 32));
      t.firstTicket = (int)ticketTask.size();
      for (int k = 0; k < t.nStripes; k++) ticketTask.push_back((int32_t)tasks.size());
      maxW = std::max<int64_t>(maxW, t.c1 - t.c0);
      psWide.cells += (double)(t.c1 - t.c0) * (double)L;
      taskTr.push_back(tr);
      tasks.push_back(t);
    }
    std::vector<int> okTr;
    if (!tasks.empty()) {
      int blocksPerSM = 0;
      CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSM, swg_score32_kernel, 128, 0));
      if (blocksPerSM < 1) throw Error("K2w does not fit on an SM");
      blocksPerSM = std::min(blocksPerSM, 4);
      const int nTickets = (int)ticketTask.size();
      const int grid = (int)std::min<int64_t>((int64_t)numSMs * blocksPerSM, ((int64_t)nTickets + 3) / 4);
      const int warps = grid * 4;
      const int nslot = std::min(nTickets, 2 * warps + 2);
      const int64_t bndStride = ((maxW + 31) >> 5) * 32 + 32;
      WideArgs a;
      d_wtasks.upload(tasks, stream2);
      d_ticketTaskW.upload(ticketTask, stream2);
      a.tasks = d_wtasks.p;
      a.ticketTask = d_ticketTaskW.p;
      a.nTickets = nTickets;
      a.codes_fwd = d_codes_fwd.p;
      a.codes_rc = d_codes_rc.p;
      a.qoff = d_qoff.p;
      a.qlen = d_qlen.p;
      a.scodes = d_scodes.p;
      a.soff = d_soff.p;
      a.slen = d_slen.p;
      memcpy(a.mat, h_mat, sizeof(h_mat));
      a.gap_init = p.gap_init;
      a.ins_ext = p.ins_gap_ext;
      a.del_ext = p.del_gap_ext;
      a.minscore = p.minscore;
      a.bnd = d_bndW.ensure((size_t)nslot * bndStride);
      a.bndStride = bndStride;
      a.nslot = nslot;
      a.progress = d_progressW.ensure((size_t)nTickets);
      a.finished = d_finishedW.ensure((size_t)nTickets);
      a.ticket = d_ticketW.ensure(1);
      a.abort_flag = d_abortW.ensure(1);
      a.ck = d_ck.p;
      a.ck32 = d_ck32.ensure((size_t)std::max<int64_t>(ckTot, 1));
      a.hot32 = d_hot32.ensure(tasks.size() * kHotWords);
      a.edgeFlag = d_edgeW.ensure(tasks.size());
      CUDA_CHECK(cudaMemsetAsync(a.progress, 0, (size_t)nTickets * sizeof(int), stream2));
      CUDA_CHECK(cudaMemsetAsync(a.finished, 0, (size_t)nTickets * sizeof(int), stream2));
      CUDA_CHECK(cudaMemsetAsync(a.ticket, 0, sizeof(int), stream2));
      CUDA_CHECK(cudaMemsetAsync(a.abort_flag, 0, sizeof(int), stream2));
      CUDA_CHECK(cudaMemsetAsync(a.hot32, 0, tasks.size() * kHotWords * 8, stream2));
      CUDA_CHECK(cudaMemsetAsync(a.edgeFlag, 0, tasks.size() * sizeof(int), stream2));
      EventTimer tm;
      tm.start(stream2);
      swg_score32_kernel<<<grid, 128, 0, stream2>>>(a);
      tm.stop(stream2);
      chainLaunched.store(1);
      CUDA_CHECK(cudaGetLastError());
      std::vector<unsigned long long> hot(tasks.size() * kHotWords);
      std::vector<int> edge(tasks.size());
      CUDA_CHECK(cudaStreamSynchronize(stream2));  // (never block inside a pageable copy while the other thread launches)
      CUDA_CHECK(cudaMemcpyAsync(hot.data(), a.hot32, hot.size() * 8, cudaMemcpyDeviceToHost, stream2));
      CUDA_CHECK(cudaMemcpyAsync(edge.data(), a.edgeFlag, edge.size() * sizeof(int), cudaMemcpyDeviceToHost, stream2));
      CUDA_CHECK(cudaStreamSynchronize(stream2));
      check_abort_on("K2w", a.abort_flag, stream2);
      psWide.launches += 1;
      psWide.d2h += (int64_t)(hot.size() * 8 + edge.size() * 4);
      int nEdge = 0;
      for (size_t i = 0; i < tasks.size(); i++) {
        Triple& T = triples[taskTr[i]];
        if (edge[i]) {  // something >= minscore leaves the window: the one-sweep path handles it (with its own redo)
          fallback.push_back(taskTr[i]);
          nEdge++;
          continue;
        }
        HotMask m;
        for (int k = 0; k < kHotWords; k++) m.w[k] = hot[i * kHotWords + k];
        T.dil32 = (m | m.shr1()) & HotMask::range(T.lo32, T.hi32);
        okTr.push_back(taskTr[i]);
      }
      if (dbgK2)
        fprintf(stderr, "[K2 dbg] wide: %zu triples %d tickets R=%d grid %d  kernel %.2f ms (done at +%.2f ms); widest window %lld columns, "
                "%.1f MB of int32 checkpoints, %d left the window\n", tasks.size(), nTickets, kRW, grid, tm.ms(), now_ms() - k2t0, (long long)maxW,
                (double)ckTot * 8e-6, nEdge);
    }
    std::vector<int> redoW;
    run_blocks(okTr, false, redoW, bb[1], stream2, foundWide, psWide, false);
    // a pseudo origin that leads out of the window (or into a block of it that was not swept): the one-sweep path
    for (int tr : redoW) fallback.push_back(tr);
    psWide.redo += (int)fallback.size();
    for (int tr : fallback) triples[tr].wide = false;
  };

  std::vector<int> blockTr, chainTr, wideTr, redo;
  for (size_t i = 0; i < triples.size(); i++) (triples[i].wide ? wideTr : triples[i].chain ? chainTr : blockTr).push_back((int)i);
  // the latency-bound paths (one dependent chain of columns per triple: a few warps busy for milliseconds) run on a
  // second stream from a second host thread while the block-parallel path fills the machine
  std::exception_ptr chainErr;
  std::thread chainThread([&] {
    try {
      if (cudaSetDevice(device) != cudaSuccess) throw Error("cudaSetDevice failed in the K2 fallback thread");
      run_chain(chainTr);
      std::vector<int> fallback;
      run_wide(wideTr, fallback);
      chainLaunched.store(1);
      run_chain(fallback);
    } catch (...) {
      chainErr = std::current_exception();
    }
    chainLaunched.store(1);
  });
  std::exception_ptr blockErr;
  static const bool serialK2 = getenv("RMGPU_K2_SERIAL") != nullptr;  // measurement: the two paths one after the other
  if (serialK2 && chainThread.joinable()) chainThread.join();
  try {
    run_blocks(blockTr, false, redo, bb[0], stream, found, psMain, !chainTr.empty());
    if (!redo.empty()) {
      psMain.redo += (int)redo.size();
      std::vector<int> none;
      run_blocks(redo, true, none, bb[0], stream, found, psMain, false);
    }
  } catch (...) {
    blockErr = std::current_exception();
  }
  if (chainThread.joinable()) chainThread.join();
  if (blockErr) std::rethrow_exception(blockErr);
  if (chainErr) std::rethrow_exception(chainErr);
  found.insert(found.end(), foundWide.begin(), foundWide.end());
  for (const PathStats* ps : {&psMain, &psWide}) {
    st.launches += ps->launches;
    st.d2h_bytes += ps->d2h;
    st.n_origin_redo += ps->redo;
    st.origin_cells += ps->cells;
  }
  found.insert(found.end(), foundChain.begin(), foundChain.end());
  st.launches += chainLaunches;
  st.d2h_bytes += chainD2H;
  st.n_origin_redo += chainRedo;
  st.origin_cells += chainCells;
  st.ms_origin += now_ms() - k2t0;  // both paths overlap: wall time of the origin pass (kernels, table downloads, host merge)
  // one candidate per (triple, origin)
  std::sort(found.begin(), found.end(), [](const std::pair<int, Candidate>& x, const std::pair<int, Candidate>& y) {
    if (x.first != y.first) return x.first < y.first;
    if (x.second.i0 != y.second.i0) return x.second.i0 < y.second.i0;
    if (x.second.j0 != y.second.j0) return x.second.j0 < y.second.j0;
    if (x.second.raw != y.second.raw) return x.second.raw > y.second.raw;
    if (x.second.i1 != y.second.i1) return x.second.i1 < y.second.i1;
    return x.second.j1 < y.second.j1;
  });
  for (size_t i = 0; i < found.size(); i++) {
    if (i > 0 && found[i].first == found[i - 1].first && found[i].second.i0 == found[i - 1].second.i0 &&
        found[i].second.j0 == found[i - 1].second.j0)
      continue;
    cands.push_back(found[i].second);
  }
  std::sort(cands.begin(), cands.end(), [](const Candidate& x, const Candidate& y) {
    if (x.q != y.q) return x.q < y.q;
    if (x.strand != y.strand) return x.strand < y.strand;
    if (x.subject != y.subject) return x.subject < y.subject;
    if (x.i0 != y.i0) return x.i0 < y.i0;
    return x.j0 < y.j0;
  });
  st.n_candidates += (int)cands.size();
}

// K3 with the band schedule W = max(bandwidth,1) * 2^k until the banded score reproduces the raw score.  The edit
// operations stay on the device (d_ops, one slot of n + m bytes per candidate): K4 spells the rows from them.
void rmgpu_session::run_traceback(const rmgpu_params& p, const std::vector<Candidate>& cands, rmgpu_stats& st) {
  const size_t nC = cands.size();
  h_candAll.resize(nC);
  int64_t opsTotal = 0;
  std::vector<int> todo(nC), W(nC);
  std::iota(todo.begin(), todo.end(), 0);
  for (size_t i = 0; i < nC; i++) {
    const Candidate& c = cands[i];
    const int n = c.i1 - c.i0 + 1, m = c.j1 - c.j0 + 1;
    W[i] = std::max(p.bandwidth, 1) + std::max(n, m) / 32;
    TraceCand& t = h_candAll[i];
    memset(&t, 0, sizeof(t));
    t.q = c.q;
    t.strand = c.strand;
    t.subject = c.subject;
    t.i0 = c.i0;
    t.j0 = c.j0;
    t.i1 = c.i1;
    t.j1 = c.j1;
    t.raw = c.raw;
    t.id = (int32_t)i;
    t.opsOff = opsTotal;
    opsTotal += n + m;
  }
  opsBytes = opsTotal;
  if (nC == 0) return;
  d_ops.ensure((size_t)opsTotal);
  d_opsLen.ensure(nC);
  d_status.ensure(nC);
  d_candAll.upload(h_candAll, stream);
  auto band_of = [&](int id) {
    const Candidate& c = cands[id];
    return (int64_t)std::abs((c.i1 - c.i0) - (c.j1 - c.j0)) + 2 * (int64_t)W[id] + 1;
  };
  const int64_t kDirBudget = 4ll << 30;  // direction bytes per launch
  std::vector<int32_t> hlen(nC);
  while (!todo.empty()) {
    std::vector<int> next;
    // narrow bands first (one warp each), then wide bands (a team of up to 512 threads of a block each)
    // ... and the longest alignments of each kind first: a candidate is one dependent chain of 2n waves, so the longest one
    // bounds the launch and must not be the last block the scheduler starts
    std::stable_sort(todo.begin(), todo.end(), [&](int x, int y) {
      const bool wx = band_of(x) > 64, wy = band_of(y) > 64;
      if (wx != wy) return wx < wy;
      return cands[x].i1 - cands[x].i0 > cands[y].i1 - cands[y].i0;
    });
    size_t at = 0;
    while (at < todo.size()) {
      std::vector<TraceCand> tc;
      int64_t dirTot = 0, rowTot = 0;
      while (at < todo.size()) {
        const int id = todo[at];
        const Candidate& c = cands[id];
        int n = c.i1 - c.i0 + 1, m = c.j1 - c.j0 + 1, w = W[id];
        int64_t bw = (int64_t)std::abs(n - m) + 2 * (int64_t)w + 1;
        // the kernel itself retries once with the doubled band (same schedule, no second launch for the few that
        // need it): the scratch is sized for that last attempt
        const int64_t bwLast = (int64_t)std::abs(n - m) + 2 * ((int64_t)w << (kTraceAttempts - 1)) + 1;
        int64_t dirBytes = (int64_t)(n + 1) * bwLast;
        if (!tc.empty() && (dirTot + dirBytes > kDirBudget || (band_of(tc[0].id) > 64) != (bw > 64))) break;
        TraceCand t = h_candAll[id];
        t.W = w;
        t.attempts = kTraceAttempts;
        t.dirOff = dirTot;
        t.rowOff = rowTot;
        dirTot += dirBytes;
        rowTot += 3 * bwLast;
        tc.push_back(t);
        at++;
      }
      TraceArgs a;
      d_cand.upload(tc, stream);
      a.cand = d_cand.p;
      a.nCand = (int)tc.size();
      a.codes_fwd = d_codes_fwd.p;
      a.codes_rc = d_codes_rc.p;
      a.qoff = d_qoff.p;
      a.scodes = d_scodes.p;
      a.soff = d_soff.p;
      memcpy(a.mat, h_mat, sizeof(h_mat));
      a.gap_init = p.gap_init;
      a.ins_ext = p.ins_gap_ext;
      a.del_ext = p.del_gap_ext;
      a.dir = d_dir.ensure((size_t)dirTot);
      a.rows = d_rows.ensure((size_t)rowTot);
      a.ops = d_ops.p;
      a.opsLen = d_opsLen.p;
      a.status = d_status.p;
      const bool wide = band_of(tc[0].id) > 64;
      EventTimer tm;
      tm.start(stream);
      if (wide)
        traceback_kernel<true><<<(int)tc.size(), kTraceBlockThreads, 0, stream>>>(a);
      else
        traceback_kernel<false><<<((int)tc.size() + kTraceWarps - 1) / kTraceWarps, kTraceWarps * 32, 0, stream>>>(a);
      tm.stop(stream);
      CUDA_CHECK(cudaGetLastError());
      CUDA_CHECK(cudaMemcpyAsync(hlen.data(), d_opsLen.p, nC * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
      CUDA_CHECK(cudaStreamSynchronize(stream));
      st.ms_traceback += tm.ms();
      st.launches += 1;
      st.d2h_bytes += (int64_t)nC * 4;
      static const bool dbgK3 = getenv("RMGPU_DEBUG_K3") != nullptr;
      if (dbgK3) {
        int maxN = 0, failed = 0, second = 0;
        int64_t maxBand = 0, longBand = 0, cells = 0, rows = 0;
        std::vector<int32_t> hst(nC);
        CUDA_CHECK(cudaMemcpy(hst.data(), d_status.p, nC * sizeof(int32_t), cudaMemcpyDeviceToHost));
        for (auto& t : tc) {
          const int n = t.i1 - t.i0 + 1;
          if (n > maxN) maxN = n, longBand = band_of(t.id);
          maxBand = std::max(maxBand, band_of(t.id));
          failed += hlen[t.id] <= 0;
          second += hst[t.id] > 1;
          cells += (int64_t)n * band_of(t.id);
          rows += n;
        }
        fprintf(stderr, "[K3 dbg] %s launch: %zu candidates, %.2f ms, longest %d rows (band %lld), widest band %lld, %d to retry, %d needed the doubled band, "
                "%.1f M band cells, %lld rows, dir bytes %.1f MB\n",
                wide ? "block" : "warp", tc.size(), tm.ms(), maxN, (long long)longBand, (long long)maxBand, failed, second, cells / 1e6, (long long)rows, dirTot / 1e6);
      }
      for (auto& t : tc) {
        const int id = t.id;
        if (hlen[id] > 0) continue;
        const Candidate& c = cands[id];
        int full = std::max(c.i1 - c.i0 + 1, c.j1 - c.j0 + 1);
        if (W[id] >= full) throw Error("traceback could not reproduce the forward score");
        W[id] <<= kTraceAttempts;
        next.push_back(id);
      }
    }
    todo.swap(next);
  }
}

// K4 + K5: rows, counts, adjusted score and Kimura counts per candidate; '*' and masklevel per query record; the
// survivors' rows packed.  Downloads the hit records and the packed rows.
void rmgpu_session::run_finalize(const rmgpu_params& p, const std::vector<Candidate>& cands, std::vector<HitDev>& hits,
                                 std::vector<char>& rows, rmgpu_stats& st) {
  const int nC = (int)cands.size(), nQ = (int)queries.seqs.size();
  hits.clear();
  rows.clear();
  if (nC == 0) return;
  // ln(k) table of the host for every count an alignment can reach
  {
    int64_t need = 2;
    for (auto& c : cands) need = std::max<int64_t>(need, std::min(c.i1 - c.i0, c.j1 - c.j0) + 2);
    if ((int64_t)h_logtab.size() < need) {
      const size_t old = h_logtab.size();
      h_logtab.resize((size_t)std::max<int64_t>(need, 4096));
      for (size_t k = old; k < h_logtab.size(); k++) h_logtab[k] = k == 0 ? 0.0 : std::log((double)k);
      d_logtab.upload(h_logtab, stream);
    }
  }
  EventTimer tm;
  tm.start(stream);
  RenderArgs r;
  r.cand = d_candAll.p;
  r.nCand = nC;
  r.codes_fwd = d_codes_fwd.p;
  r.codes_rc = d_codes_rc.p;
  r.qoff = d_qoff.p;
  r.qlen = d_qlen.p;
  r.scodes = d_scodes.p;
  r.soff = d_soff.p;
  r.ops = d_ops.p;
  r.opsLen = d_opsLen.p;
  memcpy(r.col_of_canon, matrix.col_of_canon, 16);
  r.nCols = matrix.n;
  r.lambda = matrix.lambda;
  for (int k = 0; k < 16; k++) {
    const bool use = k < matrix.n && matrix.freq[k] > 0 && std::log(matrix.freq[k]) != 0;
    r.freqUse[k] = use;
    r.logfreq[k] = use ? std::log(matrix.freq[k]) : 0.0;
  }
  r.logtab = d_logtab.p;
  r.logtabN = (int)h_logtab.size();
  r.rawMode = p.raw ? 1 : 0;
  r.minscore = p.minscore;
  r.rows = d_rowsOut.ensure((size_t)(2 * opsBytes + 2 * (int64_t)nC + 16));
  r.out = d_hits.ensure((size_t)nC);
  render_alignment_kernel<<<(nC + kRenderWarps - 1) / kRenderWarps, kRenderWarps * 32, 0, stream>>>(r);
  CUDA_CHECK(cudaGetLastError());
  // per query record: candidates are sorted by query, so each record owns a contiguous range
  std::vector<int32_t> qFirst(nQ, 0), qCount(nQ, 0);
  std::vector<int64_t> coverOff(nQ + 1, 0);
  for (int i = 0; i < nC; i++) qCount[cands[i].q]++;
  int maxCount = 0;
  for (int q = 0; q < nQ; q++) {
    qFirst[q] = q ? qFirst[q - 1] + qCount[q - 1] : 0;
    maxCount = std::max(maxCount, qCount[q]);
    coverOff[q + 1] = coverOff[q] + (p.masklevel < 101 ? ((int64_t)h_qlen[q] + 31) / 32 + 1 : 0);
  }
  FinalizeArgs f;
  d_qFirst.upload(qFirst, stream);
  d_qCount.upload(qCount, stream);
  d_coverOff.upload(coverOff, stream);
  f.hits = d_hits.p;
  f.qFirst = d_qFirst.p;
  f.qCount = d_qCount.p;
  f.qlen = d_qlen.p;
  f.nQueries = nQ;
  f.masklevel = p.masklevel;
  f.order = d_rankOrder.ensure((size_t)nC);
  f.nKept = d_nKept.ensure((size_t)nQ);
  f.cover = d_cover.ensure((size_t)std::max<int64_t>(coverOff[nQ], 1));
  f.coverOff = d_coverOff.p;
  CUDA_CHECK(cudaMemsetAsync(f.nKept, 0, (size_t)nQ * sizeof(int32_t), stream));
  star_rank_kernel<<<dim3((unsigned)((maxCount + 255) / 256), (unsigned)nQ), 256, 0, stream>>>(f);
  CUDA_CHECK(cudaGetLastError());
  st.launches += 2;
  if (p.masklevel < 101) {
    CUDA_CHECK(cudaMemsetAsync(f.cover, 0, (size_t)coverOff[nQ] * sizeof(uint32_t), stream));
    masklevel_kernel<<<(nQ + 3) / 4, 128, 0, stream>>>(f);
    CUDA_CHECK(cudaGetLastError());
    st.launches += 1;
  }
  GatherArgs g;
  g.hits = d_hits.p;
  g.nCand = nC;
  g.rows = d_rowsOut.p;
  g.packed = d_rowsPacked.ensure((size_t)(2 * opsBytes + 2 * (int64_t)nC + 16));
  g.total = d_rowsTotal.ensure(1);
  scan_rows_kernel<<<1, 1024, 0, stream>>>(g);
  if (p.want_alignments) gather_rows_kernel<<<(nC + 7) / 8, 256, 0, stream>>>(g);
  CUDA_CHECK(cudaGetLastError());
  st.launches += p.want_alignments ? 2 : 1;
  tm.stop(stream);
  hits.resize((size_t)nC);
  long long total = 0;
  CUDA_CHECK(cudaMemcpyAsync(hits.data(), d_hits.p, (size_t)nC * sizeof(HitDev), cudaMemcpyDeviceToHost, stream));
  CUDA_CHECK(cudaMemcpyAsync(&total, g.total, sizeof(total), cudaMemcpyDeviceToHost, stream));
  CUDA_CHECK(cudaStreamSynchronize(stream));
  st.d2h_bytes += (int64_t)nC * (int64_t)sizeof(HitDev) + 8;
  if (p.want_alignments && total > 0) {
    rows.resize((size_t)total);
    CUDA_CHECK(cudaMemcpyAsync(rows.data(), g.packed, (size_t)total, cudaMemcpyDeviceToHost, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));
    st.d2h_bytes += total;
  }
  ms_finalize = tm.ms();
}

namespace {
void fmt_pct(char* out, long num, long den) { snprintf(out, 8, "%.2f", den > 0 ? 100.0 * (double)num / (double)den : 0.0); }
}  // namespace

// How many query records one pass may hold: the K1 column checkpoints (kHotBlocks x rows x 8 B per query) and the stripe
// boundary ring (up to stripes x columns x 8 B per query) must fit in a share of the free HBM.
int rmgpu_session::query_capacity() const {
  static const char* capEnv = getenv("RMGPU_QUERY_CHUNK");
  if (capEnv && atoi(capEnv) > 0) return atoi(capEnv);
  // free HBM = what was free when the session opened less what this process has allocated since (cudaMemGetInfo
  // itself costs up to milliseconds per call)
  const long long freeNow = freeAtOpen - (g_devbuf_bytes[device & 63].load() - devbufAtOpen);
  const size_t freeB = (size_t)std::max<long long>(freeNow, 0);
  const double budget = 0.45 * (double)(freeB + d_ck.cap * sizeof(uint2) + slot[0].d_bnd.cap * 8);
  const double stride = (double)maxQ + 64.0;
  const double ckPerQuery = (double)kHotBlocks * nGroups * R * 8.0;
  // boundary ring: at most 15 tag generations per launch (run_score), i.e. nQ * nStripes / 14 slots, and never fewer than
  // a few per resident warp and per chain of a ticket-order group
  const double ringPerQuery = (double)nStripes / 14.0 * stride * 8.0;
  const double ringFixed = 12.0 * (double)numSMs * 16.0 * stride * 8.0;
  // (query, strand, subject) triples are numbered in 24 bits by the origin pass
  const double byTriples = std::floor(((double)(1u << 24) - 1.0) / std::max(2.0 * (double)lib.seqs.size(), 1.0));
  const double withCk = std::min((budget - ringFixed) / std::max(ckPerQuery + ringPerQuery, 1.0), (double)kCheckpointBudget / std::max(ckPerQuery, 1.0));
  double cap = withCk;
  if (withCk < 8.0) {
    // A library so large that K1's column checkpoints (1 KB per library row and query record) would leave only a few
    // records per pass: plan without them.  run_score() then disables them (they do not fit the budget) and the origin
    // pass sweeps the passing triples from column 0 (swg_origin_kernel), which at such a library is a few per cent of K1.
    const double noCk = std::min(32.0, (budget - ringFixed) / std::max(ringPerQuery, 1.0));
    cap = std::max(withCk, noCk);
  }
  return (int)std::max(1.0, std::floor(std::min(cap, byTriples)));
}

rmgpu_result* rmgpu_session::search(const rmgpu_params& p) {
  if (!haveMatrix || !haveLib || !haveQueries) throw Error("matrix, library and queries must be set before searching");
  const int nQ = (int)queries.seqs.size();
  const int cap = query_capacity();
  if (nQ <= cap) return search_batch(p);
  // too many query records for one pass: run the pipeline chunk by chunk and merge (records are independent units)
  SeqSet all = std::move(queries);
  auto* res = new rmgpu_result();
  memset(res, 0, sizeof(*res));
  auto* arena = new ResultArena();
  res->opaque = arena;
  std::vector<rmgpu_hit> hits;
  try {
    for (int q0 = 0; q0 < nQ; q0 += cap) {
      const int q1 = std::min(nQ, q0 + cap);
      SeqSet sub;
      sub.names.assign(all.names.begin() + q0, all.names.begin() + q1);
      sub.seqs.assign(all.seqs.begin() + q0, all.seqs.begin() + q1);
      set_queries(std::move(sub));
      rmgpu_result* part = search_batch(p);
      for (int i = 0; i < part->n_hits; i++) {
        rmgpu_hit h = part->hits[i];
        h.query_index += q0;
        hits.push_back(h);
      }
      const rmgpu_stats& a = part->stats;
      rmgpu_stats& b = res->stats;
      b.cells += a.cells; b.ms_total += a.ms_total; b.ms_h2d += a.ms_h2d; b.ms_d2h += a.ms_d2h; b.ms_score += a.ms_score;
      b.ms_origin += a.ms_origin; b.ms_traceback += a.ms_traceback; b.ms_host_post += a.ms_host_post;
      b.h2d_bytes += a.h2d_bytes; b.d2h_bytes += a.d2h_bytes; b.launches += a.launches; b.score_launches += a.score_launches;
      b.n_pairs_passed += a.n_pairs_passed; b.n_candidates += a.n_candidates; b.origin_cells += a.origin_cells;
      b.n_origin_redo += a.n_origin_redo;
      arena->children.emplace_back(reinterpret_cast<ResultArena*>(part->opaque));
      delete[] part->hits;
      delete part;
    }
  } catch (...) {
    queries = std::move(all);
    haveQueries = false;  // the device holds a partial batch: force a fresh set_queries
    delete arena;
    delete res;
    throw;
  }
  set_queries(std::move(all));
  res->n_hits = (int32_t)hits.size();
  res->hits = hits.empty() ? nullptr : new rmgpu_hit[hits.size()];
  if (!hits.empty()) memcpy(res->hits, hits.data(), hits.size() * sizeof(rmgpu_hit));
  return res;
}

rmgpu_result* rmgpu_session::search_batch(const rmgpu_params& p) {
  if (!haveMatrix || !haveLib || !haveQueries) throw Error("matrix, library and queries must be set before searching");
  if (p.minscore < 1) throw Error("minscore must be >= 1");
  const double t_begin = now_ms();
  auto* res = new rmgpu_result();
  memset(res, 0, sizeof(*res));
  rmgpu_stats& st = res->stats;
  st.h2d_bytes = h2d_bytes_queries;
  st.ms_h2d = ms_h2d_queries;
  const int nQ = (int)queries.seqs.size(), nS = (int)lib.seqs.size();
  {
    double qb = 0, sb = 0;
    for (int q = 0; q < nQ; q++) qb += (double)queries.seqs[q].size();
    for (int s = 0; s < nS; s++) sb += (double)lib.seqs[s].size();
    st.cells = 2.0 * qb * sb;
  }

  struct Part {  // (one part: the two-sub-batch pipeline of round 2 was measured slower)
    int q0, q1;
    std::vector<HitDev> dh;
    std::vector<char> rows;
    const char* rowBase = nullptr;
  };
  std::vector<Part> parts(1);
  parts[0].q0 = 0;
  parts[0].q1 = nQ;
  std::vector<int32_t> pairbest((size_t)nQ * nS * 2, 0);
  std::vector<unsigned long long> hotmask((size_t)nQ * nS * 2 * kHotWords, 0);
  const double w0 = now_ms();
  double wOrigin = 0, wTrace = 0, wFinal = 0, msFinal = 0;
  if (nQ > 0 && nS > 0) {
    prepare_score(p, /*exact=*/false);
    for (size_t k = 0; k < parts.size(); k++)
      launch_score(p, slot[k], parts[k].q0, parts[k].q1 - parts[k].q0, /*exact=*/false);
    st.launches += 2 * (int)parts.size();
    st.score_launches = (int)parts.size();
    st.d2h_bytes += (int64_t)pairbest.size() * 12;
    for (size_t k = 0; k < parts.size(); k++) {
      Part& P = parts[k];
      collect_score(slot[k], pairbest, hotmask);
      const double t1 = now_ms();
      std::vector<Candidate> cands;
      run_origin(p, pairbest, hotmask, P.q0, P.q1, cands, st);
      const double t2 = now_ms();
      run_traceback(p, cands, st);
      const double t3 = now_ms();
      run_finalize(p, cands, P.dh, P.rows, st);
      wOrigin += t2 - t1;
      wTrace += t3 - t2;
      wFinal += now_ms() - t3;
      msFinal += ms_finalize;
    }
    // the score pass as one span: first launch's start to last launch's end (the launches overlap)
    float f = 0;
    cudaEventElapsedTime(&f, slot[0].evStart, slot[parts.size() - 1].evStop);
    st.ms_score = f;
  }
  const double w4 = now_ms();

  // What is left for the host: the output order, the decimal text of three percentages and of the Kimura divergence
  // (one logarithm per hit), names and pointers.  No per-column work.
  const double t_post = now_ms();
  struct Key {
    int part, idx, qstart, qend;
  };
  std::vector<Key> keep;
  for (size_t pi = 0; pi < parts.size(); pi++)
    for (size_t i = 0; i < parts[pi].dh.size(); i++) {
      const HitDev& h = parts[pi].dh[i];
      if (!(h.flags & kHitKeep) || (h.flags & kHitDrop)) continue;
      const int Q = h_qlen[h.q];
      Key k;
      k.part = (int)pi;
      k.idx = (int)i;
      k.qstart = h.strand ? Q + 1 - h.i1 : h.i0;
      k.qend = h.strand ? Q + 1 - h.i0 : h.i1;
      keep.push_back(k);
    }
  // output order: query order, ascending query start, DESCENDING subject order, then subject start, query end, strand
  std::sort(keep.begin(), keep.end(), [&](const Key& x, const Key& y) {
    const HitDev &a = parts[x.part].dh[x.idx], &b = parts[y.part].dh[y.idx];
    if (a.q != b.q) return a.q < b.q;
    if (x.qstart != y.qstart) return x.qstart < y.qstart;
    if (a.subject != b.subject) return a.subject > b.subject;
    if (a.j0 != b.j0) return a.j0 < b.j0;
    if (x.qend != y.qend) return x.qend < y.qend;
    if (a.strand != b.strand) return a.strand < b.strand;
    return a.j1 < b.j1;
  });
  auto* arena = new ResultArena();
  res->opaque = arena;
  res->n_hits = (int32_t)keep.size();
  res->hits = keep.empty() ? nullptr : new rmgpu_hit[keep.size()];
  for (auto& P : parts) P.rowBase = arena->keep_buffer(std::move(P.rows));
  std::vector<const char*> qn(nQ, nullptr), sn(nS, nullptr);
  for (size_t k = 0; k < keep.size(); k++) {
    const HitDev& h = parts[keep[k].part].dh[keep[k].idx];
    const char* rowBase = parts[keep[k].part].rowBase;
    rmgpu_hit& o = res->hits[k];
    memset(&o, 0, sizeof(o));
    const int Q = h_qlen[h.q], L = (int)lib.seqs[h.subject].size();
    o.score = h.score;
    o.raw_score = h.raw;
    o.query_index = h.q;
    o.subject_index = h.subject;
    if (!qn[h.q]) qn[h.q] = arena->dup(queries.names[h.q]);
    if (!sn[h.subject]) sn[h.subject] = arena->dup(lib.names[h.subject]);
    o.query_name = qn[h.q];
    o.subject_name = sn[h.subject];
    o.qstart = keep[k].qstart;
    o.qend = keep[k].qend;
    o.qleft = Q - o.qend;
    o.sstart = h.j0;
    o.send = h.j1;
    o.sleft = L - h.j1;
    o.strand = (uint8_t)h.strand;
    o.overlap_star = (h.flags & kHitStar) ? 1 : 0;
    int nsub = h.nsub;
    if (p.want_alignments && h.outOff >= 0) {
      o.qaln = rowBase + h.outOff;
      o.saln = o.qaln + h.len + 1;
      if (queries.exotic[h.q] || lib.exotic[h.subject])
        nsub = respell_exotic(h, const_cast<char*>(o.qaln), const_cast<char*>(o.saln));
    }
    const int span = h.i1 - h.i0 + 1;
    fmt_pct(o.pct_div, nsub, span);
    fmt_pct(o.pct_del, h.ndel, span);
    fmt_pct(o.pct_ins, h.nins, span);
    o.transitions = h.ti;
    o.transversions = h.tv;
    o.ambiguous = h.amb;
    o.gap_inits = h.gap_inits;
    o.gap_bases = h.gap_bases;
    o.pct_kimura[0] = 0;
    if (p.want_alignments) {
      snprintf(o.pct_kimura, sizeof(o.pct_kimura), "%4.2f", kimura_from_counts(h.ktrans, h.ktv, h.wcb));
      o.cpg_sites = h.cpg;
      o.well_characterized = h.wcb;
    }
  }
  st.ms_host_post = now_ms() - t_post;
  st.ms_total = now_ms() - t_begin;
  if (getenv("RMGPU_DEBUG_HOST"))
    fprintf(stderr, "[host dbg] search_batch wall: setup %.2f  device part %.2f (score span %.2f, %zu sub-batch(es); origin %.2f  traceback %.2f (kernels %.2f)  finalize %.2f (kernels %.2f))  post %.2f ms\n",
            w0 - t_begin, w4 - w0, st.ms_score, parts.size(), wOrigin, wTrace, st.ms_traceback, wFinal, msFinal, st.ms_host_post);
  return res;
}

// A record with letters outside the canonical 16 (scored as N, like the oracle does) was spelled with N by the device:
// put the original letters back and recount the mismatching pairs (letters, not codes, are compared).
int rmgpu_session::respell_exotic(const HitDev& h, char* qrow, char* srow) const {
  const std::string& qf = queries.seqs[h.q];
  const std::string& sj = lib.seqs[h.subject];
  const int Q = (int)qf.size();
  int qi = h.i0 - 1, si = h.j0 - 1, nsub = 0;
  for (int c = 0; c < h.len; c++) {
    if (qrow[c] != '-') {
      qrow[c] = h.strand ? complement_letter(qf[Q - 1 - qi]) : qf[qi];
      qi++;
    }
    if (srow[c] != '-') srow[c] = sj[si++];
    if (qrow[c] != '-' && srow[c] != '-' && qrow[c] != srow[c]) nsub++;
  }
  return nsub;
}

// -------------------------------------------------------------------------------------------------------------
// C ABI
// -------------------------------------------------------------------------------------------------------------
template <typename F>
static int guarded(F&& f) {
  try {
    g_last_error.clear();
    f();
    return 0;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return 1;
  }
}

extern "C" {

const char* rmgpu_version(void) { return "rmgpu version 0.1.0 (B200 sm_100a exhaustive Smith-Waterman-Gotoh)"; }
int rmgpu_abi_version(void) { return RMGPU_ABI_VERSION; }

void rmgpu_kimura_divergence(const char* qaln, const char* saln, int strand, char text[8], double* transitions,
                             int32_t* transversions, int32_t* well_characterized, int32_t* cpg_sites) {
  const size_t len = std::min(strlen(qaln), strlen(saln));
  const KimuraResult k = kimura_divergence(qaln, saln, len, strand != 0);
  snprintf(text, 8, "%4.2f", k.kimura);
  if (transitions) *transitions = k.transitions;
  if (transversions) *transversions = k.transversions;
  if (well_characterized) *well_characterized = k.wellCharacterized;
  if (cpg_sites) *cpg_sites = k.cpgSites;
}
const char* rmgpu_last_error(void) { return g_last_error.c_str(); }

// Host-only test hook: the schedule of K1's column segments, computed by the very functions the kernel calls
// (kernels.cuh: seg_claim_decode, seg_block_range).
int rmgpu_debug_k1_schedule(int32_t n_tickets, int32_t n_seg, int32_t seg_window, int32_t claim, int32_t n_blocks,
                            int32_t seg_min_blocks, int32_t level, int32_t* ticket, int32_t* seg, int32_t* bs, int32_t* be) {
  if (n_tickets <= 0 || n_seg <= 0 || seg_window <= 0 || claim < 0 || (int64_t)claim >= (int64_t)n_tickets * n_seg) return -1;
  int t = claim, sg = 0, b0 = 0, b1 = n_blocks;
  if (n_seg > 1) seg_claim_decode(claim, n_tickets, n_seg, seg_window, t, sg);
  const int work = n_seg > 1 ? (seg_block_range(n_blocks, n_seg, sg, seg_min_blocks, level, b0, b1) ? 1 : 0) : 1;
  *ticket = t;
  *seg = sg;
  *bs = b0;
  *be = b1;
  return work;
}

rmgpu_session* rmgpu_session_open(int device) {
  rmgpu_session* s = nullptr;
  int rc = guarded([&] {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
      throw Error(std::string("no CUDA device: librmgpu has no CPU fallback (") + cudaGetErrorString(e) + ")");
    if (device < 0) CUDA_CHECK(cudaGetDevice(&device));
    CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    // the library carries sm_100a SASS only (build.py: -gencode arch=compute_100a,code=sm_100a)
    if (prop.major != 10)
      throw Error("librmgpu is built for sm_100a (B200) only; device " + std::to_string(device) + " is sm_" +
                  std::to_string(prop.major) + std::to_string(prop.minor));
    s = new rmgpu_session();
    s->device = device;
    s->numSMs = prop.multiProcessorCount;
    {
      // the score pass on its own (lowest-priority) stream, so that clearing its ring for the next call runs beside the
      // kernels that follow it
      int lo = 0, hi = 0;
      CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CUDA_CHECK(cudaStreamCreateWithPriority(&s->stream, cudaStreamNonBlocking, hi));
      for (auto& sl : s->slot) {
        CUDA_CHECK(cudaStreamCreateWithPriority(&sl.stream, cudaStreamNonBlocking, lo));
        CUDA_CHECK(cudaEventCreate(&sl.evStart));
        CUDA_CHECK(cudaEventCreate(&sl.evStop));
      }
    }
    {
      size_t freeB = 0, totalB = 0;
      CUDA_CHECK(cudaMemGetInfo(&freeB, &totalB));
      s->freeAtOpen = (long long)freeB;
      s->devbufAtOpen = g_devbuf_bytes[device & 63].load();
    }
    {
      int lo = 0, hi = 0;
      CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CUDA_CHECK(cudaStreamCreateWithPriority(&s->stream2, cudaStreamNonBlocking, hi));
    }
  });
  if (rc) {
    delete s;
    return nullptr;
  }
  return s;
}

void rmgpu_session_close(rmgpu_session* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  if (s->stream) cudaStreamDestroy(s->stream);
  if (s->stream2) cudaStreamDestroy(s->stream2);
  for (auto& sl : s->slot) {
    if (sl.stream) cudaStreamDestroy(sl.stream);
    if (sl.evStart) cudaEventDestroy(sl.evStart);
    if (sl.evStop) cudaEventDestroy(sl.evStop);
  }
  delete s;
}

int rmgpu_session_set_matrix_file(rmgpu_session* s, const char* f) {
  return guarded([&] {
    CUDA_CHECK(cudaSetDevice(s->device));
    s->matrix = ScoreMatrix::from_file(f);
    s->haveMatrix = true;
    // residue codes are matrix-independent: the packed library and batch stay; only the class map, the selector streams
    // and the row profiles are rebuilt (K0, microseconds) at the next search
    s->streamsValid = false;
  });
}

int rmgpu_session_set_library(rmgpu_session* s, int32_t n, const char* const* names, const char* const* seqs,
                              const int32_t* lens) {
  return guarded([&] {
    const double t0 = now_ms();
    CUDA_CHECK(cudaSetDevice(s->device));
    SeqSet ss = SeqSet::from_arrays(n, names, seqs, lens);
    const double t1 = now_ms();
    s->set_library(std::move(ss));
    if (getenv("RMGPU_DEBUG_HOST")) fprintf(stderr, "[host dbg] set_library: copy %.2f ms, pack+upload %.2f ms\n", t1 - t0, now_ms() - t1);
  });
}
int rmgpu_session_set_library_fasta(rmgpu_session* s, const char* f) {
  return guarded([&] {
    CUDA_CHECK(cudaSetDevice(s->device));
    s->set_library(SeqSet::from_fasta(f));
  });
}
int rmgpu_session_set_queries(rmgpu_session* s, int32_t n, const char* const* names, const char* const* seqs,
                              const int32_t* lens) {
  return guarded([&] {
    const double t0 = now_ms();
    CUDA_CHECK(cudaSetDevice(s->device));
    SeqSet ss = SeqSet::from_arrays(n, names, seqs, lens);
    const double t1 = now_ms();
    s->set_queries(std::move(ss));
    if (getenv("RMGPU_DEBUG_HOST")) fprintf(stderr, "[host dbg] set_queries: copy %.2f ms, pack+upload %.2f ms\n", t1 - t0, now_ms() - t1);
  });
}
int rmgpu_session_set_queries_fasta(rmgpu_session* s, const char* f) {
  return guarded([&] {
    CUDA_CHECK(cudaSetDevice(s->device));
    s->set_queries(SeqSet::from_fasta(f));
  });
}

int rmgpu_session_set_queries_fasta_multi(rmgpu_session* s, int32_t n_files, const char* const* files, int32_t* first_record) {
  return guarded([&] {
    CUDA_CHECK(cudaSetDevice(s->device));
    SeqSet all;
    for (int f = 0; f < n_files; f++) {
      first_record[f] = (int32_t)all.seqs.size();
      SeqSet one = SeqSet::from_fasta(files[f]);
      for (size_t i = 0; i < one.seqs.size(); i++) {
        all.names.push_back(std::move(one.names[i]));
        all.seqs.push_back(std::move(one.seqs[i]));
      }
    }
    first_record[n_files] = (int32_t)all.seqs.size();
    s->set_queries(std::move(all));
  });
}

int rmgpu_session_mask_queries(rmgpu_session* s, int32_t n, const int32_t* query_index, const int32_t* qstart, const int32_t* qend) {
  return guarded([&] {
    CUDA_CHECK(cudaSetDevice(s->device));
    s->mask_queries(n, query_index, qstart, qend);
  });
}

int rmgpu_session_search(rmgpu_session* s, const rmgpu_params* p, rmgpu_result** out) {
  *out = nullptr;
  return guarded([&] {
    CUDA_CHECK(cudaSetDevice(s->device));
    *out = s->search(*p);
  });
}

int rmgpu_session_score_only(rmgpu_session* s, const rmgpu_params* p, int32_t* best, double* ms) {
  return guarded([&] {
    CUDA_CHECK(cudaSetDevice(s->device));
    if (!s->haveMatrix || !s->haveLib || !s->haveQueries) throw Error("matrix, library and queries must be set");
    std::vector<int32_t> pb;
    std::vector<unsigned long long> hm;
    double t = 0;
    int launches = 0;
    s->run_score(*p, pb, hm, t, launches, /*exact=*/true);
    if (best) memcpy(best, pb.data(), pb.size() * sizeof(int32_t));
    if (ms) *ms = t;
  });
}

int rmgpu_session_counts(const rmgpu_session* s, int32_t* nq, int32_t* ns, int64_t* qb, int64_t* sb) {
  if (nq) *nq = (int32_t)s->queries.seqs.size();
  if (ns) *ns = (int32_t)s->lib.seqs.size();
  if (qb) *qb = s->queries.total_bases();
  if (sb) *sb = s->lib.total_bases();
  return 0;
}

int rmgpu_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    g_last_error = std::string("no CUDA device: librmgpu has no CPU fallback (") + cudaGetErrorString(e) + ")";
    return 0;
  }
  return n;
}

int64_t rmgpu_session_library_h2d_bytes(const rmgpu_session* s) { return s->h2d_bytes_library; }

int rmgpu_search(const rmgpu_params* p, const char* query_fasta, const char* subject_fasta, const char* matrix_file,
                 rmgpu_result** out) {
  *out = nullptr;
  // RMGPU_SERVER=<unix socket>: hand the search to a persistent rmgpu_server (library resident in HBM, no CUDA context
  // in this process -- RepeatMasker forks one short-lived child per batch, RepeatMasker:1043)
  if (const char* sock = getenv("RMGPU_SERVER"))
    if (*sock) {
      std::string err;
      int rc = rmgpu::remote_search(sock, p, query_fasta, subject_fasta, matrix_file, out, err);
      g_last_error = err;
      return rc;
    }
  rmgpu_session* s = rmgpu_session_open(p->device);
  if (!s) return 1;
  int rc = rmgpu_session_set_matrix_file(s, matrix_file);
  if (!rc) rc = rmgpu_session_set_library_fasta(s, subject_fasta);
  if (!rc) rc = rmgpu_session_set_queries_fasta(s, query_fasta);
  if (!rc) rc = rmgpu_session_search(s, p, out);
  std::string keep = g_last_error;
  rmgpu_session_close(s);
  g_last_error = keep;
  return rc;
}

int rmgpu_server_ping(const char* socket_path) {
  std::string err;
  int rc = rmgpu::remote_command(socket_path, rmgpu::kWirePing, err);
  g_last_error = err;
  return rc;
}
int rmgpu_server_quit(const char* socket_path) {
  std::string err;
  int rc = rmgpu::remote_command(socket_path, rmgpu::kWireQuit, err);
  g_last_error = err;
  return rc;
}

void rmgpu_free(rmgpu_result* r) {
  if (!r) return;
  delete[] r->hits;
  delete reinterpret_cast<ResultArena*>(r->opaque);
  delete r;
}

int rmgpu_matrix_info(const char* matrix_file, int32_t* n, char alphabet[17], double* lambda, double freqs[16]) {
  return guarded([&] {
    const ScoreMatrix m = ScoreMatrix::from_file(matrix_file);
    if (n) *n = m.n;
    if (alphabet) snprintf(alphabet, 17, "%s", m.alphabet.c_str());
    if (lambda) *lambda = m.lambda;
    if (freqs)
      for (int i = 0; i < kMaxAlpha; i++) freqs[i] = i < m.n ? m.freq[i] : 0.0;
  });
}

int rmgpu_adjusted_score(const char* matrix_file, int32_t raw, const char* qaln, const char* saln, int32_t* adjusted) {
  return guarded([&] {
    const ScoreMatrix m = ScoreMatrix::from_file(matrix_file);
    int64_t counts[kMaxAlpha] = {};
    const size_t len = std::min(strlen(qaln), strlen(saln));
    for (size_t i = 0; i < len; i++)
      if (qaln[i] != '-' && saln[i] != '-') counts[m.code_of[(unsigned char)qaln[i]]]++;  // aligned pairs, query letter
    *adjusted = complexity_adjust(raw, counts, m);
  });
}

static int collection_filter(int device, int32_t n, const int32_t* name_id, const int32_t* qstart, const int32_t* qend,
                             const int32_t* score, int32_t masklevel, uint8_t* keep, bool contained) {
  return guarded([&] {
    for (int i = 0; i < n; i++) keep[i] = 1;
    if (n <= 0 || masklevel >= 101) return;  // both filters return at once for masklevel >= 101
    if (device >= 0) CUDA_CHECK(cudaSetDevice(device));
    // the sort is the reference's own (a stable merge sort in perl): by name, then the filter's keys, ties in collection order
    std::vector<int32_t> order((size_t)n);
    std::iota(order.begin(), order.end(), 0);
    if (!contained)
      std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        if (name_id[a] != name_id[b]) return name_id[a] < name_id[b];
        if (score[a] != score[b]) return score[a] > score[b];
        return qend[a] - qstart[a] < qend[b] - qstart[b];
      });
    else
      std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        if (name_id[a] != name_id[b]) return name_id[a] < name_id[b];
        if (qstart[a] != qstart[b]) return qstart[a] < qstart[b];
        return qend[a] > qend[b];
      });
    std::vector<int32_t> gFirst, gCount;
    for (int i = 0; i < n; i++) {
      if (i == 0 || name_id[order[i]] != name_id[order[i - 1]]) {
        gFirst.push_back(i);
        gCount.push_back(0);
      }
      gCount.back()++;
    }
    cudaStream_t st;
    CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    DevBuf<int32_t> dStart, dEnd, dScore, dOrder, dFirst, dCount, dDom;
    DevBuf<uint8_t> dDel;
    std::vector<int32_t> vs(qstart, qstart + n), ve(qend, qend + n), vc(score, score + n);
    dStart.upload(vs, st);
    dEnd.upload(ve, st);
    dScore.upload(vc, st);
    dOrder.upload(order, st);
    dFirst.upload(gFirst, st);
    dCount.upload(gCount, st);
    CollectionFilterArgs a;
    a.start = dStart.p;
    a.end = dEnd.p;
    a.score = dScore.p;
    a.order = dOrder.p;
    a.gFirst = dFirst.p;
    a.gCount = dCount.p;
    a.nGroups = (int)gFirst.size();
    a.n = n;
    a.masklevel = masklevel;
    a.domains = dDom.ensure((size_t)n);
    a.deleted = dDel.ensure((size_t)n);
    CUDA_CHECK(cudaMemsetAsync(a.deleted, 0, (size_t)n, st));
    if (!contained)
      perl_masklevel_kernel<<<(a.nGroups + 3) / 4, 128, 0, st>>>(a);
    else
      contained_kernel<<<1, 32, 0, st>>>(a);
    CUDA_CHECK(cudaGetLastError());
    std::vector<uint8_t> del((size_t)n);
    CUDA_CHECK(cudaMemcpyAsync(del.data(), a.deleted, (size_t)n, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    cudaStreamDestroy(st);
    for (int i = 0; i < n; i++) keep[i] = del[i] ? 0 : 1;
  });
}

int rmgpu_masklevel_filter(int device, int32_t n, const int32_t* name_id, const int32_t* qstart, const int32_t* qend,
                           const int32_t* score, int32_t masklevel, uint8_t* keep) {
  return collection_filter(device, n, name_id, qstart, qend, score, masklevel, keep, false);
}
int rmgpu_contained_filter(int device, int32_t n, const int32_t* name_id, const int32_t* qstart, const int32_t* qend,
                           const int32_t* score, int32_t masklevel, uint8_t* keep) {
  return collection_filter(device, n, name_id, qstart, qend, score, masklevel, keep, true);
}

int rmgpu_measure_dpx_peak(int device, double* cups, double* dpx_instr_per_s) {
  return guarded([&] {
    if (device >= 0) CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    int dev = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 20000;
    uint32_t* d = nullptr;
    CUDA_CHECK(cudaMalloc(&d, (size_t)blocks * threads * 4));
    EventTimer tm;
    double bestMs = 1e30;
    for (int rep = 0; rep < 5; rep++) {
      tm.start(0);
      dpx_peak_kernel<8><<<blocks, threads>>>(d, iters, rep);
      tm.stop(0);
      CUDA_CHECK(cudaGetLastError());
      bestMs = std::min(bestMs, tm.ms());
    }
    if (cups) *cups = (double)blocks * threads * 8.0 * iters * 2.0 / (bestMs * 1e-3);
    bestMs = 1e30;
    for (int rep = 0; rep < 5; rep++) {
      tm.start(0);
      dpx_instr_kernel<<<blocks, threads>>>(d, iters);
      tm.stop(0);
      CUDA_CHECK(cudaGetLastError());
      bestMs = std::min(bestMs, tm.ms());
    }
    if (dpx_instr_per_s) *dpx_instr_per_s = (double)blocks * (threads / 32) * 8.0 * iters / (bestMs * 1e-3);
    cudaFree(d);
  });
}

}  // extern "C"
