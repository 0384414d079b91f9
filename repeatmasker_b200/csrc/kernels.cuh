// kernels.cuh -- sm_100a device code of librmgpu.
//
//   K0 build_streams      packed 4-bit query -> per-column PRMT selector stream (both strands) + 1-byte code streams
//   K0 build_profile      packed library rows -> 8-byte per-row score profile for the batch's residue classes
//   K1 swg_score_kernel        strand-paired int16x2 Smith-Waterman-Gotoh score pass (DPX: VIADDMNMX/VIMNMX3/VIADD.16x2):
//                              best per (query, subject, strand), hot column blocks, column checkpoints
//   K2b swg_origin_block_kernel int32 Smith-Waterman-Gotoh with alignment-origin propagation over the hot column blocks
//                              only, started from K1's checkpoints; resolve_candidates_kernel -> candidate table
//   K2 swg_origin_kernel       the same as one sweep over a window of the query (triples K1 cannot checkpoint exactly)
//   K3 traceback_kernel        banded global traceback between origin and end cell
//
// Scoring convention (SearchResult.pm:1829-1852, :1894-1917): score = M[subject][query]; the first base of a gap
// costs gap_init, further bases ins_gap_ext (gap in the subject row, "E") or del_gap_ext (gap in the query row, "F").
#pragma once
#include <type_traits>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace rmgpu {

constexpr int kPadClass = 7;            // residue class whose score is very negative: columns outside the query
constexpr uint32_t kPadSel = 0xF7F7u;   // PRMT selector picking class 7 for both strands (sign-replicated)
constexpr int kWarp = 32;
constexpr int kLag = 4;                 // blocks of slack a K1 consumer stripe rebuilds whenever it catches its producer
constexpr int kLagSweep = 1;            // the same for the latency-bound int32 sweeps of K2, whose pipelines are ~100
                                        // stripes deep: their fill time is lag x depth
constexpr int kHotWords = 2;                // 64-bit words of a hot mask
constexpr int kHotBlocks = 64 * kHotWords;  // column blocks per query record = hot-mask bits = K1 column checkpoints

// ---------------------------------------------------------------------------------------------------------------
// Inter-warp systolic pipeline through global memory.
// A "stripe" is 32 lanes x R subject rows.  A subject longer than one stripe spans consecutive stripes; stripe k+1
// consumes, per query column, the (H, F) of stripe k's last row.  Stripes are claimed through a global ticket
// counter in order, by persistent warps that are all co-resident, so a consumer only ever waits on an older ticket
// that is running or finished (no deadlock).  Progress is published per 32-column block.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// prmt.b32 in its generic form: selector nibble bit 3 replicates the sign of the selected byte.  (__byte_perm masks
// the selector with 0x7777, which would drop exactly that bit.)
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long v;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
  return v;
}
__device__ __forceinline__ uint32_t prmt(uint32_t lo, uint32_t hi, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(lo), "r"(hi), "r"(sel));
  return d;
}
// Every wait in the stripe pipeline is bounded: a wait longer than ~4 s of SM clock can only be a protocol bug, so
// it raises *abort_flag (the host turns that into an error) instead of hanging the GPU.
constexpr long long kWatchdogCycles = 8000000000ll;
__device__ __forceinline__ bool wait_ge(const int* p, int want, int* abort_flag) {
  unsigned ns = 32;
  const long long t0 = clock64();
  while (ld_acquire(p) < want) {
    __nanosleep(ns);
    if (ns < 1024) ns <<= 1;
    if (clock64() - t0 > kWatchdogCycles || ld_acquire(abort_flag) != 0) {
      if (atomicExch(abort_flag, 1) == 0) printf("[watchdog] wait_ge(%p) want %d have %d\n", (const void*)p, want, ld_acquire(p));
      return false;
    }
  }
  return true;
}

// ---------------------------------------------------------------------------------------------------------------
// K0: query streams.  qpacked: 4-bit codes, two per byte (low nibble first), query q at nibble offset qoff[q].
// sel[selOff[q] + i], i in [0, selLen[q]): 16-bit PRMT selector {cf, cf|8, cr, cr|8} (cf/cr = residue class of the
// forward / reverse-complement strand at column i); columns >= Q get kPadSel.
// codes_fwd / codes_rc: matrix-column codes per column (1 byte) for K2/K3.
// ---------------------------------------------------------------------------------------------------------------
struct StreamArgs {
  const uint8_t* qpacked;
  const int64_t* qoff;    // nibble offset of query q (even)
  const int32_t* qlen;
  const int64_t* selOff;  // element offset into sel
  const int32_t* selLen;
  uint16_t* sel;
  uint8_t* codes_fwd;     // [qoff[q] + i]
  uint8_t* codes_rc;
  int nQueries;
  uint8_t class_of_code[16];
  uint8_t comp_code[16];
};

__global__ void build_streams_kernel(StreamArgs a) {
  int q = blockIdx.y;
  int Q = a.qlen[q];
  int n = a.selLen[q];
  int64_t base = a.qoff[q];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint32_t s = kPadSel;
    if (i < Q) {
      int64_t nf = base + i, nr = base + (Q - 1 - i);
      uint8_t cf = (a.qpacked[nf >> 1] >> ((nf & 1) * 4)) & 15;
      uint8_t cr = a.comp_code[(a.qpacked[nr >> 1] >> ((nr & 1) * 4)) & 15];
      a.codes_fwd[base + i] = cf;
      a.codes_rc[base + i] = cr;
      uint32_t kf = a.class_of_code[cf], kr = a.class_of_code[cr];
      s = kf | ((kf | 8u) << 4) | (kr << 8) | ((kr | 8u) << 12);
    }
    a.sel[a.selOff[q] + i] = (uint16_t)s;
  }
}

// K0: profile.  rowcode[g*R + r] = matrix row (subject residue) or 0xFF for pad rows.
// prof[g*R + r] = 8 signed bytes: score of that subject residue against residue classes 0..6, class 7 = pad.
struct ProfileArgs {
  const uint8_t* rowcode;
  uint2* prof;
  int64_t nRows;
  int8_t class_score[16][8];  // [subject code][class]; for merged classes: max over member codes (upper bound)
};

__global__ void build_profile_kernel(ProfileArgs a) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.nRows; i += (int64_t)gridDim.x * blockDim.x) {
    uint8_t c = a.rowcode[i];
    uint32_t lo, hi;
    if (c == 0xFF) {
      lo = 0x80808080u;
      hi = 0x80808080u;
    } else {
      const int8_t* s = a.class_score[c];
      lo = (uint8_t)s[0] | ((uint32_t)(uint8_t)s[1] << 8) | ((uint32_t)(uint8_t)s[2] << 16) | ((uint32_t)(uint8_t)s[3] << 24);
      hi = (uint8_t)s[4] | ((uint32_t)(uint8_t)s[5] << 8) | ((uint32_t)(uint8_t)s[6] << 16) | (0x80u << 24);
    }
    a.prof[i] = make_uint2(lo, hi);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// K0m: mask ranges of the RESIDENT query batch (what RepeatMasker's stage chain does between two searches of the same
// batch, RepeatMasker:6036-6090: `$seqDB->setSubstr( $name, $begin - 1, $len, 'X' x $len )`).  One block per range; a
// position is one nibble of the packed batch, set with two idempotent atomics (ranges may overlap or share a word).
// ---------------------------------------------------------------------------------------------------------------
struct MaskArgs {
  uint32_t* qpacked;        // the packed batch viewed as 32-bit words (8 nibbles each)
  const int64_t* qoff;      // nibble offset of each record
  const int32_t* ranges;    // [n * 3]: record, first position (0-based), end position (exclusive)
  int n;
  uint32_t code;            // canonical code written (X)
};

__global__ void mask_queries_kernel(MaskArgs a) {
  for (int r = blockIdx.x; r < a.n; r += gridDim.x) {
    const int q = a.ranges[3 * r], p0 = a.ranges[3 * r + 1], p1 = a.ranges[3 * r + 2];
    const int64_t base = a.qoff[q];
    for (int p = p0 + (int)threadIdx.x; p < p1; p += (int)blockDim.x) {
      const int64_t nib = base + p;
      const int sh = (int)(nib & 7) * 4;
      atomicAnd(&a.qpacked[nib >> 3], ~(0xFu << sh));
      atomicOr(&a.qpacked[nib >> 3], a.code << sh);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// K1: score pass.  Each lane owns R consecutive subject rows (a "group"); the two 16-bit halves of every register
// are the forward and the reverse-complement strand of the same query (identical dimensions, so they pair
// perfectly).  With E' = E - gap_init and F' = F - gap_init every state value is >= 0:
//     E'[i][j]  = max(E'[i-1][j] + ins_ext, H[i-1][j])                 VIADDMNMX.S16x2
//     T[i][j]   = max(0, H[i-1][j-1] + s, E'[i][j] + gap_init)         PRMT, VIADD.16x2, VIADDMNMX.S16x2.RELU
//     H[i][j]   = max(T[i][j], F'[i][j] + gap_init)                    VIADDMNMX.S16x2
//     F'[i][j+1]= max(F'[i][j] + max(del_ext, gap_init), T[i][j])      VIADDMNMX.S16x2
// (the last line is F'[j+1] = max(F'[j]+del_ext, H[j]) with H expanded; it makes the F' recurrence the ONLY serial
// dependency down a column -- one instruction per row -- so a single warp keeps the DPX pipe busy)
// Lane t works on column (step - t): an anti-diagonal wavefront; the last row's (H, F') and the column's selector
// move one lane down per step by warp shuffle.  int16 overflow: a triple whose best comes near 32767 is reported as
// 32767 and swept in int32 by K2 like every triple with best >= minscore, so wrapped values never reach a result.
// ---------------------------------------------------------------------------------------------------------------
struct ScoreArgs {
  const uint2* prof;          // [nGroups * R]
  const uint8_t* gflags;      // [nGroups] bit0: group starts a subject
  const int32_t* gsubject;    // [nGroups] subject id or -1 (unused lanes)
  const uint8_t* sflags;      // [nStripes] bit0 has_prev, bit1 has_next
  const int32_t* order;       // [nStripes] ticket position -> stripe (level order inside groups, see engine.cu)
  const int32_t* prodPos;     // [nStripes] position of the stripe above (producer), -1 if none
  const int32_t* consPos;     // [nStripes] position of the stripe below (consumer), -1 if none
  const uint16_t* sel;
  const int64_t* selOff;
  const int32_t* qlen;
  const int32_t* hotShift;    // [nQueries] 32-column blocks per hot-mask bit (power of two exponent)
  int nQueries, nStripes, nSubjects;
  int32_t* pairbest;          // [nQueries * nSubjects * 2] raw best per strand (atomicMax)
  unsigned long long* hotmask;  // [nQueries * nSubjects * 2] bit c: some cell of column block c reached minscore
  unsigned long long* bnd;    // [nslot * bndStride] tagged (H, F') of the stripe's last row per column
  int64_t bndStride;
  int nslot;
  int* finished;              // [nTickets]
  int* ticket;                // global ticket counter
  uint32_t gi2, ie2, df2;     // gap_init, ins_gap_ext, max(del_gap_ext, gap_init), replicated in both halves
  int minscore;               // hot-mask threshold (already lowered by the sampling slack in the search path)
  int ovfThresh;              // a best at or above this is reported as 32767 = "int16 may have wrapped, redo in int32"
  int* abort_flag;            // set by the watchdog
  uint2* ck;                  // optional column checkpoints [(q*kHotBlocks + cb) * ckRows + g*R + r] = (H, E') before column cb*CB
  int64_t ckRows;             // nGroups * R
  unsigned long long* dbg;    // optional [8]: cycles waiting on the producer, cycles waiting on a slot, tasks, task cycles,
                              // sum and max over warps of (exit time - first warp's start) in ns, cycles waiting on the previous segment, base time
  // column segments (see swg_score_kernel): a ticket is claimed as nSeg sub-tickets, one after the other
  int nSeg;                   // 1 = whole tickets
  int segWindow;              // tickets per window of the claim order (segment 0 of a window's tickets, then segment 1, ...)
  int segMinBlocks;           // a query record with fewer 32-column blocks than this runs as one segment
  const uint16_t* level;      // [nStripes] ticket position -> depth of the stripe in its chain (0 = no stripe above)
  uint32_t* segState;         // [nTickets][kSegStateWords][32 lanes] a ticket's registers between two of its segments
  int* segDone;               // [nTickets] segments of the ticket that are finished
  unsigned long long* trace;  // optional [claims][4]: ticket * 16 + segment, claim time, start time (after the wait for the
                              // previous segment), end time (globaltimer ns): RMGPU_K1_TRACE=<file>
};

constexpr int kSegStateWords = 40;  // H[16], E[16], best, bestBlk, outH, outF, diag, curSel, nextSel (+1 spare)

// The schedule of the column segments, shared by the kernel and by the host-side test hook (rmgpu_debug_k1_schedule).
// Claim number -> (ticket, segment): windows of segWindow tickets, segment-major inside a window.
__host__ __device__ __forceinline__ void seg_claim_decode(int claim, int nTickets, int nSeg, int segWindow, int& t, int& seg) {
  const int perWin = segWindow * nSeg;
  const int w = claim / perWin, r = claim - w * perWin;
  const int left = nTickets - w * segWindow;
  const int nW = left < segWindow ? left : segWindow;
  seg = r / nW;
  t = w * segWindow + (r - seg * nW);
}
// 32-column blocks [bs, be) of segment seg of a ticket whose stripe lies at depth lvl of its chain: segment s covers
// [s*len - lvl, (s+1)*len - lvl), the first one from 0 and the last one to nBlocks.  A record with fewer than segMinBlocks
// blocks is not cut: its segment 0 is the whole ticket and the others are empty (returns false).
__host__ __device__ __forceinline__ bool seg_block_range(int nBlocks, int nSeg, int seg, int segMinBlocks, int lvl, int& bs, int& be) {
  bs = 0;
  be = nBlocks;
  if (nBlocks < segMinBlocks) return seg == 0;
  const int segLen = (nBlocks + nSeg - 1) / nSeg;
  if (seg > 0) bs = seg * segLen - lvl;
  if (seg < nSeg - 1) be = (seg + 1) * segLen - lvl;
  return true;
}

// Boundary elements are single 64-bit words {H | F' << 32}; every 16-bit half is < 32768, so bit 15 of each half is
// free and carries a 4-bit generation tag: a consumer simply re-reads an element until the tag is the one its
// producer writes (64-bit scalar accesses are single-copy atomic), so no fence, flag or progress counter is needed.
// The ring is cleared (tag 0 = never valid) before every launch.
constexpr unsigned long long kTagMask = 0x8000800080008000ull;
__device__ __forceinline__ unsigned long long tag_bits(int ticket, int nslot) {
  const unsigned tag = (unsigned)((ticket / nslot) % 15) + 1u;
  return ((unsigned long long)(tag & 1u) << 15) | ((unsigned long long)((tag >> 1) & 1u) << 31) |
         ((unsigned long long)((tag >> 2) & 1u) << 47) | ((unsigned long long)((tag >> 3) & 1u) << 63);
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

constexpr int kScoreWarps = 8;
constexpr int kK1Unroll = 8;  // columns per trip of K1's inner loop (4: +3 % time, 16: the loop outgrows the instruction cache)  // columns per trip of K1's inner loop

// EXACT: the running best is folded in at every cell (rmgpu_session_score_only, tests).  !EXACT (the search path, where
// K1 is only a filter): the best is sampled from E' every 8th column.  E'[i][j] = max_k (H[i-k][j] + (k-1)*ins_ext), so
// a sample sees every H of the previous 8 columns of its row less at most 7*|ins_ext|; the host lowers its thresholds
// by that slack (ScoreArgs::minscore already includes it), which keeps the filter free of false negatives and saves
// 6 of the 104 integer-pipe instructions per lane-column.
// Column segments (ScoreArgs::nSeg > 1).  A ticket is 22 ms of work on the bench and a warp gets 8.7 of them: the last of
// the nine waves runs at 74 % occupancy.  With segments a ticket is claimed as nSeg sub-tickets: the warp that finishes
// segment s writes the lanes' registers (kSegStateWords words per lane, each lane in the middle of its own column: the
// wavefront is frozen as it stands) to segState and publishes segDone = s + 1; whoever claims segment s + 1 waits for that,
// reads the registers back and continues the SAME loop at the next 32-column block -- no variant of the hot loop, the
// segment code is all outside it.  Claim order: windows of segWindow tickets, inside a window segment 0 of every ticket,
// then segment 1 of every ticket, ...; a window is at least twice the resident warps, so a segment's predecessor finished
// long before it is claimed, and the work is dealt in quarter tickets.  Deadlock freedom: a stripe of depth l in its chain
// cuts its segments l blocks EARLIER than a stripe of depth 0 (segment s = blocks [s*len - l, (s+1)*len - l)): the
// consumer of block b needs its producer's block b + 1, so segment s of a consumer needs exactly the producer's
// segments <= s, and everything a sub-ticket waits for (own segment s-1, producer's segment s, the old reader of its ring
// slot) was claimed before it: the oldest unfinished sub-ticket can always run.
// GI / IE / DF != 0: gap_init, ins_gap_ext and max(del_gap_ext, gap_init) are compile-time constants, so the four
// VIADDMNMX of a packed cell carry them as 32-bit immediates (two register reads instead of three, three registers
// free); instantiated for the gap penalties of RepeatMasker's crossmatch recipes (RepeatMasker:2146-2366), any other
// penalties run the generic instantiation with the constants in registers.
template <int R, bool EXACT, int GI = 0, int IE = 0, int DF = 0>
__global__ void __launch_bounds__(kScoreWarps * 32, 2) swg_score_kernel(ScoreArgs a) {
  const uint32_t gi2 = GI ? (uint32_t)(uint16_t)GI * 0x10001u : a.gi2;
  const uint32_t ie2 = IE ? (uint32_t)(uint16_t)IE * 0x10001u : a.ie2;
  const uint32_t df2 = DF ? (uint32_t)(uint16_t)DF * 0x10001u : a.df2;
  __shared__ uint4 s_feed[kScoreWarps][33];  // (selector, H above, F' above) for the 32 columns of the current block
                                             // + the selector of the next block's first column
  __shared__ uint2 s_out[kScoreWarps][32];   // last row's (H, F') of the last 32 columns lane 31 finished
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int nTickets = a.nQueries * a.nStripes;
  const int nSeg = a.nSeg;
  const long long nClaims = (long long)nTickets * nSeg;
  if (a.dbg && lane == 0) atomicCAS(&a.dbg[7], 0ull, globaltimer_ns());  // time base: the first warp to get here
  for (;;) {
    int t = 0, seg = 0;
    if (lane == 0) t = atomicAdd(a.ticket, 1);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= nClaims) break;
    const int claim = t;
    if (a.trace && lane == 0) a.trace[(int64_t)claim * 4 + 1] = globaltimer_ns();
    if (nSeg > 1) seg_claim_decode(claim, nTickets, nSeg, a.segWindow, t, seg);
    const int pos = t / a.nQueries, q = t - pos * a.nQueries;  // query fastest: a level of the order spans all queries
    const int k = a.order[pos];
    const int g = k * 32 + lane;
    const int Q = a.qlen[q];
    const uint8_t sf = a.sflags[k];
    const bool hasPrev = sf & 1, hasNext = sf & 2;
    const bool startFlag = a.gflags[g] & 1;
    const uint16_t* sel = a.sel + a.selOff[q];
    uint32_t Plo[R], Phi[R], H[R], E[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      uint2 p = __ldg(&a.prof[(int64_t)g * R + r]);
      Plo[r] = p.x;
      Phi[r] = p.y;
      H[r] = 0;
      E[r] = 0;
    }
    const int qBlocks = (Q + 31) >> 5;
    const int nBlocks = qBlocks + 1;  // lane 31 reaches column qBlocks*32 >= Q
    // this sub-ticket's blocks [bs, be)
    int bs = 0, be = nBlocks;
    if (nSeg > 1 && !seg_block_range(nBlocks, nSeg, seg, a.segMinBlocks, a.level[pos], bs, be)) continue;
    const bool lastSeg = be == nBlocks;
    const int tProd = hasPrev ? a.prodPos[pos] * a.nQueries + q : 0;  // ticket of the stripe above (always older)
    const unsigned long long* bndIn = a.bnd + (int64_t)(tProd % a.nslot) * a.bndStride;
    unsigned long long* bndOut = a.bnd + (int64_t)(t % a.nslot) * a.bndStride;
    const unsigned long long tagIn = tag_bits(tProd, a.nslot), tagOut = tag_bits(t, a.nslot);
    if (hasNext && bs == 0) {
      // The slot still holds the boundary of the last ticket that WROTE it (t - nslot, t - 2 nslot, ... skipping
      // tickets whose stripe has no stripe below): its reader must be done before this ticket overwrites it.
      for (int tOld = t - a.nslot; tOld >= 0; tOld -= a.nslot) {
        const int pOld = tOld / a.nQueries;
        const int cOld = a.consPos[pOld];
        if (cOld < 0) continue;
        const long long w0 = clock64();
        if (lane == 0) wait_ge(&a.finished[cOld * a.nQueries + (tOld - pOld * a.nQueries)], 1, a.abort_flag);
        __syncwarp();
        if (a.dbg && lane == 0) atomicAdd(&a.dbg[1], (unsigned long long)(clock64() - w0));
        break;
      }
    }
    if (a.trace && lane == 0) a.trace[(int64_t)claim * 4 + 0] = (unsigned long long)t * 16 + seg;
    const long long task0 = clock64();
    long long waitCyc = 0;
    const int hotShift = a.hotShift[q];
    const int ckBlkMask = (1 << hotShift) - 1;
    const int sj = a.gsubject[g];
    uint32_t best = 0, bestBlk = 0, outH = 0, outF = 0, diag = 0, curSel = kPadSel;
    uint32_t nextSel = lane == 0 ? (uint32_t)sel[0] : kPadSel;  // selector of this lane's next column, one step ahead
    // 0/1 factors, made opaque so that they stay multiplications
    uint32_t keepSel = lane != 0, feedSel = lane == 0, keepHF = lane != 0 && !startFlag, feedHF = lane == 0 && !startFlag;
    asm volatile("" : "+r"(keepSel), "+r"(feedSel), "+r"(keepHF), "+r"(feedHF));
    if (nSeg > 1) {
      if (bs > 0) {
        const uint32_t* segSt = a.segState + (int64_t)t * (kSegStateWords * 32) + lane;
        // the registers the previous segment left behind (written by another SM: read past L1)
        const long long w0 = clock64();
        if (lane == 0) wait_ge(&a.segDone[t], seg, a.abort_flag);
        __syncwarp();
        if (a.dbg && lane == 0) atomicAdd(&a.dbg[6], (unsigned long long)(clock64() - w0));
        __threadfence();
#pragma unroll
        for (int r = 0; r < R; r++) {
          H[r] = __ldcg(segSt + r * 32);
          E[r] = __ldcg(segSt + (R + r) * 32);
        }
        best = __ldcg(segSt + (2 * R + 0) * 32);
        bestBlk = __ldcg(segSt + (2 * R + 1) * 32);
        outH = __ldcg(segSt + (2 * R + 2) * 32);
        outF = __ldcg(segSt + (2 * R + 3) * 32);
        diag = __ldcg(segSt + (2 * R + 4) * 32);
        curSel = __ldcg(segSt + (2 * R + 5) * 32);
        nextSel = __ldcg(segSt + (2 * R + 6) * 32);
        // lane 31 had finished column bs*32 when the previous segment stopped: the pending entry of the output line
        if (lane == 31) s_out[wib][0] = make_uint2(outH, outF);
        __syncwarp();
      }
    }
    if (a.trace && lane == 0) a.trace[(int64_t)claim * 4 + 2] = globaltimer_ns();
    unsigned long long nextRaw = 0;
    if (hasPrev && bs < qBlocks) nextRaw = ld_relaxed_u64(&bndIn[bs * 32 + lane]);
    for (int b = bs; b < be; b++) {
      uint32_t selv = sel[b * 32 + lane];
      unsigned long long raw = 0;
      if (hasPrev && b < qBlocks) {
        raw = nextRaw;
        if (!__all_sync(0xffffffffu, (raw & kTagMask) == tagIn)) {
          // the prefetch was too early: look again before deciding that the producer really is not ahead
          raw = ld_relaxed_u64(&bndIn[b * 32 + lane]);
          if (!__all_sync(0xffffffffu, (raw & kTagMask) == tagIn)) {
            const long long w0 = clock64();
            // caught up with the producer: sleep until it is kLag blocks ahead (hysteresis), then take block b
            const int far = min(min(b + kLag, qBlocks - 1), be - 1);  // never beyond what this segment itself will need
            unsigned ns = 128;
            bool dead = false;
            for (;;) {
              unsigned long long probe = ld_relaxed_u64(&bndIn[far * 32 + lane]);
              if (__all_sync(0xffffffffu, (probe & kTagMask) == tagIn)) break;
              __nanosleep(ns);
              if (ns < 1024) ns <<= 1;
              if (clock64() - w0 > kWatchdogCycles || ld_acquire(a.abort_flag) != 0) {
                dead = true;
                break;
              }
            }
            dead = __any_sync(0xffffffffu, dead);
            while (!dead) {
              raw = ld_relaxed_u64(&bndIn[b * 32 + lane]);
              if (__all_sync(0xffffffffu, (raw & kTagMask) == tagIn)) break;
              dead = __any_sync(0xffffffffu, clock64() - w0 > kWatchdogCycles);
            }
            if (dead && lane == 0) {
              if (atomicExch(a.abort_flag, 1) == 0)
                printf("[K1 watchdog] ticket %d (q %d pos %d stripe %d) block %d/%d waits on producer ticket %d slot %d tagIn %llx raw %llx far %d\n", t, q, pos, k, b, qBlocks, tProd, tProd % a.nslot, tagIn, raw & kTagMask, far);
            }
            waitCyc += clock64() - w0;
          }
        }
        raw &= ~kTagMask;
        if (b + 1 < qBlocks) nextRaw = ld_relaxed_u64(&bndIn[(b + 1) * 32 + lane]);  // prefetch, validated next block
      }
      // lane l reaches column b*32 at step j == l; that column starts a checkpointed block iff b is a multiple of CB/32
      const bool ckBlock = a.ck != nullptr && b > 0 && (b & ckBlkMask) == 0 && b * 32 < Q;
      __syncwarp();  // everyone is done reading the previous block's feed
      s_feed[wib][lane] = make_uint4(selv, (uint32_t)raw, (uint32_t)(raw >> 32), 0u);
      if (lane == 0) s_feed[wib][32].x = b + 1 < nBlocks ? (uint32_t)sel[(b + 1) * 32] : kPadSel;
      __syncwarp();
      // The 32 columns of the block.  CK: this block starts a checkpointed column block (rare) -- a separate
      // instantiation keeps the test out of the common loop.
      auto columns = [&](auto ckTag) {
        constexpr bool CK = decltype(ckTag)::value;
#pragma unroll kK1Unroll
        for (int j = 0; j < 32; j++) {
          // the selector does not depend on the sweep: it is passed down the lanes one step AHEAD, so a step never starts
          // by waiting for a shuffle
          curSel = nextSel;
          uint32_t inSel = __shfl_up_sync(0xffffffffu, curSel, 1);
          uint32_t inH = __shfl_up_sync(0xffffffffu, outH, 1);
          uint32_t inF = __shfl_up_sync(0xffffffffu, outF, 1);
          const uint4 fd = s_feed[wib][j];  // broadcast read
          const uint32_t fdSelNext = s_feed[wib][j + 1].x;
          // lane 0 takes the feed, a lane that starts a subject takes zeros: as multiply-adds by 0/1 factors, which
          // issue on the FMA pipe instead of five SELs on the (saturated) integer pipe
          nextSel = inSel * keepSel + fdSelNext * feedSel;
          inH = inH * keepHF + fd.y * feedHF;
          inF = inF * keepHF + fd.z * feedHF;
          if (CK && j == lane) {
            // column checkpoint for the origin pass: the (H, E') this lane holds before it starts column b*32 = cb*CB
            uint2* dst = a.ck + ((int64_t)q * kHotBlocks + (b >> hotShift)) * a.ckRows + (int64_t)g * R;
#pragma unroll
            for (int r = 0; r < R; r++) dst[r] = make_uint2(H[r], E[r]);
          }
          uint32_t hd = diag, f = inF, h = 0;
          diag = inH;
#pragma unroll
          for (int r = 0; r < R; r++) {
            const uint32_t s = prmt(Plo[r], Phi[r], curSel);
            const uint32_t hl = H[r];
            E[r] = __viaddmax_s16x2(E[r], ie2, hl);                    // E'[i][j]
            const uint32_t d = __vadd2(hd, s);                         // H[i-1][j-1] + s
            const uint32_t tt = __viaddmax_s16x2_relu(E[r], gi2, d);    // T = max(0, diag, E'+gi): H without F
            h = __viaddmax_s16x2(f, gi2, tt);                          // H = max(T, F'+gi)          (off the chain)
            f = __viaddmax_s16x2(f, df2, tt);                          // F' of the next row         (the only chain)
            hd = hl;
            H[r] = h;
            if (EXACT) bestBlk = __vmaxs2(bestBlk, h);
          }
          if (!EXACT && (j & 7) == 7) {
#pragma unroll
            for (int r = 0; r < R; r++) bestBlk = __vmaxs2(bestBlk, E[r]);
          }
          outH = h;
          outF = f;
          if (hasNext) {
            // lane 31 has just finished column c = b*32 + j - 31
            const int c = b * 32 + j - 31;
            if (lane == 31) s_out[wib][c & 31] = make_uint2(outH, outF);
            if (c >= 0 && (c & 31) == 31) {  // column block complete: one coalesced, tagged 64-bit store per lane
              __syncwarp();
              const uint2 o = s_out[wib][lane];
              st_relaxed_u64(&bndOut[(c & ~31) + lane],
                             ((((unsigned long long)o.y << 32) | o.x) & ~kTagMask) | tagOut);
              __syncwarp();
            }
          }
        }
      };
      if (ckBlock)
        columns(std::true_type{});
      else
        columns(std::false_type{});
      // column-block bookkeeping for the origin pass: which blocks hold a cell >= minscore
      if ((((b + 1) >> hotShift) != (b >> hotShift)) || b == nBlocks - 1) {
        if (sj >= 0) {
          // per-half compare bestBlk >= minscore
          const bool hf = (int)(int16_t)(bestBlk & 0xffffu) >= a.minscore;
          const bool hr = (int)(int16_t)(bestBlk >> 16) >= a.minscore;
          if (hf | hr) {
            unsigned long long* hm = a.hotmask + ((int64_t)q * a.nSubjects + sj) * 2 * kHotWords;  // [strand][word]
            const int blk = min(b >> hotShift, kHotBlocks - 1);
            const unsigned long long bit = 1ull << (blk & 63);
            if (hf) atomicOr(hm + (blk >> 6), bit);
            if (hr) atomicOr(hm + kHotWords + (blk >> 6), bit);
          }
        }
        best = __vmaxs2(best, bestBlk);
        bestBlk = 0;
      }
    }
    if (!lastSeg) {
      // hand the frozen wavefront to the next segment of this ticket
      uint32_t* segSt = a.segState + (int64_t)t * (kSegStateWords * 32) + lane;
#pragma unroll
      for (int r = 0; r < R; r++) {
        __stcg(segSt + r * 32, H[r]);
        __stcg(segSt + (R + r) * 32, E[r]);
      }
      __stcg(segSt + (2 * R + 0) * 32, best);
      __stcg(segSt + (2 * R + 1) * 32, bestBlk);
      __stcg(segSt + (2 * R + 2) * 32, outH);
      __stcg(segSt + (2 * R + 3) * 32, outF);
      __stcg(segSt + (2 * R + 4) * 32, diag);
      __stcg(segSt + (2 * R + 5) * 32, curSel);
      __stcg(segSt + (2 * R + 6) * 32, nextSel);
      __threadfence();
      __syncwarp();
      if (lane == 0) {
        st_release(&a.segDone[t], seg + 1);
        if (a.trace) a.trace[(int64_t)claim * 4 + 3] = globaltimer_ns();
        if (a.dbg) {
          atomicAdd(&a.dbg[0], (unsigned long long)waitCyc);
          atomicAdd(&a.dbg[2], 1ull);
          atomicAdd(&a.dbg[3], (unsigned long long)(clock64() - task0));
        }
      }
      continue;
    }
    // per (query, subject, strand) best
    if (sj >= 0) {
      int f = (int)(int16_t)(best & 0xffffu), rc = (int)(int16_t)(best >> 16);
      // a best this close to 32767 may hide an int16 wrap: report 32767 so that the triple is redone in int32
      if (f < 0 || f >= a.ovfThresh) f = 32767;
      if (rc < 0 || rc >= a.ovfThresh) rc = 32767;
      int32_t* pb = a.pairbest + ((int64_t)q * a.nSubjects + sj) * 2;
      if (f > 0) atomicMax(pb, f);
      if (rc > 0) atomicMax(pb + 1, rc);
    }
    __syncwarp();
    if (lane == 0) {
      st_release(&a.finished[t], 1);
      if (a.trace) a.trace[(int64_t)claim * 4 + 3] = globaltimer_ns();
      if (a.dbg) {
        atomicAdd(&a.dbg[0], (unsigned long long)waitCyc);
        atomicAdd(&a.dbg[2], 1ull);
        atomicAdd(&a.dbg[3], (unsigned long long)(clock64() - task0));
      }
    }
  }
  if (a.dbg && lane == 0) {
    const unsigned long long dt = globaltimer_ns() - *(volatile unsigned long long*)&a.dbg[7];
    atomicAdd(&a.dbg[4], dt);
    atomicMax(&a.dbg[5], dt);
  }
}

// Register-resident loop of the bare K1 recurrence: the measured integer-SIMD peak used as the roofline denominator
// (SURVEY 8d).  No shuffles, no memory: 32 lanes x R rows x iters columns x 2 strands cells per warp.
template <int R>
__global__ void dpx_peak_kernel(uint32_t* out, int iters, uint32_t seed) {
  uint32_t Plo[R], Phi[R], H[R], E[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    Plo[r] = 0x090af1f0u + r * 0x01010101u + threadIdx.x;
    Phi[r] = 0xf1f60a09u ^ (r * 0x00010001u);
    H[r] = 0;
    E[r] = 0;
  }
  uint32_t best = 0, sel = 0x9180u | (seed & 0x0303u), diag = 0, hu = 0, f = 0;
  const uint32_t gi2 = 0xffe2ffe2u, ie2 = 0xfffafffau, de2 = 0xfffbfffbu;
  for (int it = 0; it < iters; it++) {
    uint32_t hd = diag;
    diag = hu;
    sel = ((sel + 0x0101u) & 0x0303u) | 0x9180u;
#pragma unroll
    for (int r = 0; r < R; r++) {
      uint32_t s = prmt(Plo[r], Phi[r], sel);
      uint32_t hl = H[r];
      E[r] = __viaddmax_s16x2(E[r], ie2, hl);
      uint32_t d = __vadd2(hd, s);
      uint32_t tt = __viaddmax_s16x2_relu(E[r], gi2, d);
      uint32_t h = __viaddmax_s16x2(f, gi2, tt);
      f = __viaddmax_s16x2(f, de2, tt);
      hd = hl;
      H[r] = h;
      hu = h;
      best = __vmaxs2(best, h);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = best;
}

// Pure VIADDMNMX.S16x2 issue rate (8 independent chains per thread).
__global__ void dpx_instr_kernel(uint32_t* out, int iters) {
  uint32_t x[8];
#pragma unroll
  for (int r = 0; r < 8; r++) x[r] = threadIdx.x + r;
  const uint32_t c1 = 0x00010003u, c2 = 0x00050007u;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++) x[r] = __viaddmax_s16x2(x[r], c1, c2 + it);
  }
  uint32_t s = 0;
#pragma unroll
  for (int r = 0; r < 8; r++) s ^= x[r];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---------------------------------------------------------------------------------------------------------------
// K2: origin pass.  int32 Smith-Waterman-Gotoh over one (query, strand, subject) triple, every positive cell carrying
// the origin (i0, j0) of its alignment.  Tie-breaks: E over F over diagonal; extension over opening (the rule that
// reproduces real cross_match alignments, tests/test_oracle_golden.py).  A cell reached by the diagonal with
// H >= minscore is a candidate end: per (task, origin) the table keeps max score, then min i, then min j.
// Same stripe pipeline as K1 with R2 rows per lane.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kR2 = 8;
constexpr uint64_t kEmptyKey = 0xFFFFFFFFFFFFFFFFull;
constexpr uint32_t kOriginLeft = 0xFFFFFFFFu;  // "the origin lies in a part of the matrix swept without origins"

struct OriginTask {
  int32_t q, strand, subject;
  int32_t firstTicket, nStripes;  // stripes of kR2*32 rows
  unsigned long long mode[kHotWords];  // bit c: sweep column block c with origin propagation (else score only)
  int32_t hotShift;               // column block c = ((column >> 5) >> hotShift), capped at kHotBlocks - 1
  int32_t c0, c1;                 // query columns [c0, c1) swept by this task (0-based)
  int32_t pad_;
  int64_t ckOff;                  // offset of the subject's first row in the K1 checkpoint taken before column c0, or -1
};

struct OriginBoundary {  // per column, last row of the stripe above
  int32_t h, f;
  uint32_t oh, of;
};

struct OriginArgs {
  const OriginTask* tasks;
  const int32_t* ticketTask;   // [nTickets] task index
  int nTickets;
  const uint8_t* codes_fwd;    // query columns
  const uint8_t* codes_rc;
  const int64_t* qoff;
  const int32_t* qlen;
  const uint8_t* scodes;       // subject rows (1 byte per base)
  const int64_t* soff;
  const int32_t* slen;
  int32_t mat[16][16];         // [subject][query]
  int32_t gap_init, ins_ext, del_ext, minscore;
  OriginBoundary* bnd;         // [nslot * bndStride]
  int64_t bndStride;
  int nslot;
  int* progress;
  int* finished;
  int* ticket;
  int* abort_flag;
  const uint2* ck;             // K1 column checkpoints (H, E' both strands), see ScoreArgs
  unsigned long long* dbg;     // optional [4]: guard-wait cycles, producer-wait cycles, tickets, ticket cycles
  int* leftFlag;               // [nTasks] set when a hit needs origins from a score-only region: redo that task fully
  unsigned long long* keys;    // open-addressing table
  unsigned long long* vals;
  uint32_t tableMask;
  int* overflow;
};

// An origin (i0, j0) is carried as one 32-bit word (i0 - c0) * (L + 1) + j0 (L = subject length, c0 = first column of the
// task's window); the host checks (Q + 2) * (L + 1) < 2^32 - 1 per sweep (any SimpleBatcher fragment, <= 75 kbp, against
// subjects up to 57 kbp).
// value of a candidate end: larger = better score, then smaller i, then smaller j
__device__ __forceinline__ unsigned long long cand_val(int h, int i, int j) {
  return ((unsigned long long)(uint32_t)h << 40) | ((unsigned long long)(0xFFFFFu - (uint32_t)i) << 20) |
         (unsigned long long)(0xFFFFFu - (uint32_t)j);
}
__device__ __forceinline__ void origin_emit(const OriginArgs& a, int task, uint32_t origin, unsigned long long val) {
  uint64_t key = ((uint64_t)(uint32_t)task << 32) | origin;
  uint32_t slot = (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 32) & a.tableMask;
  for (uint32_t probe = 0; probe < 128u && probe <= a.tableMask; probe++) {
    unsigned long long old = atomicCAS(&a.keys[slot], kEmptyKey, key);
    if (old == kEmptyKey || old == key) {
      atomicMax(&a.vals[slot], val);
      return;
    }
    slot = (slot + 1) & a.tableMask;
  }
  atomicExch(a.overflow, 1);
}

// One origin-propagating cell, branch-free (the tie-breaks of DESIGN.md section 3: E over F over diagonal, extension over
// opening).  Rows beyond the subject and columns beyond the window need no special case: their score is kPadScore, so
// the diagonal never wins there, nothing in them is a candidate end, and what their E / F carry only flows further down
// and right, away from every real cell.
constexpr int kPadCode = 16;                 // residue code of a row / column outside the matrix
constexpr int kPadScore = -(1 << 20);
constexpr int kMatStride = 17;               // score table [17][17]: [subject code][query code], code 16 = pad
struct OriginCell {
  int h;
  uint32_t o;
  bool viaDiag;
};
__device__ __forceinline__ OriginCell origin_cell(int hl, uint32_t ol, int hu, uint32_t ou, int hd, uint32_t od, int s,
                                                  uint32_t fresh, int& E, uint32_t& oE, int& f, uint32_t& of, int gap_init,
                                                  int ins_ext, int del_ext) {
  const int e_ext = E + ins_ext, e_open = hl + gap_init;
  const bool eOpen = e_open > e_ext;
  E = max(e_ext, e_open);
  oE = eOpen ? ol : oE;
  const int f_ext = f + del_ext, f_open = hu + gap_init;
  const bool fOpen = f_open > f_ext;
  f = max(f_ext, f_open);
  of = fOpen ? ou : of;
  const int d = hd + s;
  const uint32_t odd = hd > 0 ? od : fresh;
  const bool c1 = f > E;
  int h = max(E, f);
  uint32_t o = c1 ? of : oE;
  const bool c2 = d > h;
  h = max(h, d);
  o = c2 ? odd : o;
  // a cell that is not positive keeps whatever origin the selects left in it: nothing ever reads it (the diagonal
  // restarts from `fresh` when hd <= 0, and an E or F opened from H = 0 is negative and never wins a positive cell);
  // viaDiag is only meaningful for h >= minscore >= 1
  OriginCell r;
  r.h = max(h, 0);
  r.o = o;
  r.viaDiag = c2;
  return r;
}

__device__ __forceinline__ void load_score_table(int32_t* smat, const int32_t (*mat)[16]) {
  for (int i = threadIdx.x; i < kMatStride * kMatStride; i += blockDim.x) {
    const int sr = i / kMatStride, qc = i - sr * kMatStride;
    smat[i] = (sr < 16 && qc < 16) ? mat[sr][qc] : kPadScore;
  }
}

// Column blocks that K1 saw reach minscore (dilated by one block on either side by the host) are swept with origin
// propagation; all other columns with the plain int32 recurrence (same H/E/F values, about a quarter of the
// instructions).  Entering an origin block, every live path gets the origin kOriginLeft; a candidate carrying it,
// or any cell >= minscore met in score-only mode, flags the task and the host redoes it with mode = all ones.
template <int R>
__global__ void __launch_bounds__(128) swg_origin_kernel(OriginArgs a) {
  __shared__ int32_t smat[kMatStride * kMatStride];
  load_score_table(smat, a.mat);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int NEGI = -(1 << 28);
  for (;;) {
    int t = 0;
    if (lane == 0) t = atomicAdd(a.ticket, 1);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= a.nTickets) break;
    const int ti = a.ticketTask[t];
    const OriginTask tk = a.tasks[ti];
    const int k = t - tk.firstTicket;
    const int Q = a.qlen[tk.q], L = a.slen[tk.subject];
    const uint32_t Lp1 = (uint32_t)L + 1u;
    const uint8_t* qc = (tk.strand ? a.codes_rc : a.codes_fwd) + a.qoff[tk.q];
    const uint8_t* sc = a.scodes + a.soff[tk.subject];
    const bool hasPrev = k > 0, hasNext = k + 1 < tk.nStripes;
    const int row0 = (k * 32 + lane) * R;  // 0-based subject row of this lane's first row
    int srow[R];
#pragma unroll
    for (int r = 0; r < R; r++) srow[r] = ((row0 + r < L) ? (int)sc[row0 + r] : kPadCode) * kMatStride;
    int H[R], E[R];
    uint32_t oH[R], oE[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      H[r] = 0;
      E[r] = NEGI;
      oH[r] = 0;
      oE[r] = 0;
    }
    int diagH = 0;
    if (tk.ckOff >= 0) {
      // window starting at a K1 column checkpoint (exact H, E' before column c0); whatever is alive there comes "from
      // the left": a hit that carries that origin makes the host redo the task from column 0
      const uint2* ck = a.ck + tk.ckOff;
      const int sh = tk.strand ? 16 : 0;
#pragma unroll
      for (int r = 0; r < R; r++)
        if (row0 + r < L) {
          const uint2 v = ck[row0 + r];
          H[r] = (int)((v.x >> sh) & 0xffffu);
          E[r] = (int)((v.y >> sh) & 0xffffu) + a.gap_init;
          oH[r] = kOriginLeft;
          oE[r] = kOriginLeft;
        }
      if (row0 > 0 && row0 - 1 < L) diagH = (int)((ck[row0 - 1].x >> sh) & 0xffffu);
    }
    const long long k2t0 = clock64();
    long long k2wait = 0;
    const int W = tk.c1 - tk.c0;  // columns of this task's window
    const int qBlocks = (W + 31) >> 5, nBlocks = qBlocks + 1;
    const OriginBoundary* bndIn = a.bnd + (int64_t)((t - 1 + a.nslot) % a.nslot) * a.bndStride;
    OriginBoundary* bndOut = a.bnd + (int64_t)(t % a.nslot) * a.bndStride;
    if (hasNext) {
      // the reader (ticket + 1) of the last ticket that wrote this ring slot must be done
      for (int tOld = t - a.nslot; tOld >= 0; tOld -= a.nslot) {
        const OriginTask& ot = a.tasks[a.ticketTask[tOld]];
        if (tOld - ot.firstTicket + 1 >= ot.nStripes) continue;  // that ticket had no stripe below: it wrote nothing
        const long long g0 = clock64();
        if (lane == 0) wait_ge(&a.finished[tOld + 1], 1, a.abort_flag);
        __syncwarp();
        if (a.dbg && lane == 0) atomicAdd(&a.dbg[0], (unsigned long long)(clock64() - g0));
        break;
      }
    }
    int outH = 0, outF = NEGI, curQ = kPadCode;
    uint32_t outOH = 0, outOF = 0, diagO = kOriginLeft;  // only read when the diagonal source is alive (checkpoint start)
    OriginBoundary wb = {0, NEGI, 0, 0};
    uint32_t lastOrigin = kOriginLeft;  // nothing cached
    unsigned long long lastVal = 0;
    bool prevMode = false, needFull = false;
    for (int b = 0; b < nBlocks; b++) {
      int col = b * 32 + lane;  // column offset inside the window
      int qv = (col < W) ? (int)qc[tk.c0 + col] : kPadCode;
      OriginBoundary bv = {0, NEGI, 0, 0};
      if (hasPrev && b < qBlocks) {
        // hysteresis: when the producer is not ahead, let it get kLag blocks ahead so that later blocks need no sleep
        const long long w0 = clock64();
        if (lane == 0 && ld_acquire(&a.progress[t - 1]) < b + 1) wait_ge(&a.progress[t - 1], min(b + kLagSweep, qBlocks), a.abort_flag);
        __syncwarp();
        k2wait += clock64() - w0;
        const int4 raw = __ldcg(reinterpret_cast<const int4*>(&bndIn[col]));
        bv.h = raw.x;
        bv.f = raw.y;
        bv.oh = (uint32_t)raw.z;
        bv.of = (uint32_t)raw.w;
      }
      // mode of this lane's column: lanes lag the block by `lane` columns, so two 32-column blocks are involved
      const int b0 = tk.c0 >> 5;  // windows start on 32-column boundaries
      auto mode_bit = [&](int blk32) {
        const int blk = min(blk32 >> tk.hotShift, kHotBlocks - 1);
        unsigned long long w = tk.mode[0];
#pragma unroll
        for (int k = 1; k < kHotWords; k++) w = (blk >> 6) == k ? tk.mode[k] : w;
        return (bool)((w >> (blk & 63)) & 1ull);
      };
      const bool modeCur = mode_bit(b0 + b);
      const bool modeOld = b > 0 ? mode_bit(b0 + b - 1) : modeCur;
      for (int j = 0; j < 32; j++) {
        int inQ = __shfl_up_sync(0xffffffffu, curQ, 1);
        int inH = __shfl_up_sync(0xffffffffu, outH, 1);
        int inF = __shfl_up_sync(0xffffffffu, outF, 1);
        uint32_t inOH = __shfl_up_sync(0xffffffffu, outOH, 1);
        uint32_t inOF = __shfl_up_sync(0xffffffffu, outOF, 1);
        int fq = __shfl_sync(0xffffffffu, qv, j);
        int fh = __shfl_sync(0xffffffffu, bv.h, j);
        int ff = __shfl_sync(0xffffffffu, bv.f, j);
        uint32_t foh = __shfl_sync(0xffffffffu, bv.oh, j);
        uint32_t fof = __shfl_sync(0xffffffffu, bv.of, j);
        if (lane == 0) {
          inQ = fq;
          inH = fh;
          inF = ff;
          inOH = foh;
          inOF = fof;
        }
        curQ = inQ;
        const int cw = b * 32 + j - lane;  // column inside the window, 0-based
        const int i1 = tk.c0 + cw + 1;     // 1-based query column of this lane at this step
        const bool mode = (j >= lane) ? modeCur : modeOld;
        int hd = diagH, hu = inH, f = inF;
        if (cw < 0) {
          // before its first column a lane keeps its (checkpointed) state untouched
        } else if (!mode) {
          diagH = inH;
          // ---- score-only sweep: identical H/E/F values, no origins
#pragma unroll
          for (int r = 0; r < R; r++) {
            const int hl = H[r];
            const int e = max(E[r] + a.ins_ext, hl + a.gap_init);
            const int fn = max(f + a.del_ext, hu + a.gap_init);
            const int d = hd + smat[srow[r] + curQ];
            const int hnew = max(max(e, fn), max(d, 0));
            E[r] = e;
            f = fn;
            needFull |= hnew >= a.minscore;  // (a pad cell only reaches minscore below / right of a real cell that did)
            hd = hl;
            H[r] = hnew;
            hu = hnew;
          }
          diagO = kOriginLeft;
          outOH = kOriginLeft;
          outOF = kOriginLeft;
          prevMode = false;
        } else {
          diagH = inH;
          if (!prevMode) {  // entering an origin block: everything alive comes from the left
#pragma unroll
            for (int r = 0; r < R; r++) {
              oH[r] = kOriginLeft;
              oE[r] = kOriginLeft;
            }
          }
          uint32_t od = diagO, ou = inOH, of = inOF;
          diagO = inOH;
          const uint32_t freshBase = (uint32_t)(cw + 1) * Lp1 + (uint32_t)(row0 + 1);
          uint32_t ends = 0;  // candidate ends of this column, handled after the straight-line cells (see K2b)
#pragma unroll
          for (int r = 0; r < R; r++) {
            const int hl = H[r];
            const uint32_t ol = oH[r];
            const int sc_ = smat[srow[r] + curQ];
            const OriginCell c = origin_cell(hl, ol, hu, ou, hd, od, sc_, freshBase + (uint32_t)r, E[r], oE[r], f, of, a.gap_init,
                                             a.ins_ext, a.del_ext);
            ends |= (c.viaDiag && c.h >= a.minscore) ? (1u << r) : 0u;
            hd = hl;
            od = ol;
            H[r] = c.h;
            oH[r] = c.o;
            hu = c.h;
            ou = c.o;
          }
          if (ends) {
#pragma unroll
            for (int r = 0; r < R; r++)
              if (ends & (1u << r)) {
                if (oH[r] == kOriginLeft) {
                  needFull = true;
                } else {
                  // per-thread cache: one table update per (origin, lane crossing) instead of one per cell
                  const unsigned long long val = cand_val(H[r], i1, row0 + r + 1);
                  if (oH[r] == lastOrigin) {
                    lastVal = val > lastVal ? val : lastVal;
                  } else {
                    if (lastOrigin != kOriginLeft) origin_emit(a, ti, lastOrigin, lastVal);
                    lastOrigin = oH[r];
                    lastVal = val;
                  }
                }
              }
          }
          outOH = ou;
          outOF = of;
          prevMode = true;
        }
        if (cw == W - 1 && tk.c1 < Q) {
          // the window ends before the query does: nothing >= minscore may leave it (the host then redoes the task
          // over the whole query); everything further right was cold in K1, see run_origin
#pragma unroll
          for (int r = 0; r < R; r++) needFull |= H[r] >= a.minscore;
        }
        outH = hu;
        outF = f;
        if (hasNext) {
          int oh = __shfl_sync(0xffffffffu, outH, 31);
          int of2 = __shfl_sync(0xffffffffu, outF, 31);
          uint32_t ooh = __shfl_sync(0xffffffffu, outOH, 31);
          uint32_t oof = __shfl_sync(0xffffffffu, outOF, 31);
          int c = b * 32 + j - 31;
          if (lane == (c & 31)) {
            wb.h = oh;
            wb.f = of2;
            wb.oh = ooh;
            wb.of = oof;
          }
          if (c >= 0 && (c & 31) == 31) {
            *reinterpret_cast<int4*>(&bndOut[(c & ~31) + lane]) = make_int4(wb.h, wb.f, (int)wb.oh, (int)wb.of);
            // the warp barrier orders every lane's store before lane 0's release store, and a release is cumulative:
            // no fence of its own per lane (a MEMBAR.SC.GPU per 32 columns on the chain's critical path)
            __syncwarp();
            if (lane == 0) st_release(&a.progress[t], (c >> 5) + 1);
          }
        }
      }
    }
    if (lastOrigin != kOriginLeft) origin_emit(a, ti, lastOrigin, lastVal);
    if (__any_sync(0xffffffffu, needFull) && lane == 0) atomicExch(&a.leftFlag[ti], 1);
    __syncwarp();
    if (lane == 0) {
      st_release(&a.finished[t], 1);
      if (a.dbg) {
        atomicAdd(&a.dbg[1], (unsigned long long)k2wait);
        atomicAdd(&a.dbg[2], 1ull);
        atomicAdd(&a.dbg[3], (unsigned long long)(clock64() - k2t0));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// K2w: int32 score pass for the few triples whose int16 values wrapped in K1 (raw scores >= ~32,000).  K1's checkpoints
// and hot mask are exact only left of the wrap, so such a triple could not use the block-parallel origin pass and was
// swept by swg_origin_kernel in one dependent chain of ~170 instructions per column.  This kernel is the cheap part of
// that chain only: the plain H/E/F recurrence in int32 (no origins, ~40 instructions per column at 2 rows per lane)
// over the triple's window, started from K1's last exact checkpoint.  It writes what K1 could not: an exact (H, E)
// checkpoint of every subject row at every column-block boundary of the window, the exact hot mask of the window, and
// a flag when something >= minscore leaves the window on the right.  With those the triple's hot blocks become
// ordinary K2b tasks that run in parallel.  Same stripe pipeline (ring + progress flags) as swg_origin_kernel.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kRW = 2;  // subject rows per lane: the launch is one dependent chain per triple, so the per-column latency counts

struct WideTask {
  int32_t q, strand, subject;
  int32_t firstTicket, nStripes;  // stripes of kRW*32 rows
  int32_t c0, c1;                 // window [c0, c1) of query columns; c0 is a multiple of the checkpoint spacing
  int32_t hotShift;               // column block = (column >> 5) >> hotShift
  int64_t ckOff;                  // subject's first row in K1's (exact) checkpoint taken before column c0, or -1 (c0 == 0)
  int64_t ck32Off;                // ck32[ck32Off + n * L + row] = (H, E) before column c0 + n * CB, n >= 1, CB = 32 << hotShift
};

struct WideArgs {
  const WideTask* tasks;
  const int32_t* ticketTask;
  int nTickets;
  const uint8_t* codes_fwd;
  const uint8_t* codes_rc;
  const int64_t* qoff;
  const int32_t* qlen;
  const uint8_t* scodes;
  const int64_t* soff;
  const int32_t* slen;
  int32_t mat[16][16];
  int32_t gap_init, ins_ext, del_ext, minscore;
  int2* bnd;                     // [nslot * bndStride] (H, F) of the stripe's last row per column
  int64_t bndStride;
  int nslot;
  int* progress;
  int* finished;
  int* ticket;
  int* abort_flag;
  const uint2* ck;               // K1 checkpoints
  int2* ck32;                    // out: int32 checkpoints
  unsigned long long* hot32;     // out: [nTasks * kHotWords] bit c: a cell of column block c reached minscore (filed, like
                                 //      K1's mask, under the block the warp is in: the cell's own block or the one after)
  int* edgeFlag;                 // out: [nTasks] a cell >= minscore in the last column of a window that ends before the query
};

__global__ void __launch_bounds__(128) swg_score32_kernel(WideArgs a) {
  __shared__ int32_t smat[kMatStride * kMatStride];
  __shared__ int4 s_feed[4][32];  // per warp: (query code, H above, F above) for the 32 columns of the block
  __shared__ int2 s_out[4][32];   //           last row's (H, F) of the 32 columns lane 31 finished last
  load_score_table(smat, a.mat);
  __syncthreads();
  constexpr int R = kRW;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int NEGI = -(1 << 28);
  for (;;) {
    int t = 0;
    if (lane == 0) t = atomicAdd(a.ticket, 1);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= a.nTickets) break;
    const int ti = a.ticketTask[t];
    const WideTask tk = a.tasks[ti];
    const int k = t - tk.firstTicket;
    const int Q = a.qlen[tk.q], L = a.slen[tk.subject];
    const uint8_t* qc = (tk.strand ? a.codes_rc : a.codes_fwd) + a.qoff[tk.q];
    const uint8_t* sc = a.scodes + a.soff[tk.subject];
    const bool hasPrev = k > 0, hasNext = k + 1 < tk.nStripes;
    const int row0 = (k * 32 + lane) * R;
    int srow[R], H[R], E[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      srow[r] = ((row0 + r < L) ? (int)sc[row0 + r] : kPadCode) * kMatStride;
      H[r] = 0;
      E[r] = NEGI;
    }
    int diagH = 0;
    if (tk.ckOff >= 0) {
      const uint2* ck = a.ck + tk.ckOff;
      const int sh = tk.strand ? 16 : 0;
#pragma unroll
      for (int r = 0; r < R; r++)
        if (row0 + r < L) {
          const uint2 v = ck[row0 + r];
          H[r] = (int)((v.x >> sh) & 0xffffu);
          E[r] = (int)((v.y >> sh) & 0xffffu) + a.gap_init;
        }
      if (row0 > 0 && row0 - 1 < L) diagH = (int)((ck[row0 - 1].x >> sh) & 0xffffu);
    }
    const int W = tk.c1 - tk.c0;
    const int qBlocks = (W + 31) >> 5, nBlocks = qBlocks + 1;
    const int ckMask = (32 << tk.hotShift) - 1, ckShift = 5 + tk.hotShift;
    const int b0 = tk.c0 >> 5;
    const int2* bndIn = a.bnd + (int64_t)((t - 1 + a.nslot) % a.nslot) * a.bndStride;
    int2* bndOut = a.bnd + (int64_t)(t % a.nslot) * a.bndStride;
    if (hasNext) {
      // the reader (ticket + 1) of the last ticket that wrote this ring slot must be done
      for (int tOld = t - a.nslot; tOld >= 0; tOld -= a.nslot) {
        const WideTask& ot = a.tasks[a.ticketTask[tOld]];
        if (tOld - ot.firstTicket + 1 >= ot.nStripes) continue;  // that ticket had no stripe below: it wrote nothing
        if (lane == 0) wait_ge(&a.finished[tOld + 1], 1, a.abort_flag);
        __syncwarp();
        break;
      }
    }
    int outH = 0, outF = NEGI, curQ = kPadCode, blkMax = 0;
    bool edge = false;
    for (int b = 0; b < nBlocks; b++) {
      const int col = b * 32 + lane;
      int4 fv = make_int4((col < W) ? (int)qc[tk.c0 + col] : kPadCode, 0, NEGI, 0);
      if (hasPrev && b < qBlocks) {
        if (lane == 0 && ld_acquire(&a.progress[t - 1]) < b + 1) wait_ge(&a.progress[t - 1], min(b + kLagSweep, qBlocks), a.abort_flag);
        __syncwarp();
        const int2 raw = __ldcg(&bndIn[col]);
        fv.y = raw.x;
        fv.z = raw.y;
      }
      __syncwarp();  // everyone is done with the previous block's feed
      s_feed[wib][lane] = fv;
      __syncwarp();
      // The 32 columns of the block.  SLOW: the first block (lanes wait for their first column), a block that starts with
      // a checkpoint, and the last two (edge test); every other block runs without those per-column tests.
      auto columns = [&](auto slowTag) {
        constexpr bool SLOW = decltype(slowTag)::value;
#pragma unroll 4
        for (int j = 0; j < 32; j++) {
          const int4 fd = s_feed[wib][j];  // broadcast read; only lane 0 uses it
          int inQ = __shfl_up_sync(0xffffffffu, curQ, 1);
          int inH = __shfl_up_sync(0xffffffffu, outH, 1);
          int inF = __shfl_up_sync(0xffffffffu, outF, 1);
          inQ = lane == 0 ? fd.x : inQ;
          inH = lane == 0 ? fd.y : inH;
          inF = lane == 0 ? fd.z : inF;
          curQ = inQ;
          const int cw = b * 32 + j - lane;  // column inside the window, 0-based
          int hu = inH, f = inF;
          if (!SLOW || cw >= 0) {  // before its first column a lane keeps its (checkpointed) state untouched
            if (SLOW && cw > 0 && (cw & ckMask) == 0 && cw < W) {
              int2* dst = a.ck32 + tk.ck32Off + (int64_t)(cw >> ckShift) * L;
#pragma unroll
              for (int r = 0; r < R; r++)
                if (row0 + r < L) dst[row0 + r] = make_int2(H[r], E[r]);
            }
            int hd = diagH;
            diagH = inH;
            const int* scol = smat + curQ;
#pragma unroll
            for (int r = 0; r < R; r++) {
              const int hl = H[r];
              const int e = max(E[r] + a.ins_ext, hl + a.gap_init);
              const int fn = max(f + a.del_ext, hu + a.gap_init);
              const int d = hd + scol[srow[r]];
              const int hnew = max(max(e, fn), max(d, 0));
              E[r] = e;
              f = fn;
              hd = hl;
              H[r] = hnew;
              hu = hnew;
              blkMax = max(blkMax, hnew);
            }
            if (SLOW && cw == W - 1 && tk.c1 < Q) {
#pragma unroll
              for (int r = 0; r < R; r++) edge |= (row0 + r < L) && H[r] >= a.minscore;
            }
          }
          outH = hu;
          outF = f;
          if (hasNext) {
            const int c = b * 32 + j - 31;  // the column lane 31 has just finished
            if (lane == 31) s_out[wib][c & 31] = make_int2(outH, outF);
            if (c >= 0 && (c & 31) == 31) {
              __syncwarp();
              bndOut[(c & ~31) + lane] = s_out[wib][lane];
              __syncwarp();  // orders the lanes' stores before lane 0's cumulative release (see swg_origin_kernel)
              if (lane == 0) st_release(&a.progress[t], (c >> 5) + 1);
            }
          }
        }
      };
      if (b == 0 || ((b * 32) & ckMask) == 0 || b >= qBlocks - 1)
        columns(std::true_type{});
      else
        columns(std::false_type{});
      if (((((b0 + b + 1) >> tk.hotShift) != ((b0 + b) >> tk.hotShift)) || b == nBlocks - 1)) {
        if (blkMax >= a.minscore) {
          const int blk = min((b0 + b) >> tk.hotShift, kHotBlocks - 1);
          atomicOr(&a.hot32[(int64_t)ti * kHotWords + (blk >> 6)], 1ull << (blk & 63));
        }
        blkMax = 0;
      }
    }
    if (__any_sync(0xffffffffu, edge) && lane == 0) atomicExch(&a.edgeFlag[ti], 1);
    __syncwarp();
    if (lane == 0) st_release(&a.finished[t], 1);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// K2b: block-parallel origin pass (the common path).  A task = one (query, strand, subject) triple x one hot column
// block [c0, c1) (CB columns, CB = the K1 checkpoint spacing).  Every task starts from K1's exact (H, E') column
// checkpoint, so all blocks of a triple run in parallel; the subject's 256-row stripes of a task are tickets that
// different warps pipeline through a boundary ring with progress flags (same protocol as swg_origin_kernel; the
// host orders the tickets so that long tasks start first).  Paths that are alive at the left edge carry a PSEUDO origin naming the boundary cell
// (row, H-or-E state); each task snapshots the origins of its last column, and resolve_candidates_kernel follows
// pseudo origins through those snapshots back to the real origin cell.  Origins are encoded relative to the task,
// (i0 - c0) * (L + 1) + j0, which must stay below 2^31 (bit 31 marks pseudo origins).
// ---------------------------------------------------------------------------------------------------------------
constexpr uint32_t kPseudoBit = 0x80000000u;
constexpr uint32_t kPseudoEBit = 1u << 20;  // pseudo origin of the E state (else H) of boundary row (bits 0..19, 1-based)

struct BlockTask {
  int32_t q, strand, subject, triple;
  int32_t c0, c1;     // query columns [c0, c1), 0-based
  int32_t prevTask;   // task of the same triple covering [.., c0), or -1
  int32_t firstTicket, nStripes;  // one ticket per 256-row stripe; the stripes of a task form a pipeline
  int32_t pad_;
  int64_t ckOff;      // subject's first row in the K1 checkpoint taken before column c0, or -1 (c0 == 0)
  int64_t ck32Off;    // >= 0: start from this int32 checkpoint of swg_score32_kernel instead (ck32[ck32Off + row] = (H, E))
  int64_t snapOff;    // this task's origin snapshot: snap[snapOff + row] = oH, snap[snapOff + L + row] = oE (0-based row)
};

struct BlockArgs {
  const BlockTask* tasks;
  const int32_t* ticketTask;   // [nTickets] task of the ticket; tickets go stripe level by stripe level over all tasks
  const int32_t* ticketStripe; // [nTickets] stripe of the ticket
  const int32_t* prodTicket;   // [nTickets] ticket of (task, stripe-1) or -1: always an older ticket
  const int32_t* consTicket;   // [nTickets] ticket of (task, stripe+1) or -1
  int nTickets;
  const uint8_t* codes_fwd;
  const uint8_t* codes_rc;
  const int64_t* qoff;
  const uint8_t* scodes;
  const int64_t* soff;
  const int32_t* slen;
  int32_t mat[16][16];
  int32_t gap_init, ins_ext, del_ext, minscore;
  const uint2* ck;
  const int2* ck32;            // int32 checkpoints of the triples whose K1 values wrapped (swg_score32_kernel)
  OriginBoundary* bnd;         // [nslot * bndStride] stripe boundary ring (same protocol as swg_origin_kernel)
  int64_t bndStride;
  int nslot;
  int* progress;
  int* finished;
  int* abort_flag;
  uint32_t* snap;
  int* ticket;
  unsigned long long* keys;    // (task << 32 | task-relative origin) -> best end
  unsigned long long* vals;
  uint32_t tableMask;
  int* overflow;
  unsigned long long* pbest;   // best end per PSEUDO origin, dense, laid out like snap: [snapOff + (E ? L : 0) + row]
  unsigned long long* dbg;     // optional [8]: producer-wait cycles, slot-wait cycles, tickets, ticket cycles, sum and max of
                               // (warp exit - first warp start) in ns, -, time base
};

__device__ __forceinline__ void table_emit(unsigned long long* keys, unsigned long long* vals, uint32_t mask, int* overflow,
                                           unsigned long long key, unsigned long long val) {
  uint32_t slot = (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 32) & mask;
  for (uint32_t probe = 0; probe < 128u && probe <= mask; probe++) {  // a long probe sequence = table too full: grow it
    unsigned long long old = atomicCAS(&keys[slot], kEmptyKey, key);
    if (old == kEmptyKey || old == key) {
      atomicMax(&vals[slot], val);
      return;
    }
    slot = (slot + 1) & mask;
  }
  atomicExch(overflow, 1);
}

__global__ void __launch_bounds__(128, 4) swg_origin_block_kernel(BlockArgs a) {
  __shared__ int32_t smat[kMatStride * kMatStride];
  __shared__ int4 s_feed[4][32];   // per warp: (H, F, origin of H, origin of F) of the row above, one entry per column of the block
  __shared__ int s_feedq[4][32];   //           query residue of the column
  __shared__ int4 s_out[4][32];    //           last row of this stripe, collected column by column for the stripe below
  load_score_table(smat, a.mat);
  __syncthreads();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  constexpr int R = kR2;
  const int NEGI = -(1 << 28);
  if (a.dbg && lane == 0) atomicCAS(&a.dbg[7], 0ull, globaltimer_ns());
  for (;;) {
    int t = 0;
    if (lane == 0) t = atomicAdd(a.ticket, 1);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= a.nTickets) break;
    const long long tk0 = clock64();
    long long waitCyc = 0;
    const int ti = a.ticketTask[t];
    const BlockTask tk = a.tasks[ti];
    const int st = a.ticketStripe[t];  // stripe of this ticket
    const int L = a.slen[tk.subject];
    const uint32_t Lp1 = (uint32_t)L + 1u;
    const uint8_t* qc = (tk.strand ? a.codes_rc : a.codes_fwd) + a.qoff[tk.q];
    const uint8_t* sc = a.scodes + a.soff[tk.subject];
    const int W = tk.c1 - tk.c0;
    const int qBlocks = (W + 31) >> 5, nBlocks = qBlocks + 1;
    const int sh = tk.strand ? 16 : 0;
    const int tProd = a.prodTicket[t];
    const bool hasPrev = tProd >= 0, hasNext = a.consTicket[t] >= 0;
    const OriginBoundary* bndIn = a.bnd + (int64_t)((hasPrev ? tProd : 0) % a.nslot) * a.bndStride;
    OriginBoundary* bndOut = a.bnd + (int64_t)(t % a.nslot) * a.bndStride;
    if (hasNext) {
      // the reader of the last ticket that wrote this ring slot must be done
      for (int tOld = t - a.nslot; tOld >= 0; tOld -= a.nslot) {
        const int cOld = a.consTicket[tOld];
        if (cOld < 0) continue;  // that ticket had no stripe below: it wrote nothing
        const long long g0 = clock64();
        if (lane == 0) wait_ge(&a.finished[cOld], 1, a.abort_flag);
        __syncwarp();
        if (a.dbg && lane == 0) atomicAdd(&a.dbg[1], (unsigned long long)(clock64() - g0));
        break;
      }
    }
    uint32_t lastOrigin = 0xFFFFFFFFu;  // nothing cached
    unsigned long long lastVal = 0;
    const int row0 = (st * 32 + lane) * R;
    int srow[R], H[R], E[R];
    uint32_t oH[R], oE[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      srow[r] = ((row0 + r < L) ? (int)sc[row0 + r] : kPadCode) * kMatStride;
      H[r] = 0;
      E[r] = NEGI;
      oH[r] = 0;
      oE[r] = 0;
    }
    int diagH = 0;
    uint32_t diagO = 0;
    if (tk.ck32Off >= 0) {
      const int2* ck = a.ck32 + tk.ck32Off;
#pragma unroll
      for (int r = 0; r < R; r++)
        if (row0 + r < L) {
          const int2 v = ck[row0 + r];
          H[r] = v.x;
          E[r] = v.y;
          oH[r] = kPseudoBit | (uint32_t)(row0 + r + 1);
          oE[r] = kPseudoBit | kPseudoEBit | (uint32_t)(row0 + r + 1);
        }
      if (row0 > 0 && row0 - 1 < L) {
        diagH = ck[row0 - 1].x;
        diagO = kPseudoBit | (uint32_t)row0;
      }
    } else if (tk.ckOff >= 0) {
      const uint2* ck = a.ck + tk.ckOff;
#pragma unroll
      for (int r = 0; r < R; r++)
        if (row0 + r < L) {
          const uint2 v = ck[row0 + r];
          H[r] = (int)((v.x >> sh) & 0xffffu);
          E[r] = (int)((v.y >> sh) & 0xffffu) + a.gap_init;
          oH[r] = kPseudoBit | (uint32_t)(row0 + r + 1);
          oE[r] = kPseudoBit | kPseudoEBit | (uint32_t)(row0 + r + 1);
        }
      if (row0 > 0 && row0 - 1 < L) {
        diagH = (int)((ck[row0 - 1].x >> sh) & 0xffffu);
        diagO = kPseudoBit | (uint32_t)row0;
      }
    }
    int outH = 0, outF = NEGI, curQ = kPadCode;
    uint32_t outOH = 0, outOF = 0;
    for (int b = 0; b < nBlocks; b++) {
      const int col = b * 32 + lane;
      int4 fv = make_int4(0, NEGI, 0, 0);
      if (hasPrev && b < qBlocks) {
        const long long w0 = clock64();
        if (lane == 0 && ld_acquire(&a.progress[tProd]) < b + 1) wait_ge(&a.progress[tProd], b + 1, a.abort_flag);
        __syncwarp();
        waitCyc += clock64() - w0;
        fv = __ldcg(reinterpret_cast<const int4*>(&bndIn[col]));
      }
      __syncwarp();  // everyone is done with the previous block's feed
      s_feed[wib][lane] = fv;
      s_feedq[wib][lane] = (col < W) ? (int)qc[tk.c0 + col] : kPadCode;
      __syncwarp();
      for (int j = 0; j < 32; j++) {
        // values moving one lane down per step; lane 0 takes the block's feed (one broadcast read of shared memory)
        int inQ = __shfl_up_sync(0xffffffffu, curQ, 1);
        int inH = __shfl_up_sync(0xffffffffu, outH, 1);
        int inF = __shfl_up_sync(0xffffffffu, outF, 1);
        uint32_t inOH = __shfl_up_sync(0xffffffffu, outOH, 1);
        uint32_t inOF = __shfl_up_sync(0xffffffffu, outOF, 1);
        if (lane == 0) {
          const int4 fd = s_feed[wib][j];
          inQ = s_feedq[wib][j];
          inH = fd.x;
          inF = fd.y;
          inOH = (uint32_t)fd.z;
          inOF = (uint32_t)fd.w;
        }
        curQ = inQ;
        const int cw = b * 32 + j - lane;  // column inside the block, 0-based
        int hd = diagH, hu = inH, f = inF;
        uint32_t od = diagO, ou = inOH, of = inOF;
        if (cw >= 0) {  // before its first column a lane must keep the checkpointed state untouched
          diagH = inH;
          diagO = inOH;
          // the R cells of the column are straight-line code; the rare candidate ends are only marked and handled after
          const uint32_t freshBase = (uint32_t)(cw + 1) * Lp1 + (uint32_t)(row0 + 1);
          const int* scol = smat + curQ;
          uint32_t ends = 0;
#pragma unroll
          for (int r = 0; r < R; r++) {
            const int hl = H[r];
            const uint32_t ol = oH[r];
            const OriginCell c = origin_cell(hl, ol, hu, ou, hd, od, scol[srow[r]], freshBase + (uint32_t)r, E[r], oE[r], f, of,
                                             a.gap_init, a.ins_ext, a.del_ext);
            if (c.viaDiag && c.h >= a.minscore) ends |= 1u << r;
            hd = hl;
            od = ol;
            H[r] = c.h;
            oH[r] = c.o;
            hu = c.h;
            ou = c.o;
          }
          if (ends) {
            const int i1 = tk.c0 + cw + 1;  // 1-based query column
#pragma unroll
            for (int r = 0; r < R; r++)
              if (ends & (1u << r)) {
                const unsigned long long val = cand_val(H[r], i1, row0 + r + 1);
                const uint32_t o = oH[r];
                if (o & kPseudoBit) {
                  // a path that entered through the left edge: one dense slot per boundary cell and a reduction that
                  // returns nothing.  (To the right of a high-scoring alignment its decaying fan crosses the edge in
                  // thousands of rows, each with its own pseudo origin: through the hash table every such end was a
                  // compare-and-swap round trip to L2, and those tickets ran 15x longer than the others.)
                  atomicMax(&a.pbest[tk.snapOff + ((o & kPseudoEBit) ? L : 0) + (int64_t)((o & 0xFFFFFu) - 1u)], val);
                } else if (o == lastOrigin) {
                  lastVal = val > lastVal ? val : lastVal;
                } else {
                  if (lastOrigin != 0xFFFFFFFFu)
                    table_emit(a.keys, a.vals, a.tableMask, a.overflow, ((unsigned long long)(uint32_t)ti << 32) | lastOrigin, lastVal);
                  lastOrigin = o;
                  lastVal = val;
                }
              }
          }
          if (cw == W - 1) {  // last column of the block: snapshot the origins of (H, E) for the next block's tasks
            uint32_t* sn = a.snap + tk.snapOff;
#pragma unroll
            for (int r = 0; r < R; r++)
              if (row0 + r < L) {
                sn[row0 + r] = oH[r];
                sn[L + row0 + r] = oE[r];
              }
          }
        }
        outH = hu;
        outF = f;
        outOH = ou;
        outOF = of;
        if (hasNext) {
          // lane 31 has just finished column c = b*32 + j - 31 of the stripe's last row
          const int c = b * 32 + j - 31;
          if (lane == 31) s_out[wib][c & 31] = make_int4(outH, outF, (int)outOH, (int)outOF);
          if (c >= 0 && (c & 31) == 31) {  // a block of 32 columns is complete: one coalesced store, then publish it
            __syncwarp();
            *reinterpret_cast<int4*>(&bndOut[(c & ~31) + lane]) = s_out[wib][lane];
            __syncwarp();  // (orders the lanes' stores before lane 0's cumulative release, see swg_origin_kernel)
            if (lane == 0) st_release(&a.progress[t], (c >> 5) + 1);
          }
        }
      }
    }
    if (lastOrigin != 0xFFFFFFFFu)
      table_emit(a.keys, a.vals, a.tableMask, a.overflow, ((unsigned long long)(uint32_t)ti << 32) | lastOrigin, lastVal);
    __syncwarp();
    if (lane == 0) {
      st_release(&a.finished[t], 1);
      if (a.dbg) {
        atomicAdd(&a.dbg[0], (unsigned long long)waitCyc);
        atomicAdd(&a.dbg[2], 1ull);
        atomicAdd(&a.dbg[3], (unsigned long long)(clock64() - tk0));
        // the slowest ticket: duration and producer wait in units of 1024 cycles, and which ticket it was
        atomicMax(&a.dbg[6], ((unsigned long long)((clock64() - tk0) >> 10) << 44) | ((unsigned long long)((waitCyc >> 10) & 0xFFFFF) << 24) |
                                 (unsigned long long)(t & 0xFFFFFF));
        // per stripe level (first 64): when its last ticket ended, relative to the first warp's start
        if (st < 64) atomicMax(&a.dbg[8 + st], globaltimer_ns() - *(volatile unsigned long long*)&a.dbg[7]);
      }
    }
  }
  if (a.dbg && lane == 0) {
    const unsigned long long dt = globaltimer_ns() - *(volatile unsigned long long*)&a.dbg[7];
    atomicAdd(&a.dbg[4], dt);
    atomicMax(&a.dbg[5], dt);
  }
}

// Follows pseudo origins through the per-task snapshots to the real origin cell and merges every candidate into the
// final table keyed by (triple, i0, j0).  A pseudo origin that leads into a column block that was not swept marks
// the triple for a redo with all blocks.
struct ResolveArgs {
  const BlockTask* tasks;
  int nTasks;
  const int32_t* slen;
  const uint32_t* snap;
  const unsigned long long* keys;
  const unsigned long long* vals;
  uint32_t tableSize;
  const unsigned long long* pbest;  // dense bests of the pseudo origins (BlockArgs::pbest)
  unsigned long long* fkeys;   // (triple << 40 | i0 << 20 | j0)
  unsigned long long* fvals;
  uint32_t fmask;
  int* overflow;
  int* redo;                   // [nTriples]
  int verbose;
};

__device__ __forceinline__ void resolve_one(const ResolveArgs& a, const int t0, const uint32_t o0, const unsigned long long val) {
  int t = t0;
  uint32_t o = o0;
  const int triple = a.tasks[t].triple;
  const int L = a.slen[a.tasks[t].subject];
  bool ok = true;
  for (int guard = 0; (o & kPseudoBit) && guard < 128; guard++) {
    const int pt = a.tasks[t].prevTask;
    if (pt < 0) {
      ok = false;
      break;
    }
    const uint32_t row = o & 0xFFFFFu;  // 1-based
    const bool isE = o & kPseudoEBit;
    o = a.snap[a.tasks[pt].snapOff + (isE ? L : 0) + (row - 1)];
    t = pt;
    if (o == 0) {
      ok = false;
      break;
    }
  }
  if (!ok || (o & kPseudoBit)) {
    if (atomicExch(&a.redo[triple], 1) == 0 && a.verbose)
      printf("[resolve] triple %d: task %d (c0 %d) origin %x -> stuck at task %d (c0 %d prev %d) o %x val %llx\n", triple, t0, a.tasks[t0].c0, o0, t,
             a.tasks[t].c0, a.tasks[t].prevTask, o, val);
    return;
  }
  const uint32_t Lp1 = (uint32_t)L + 1u;
  const unsigned long long i0 = (unsigned long long)(o / Lp1) + (unsigned long long)a.tasks[t].c0;
  const unsigned long long j0 = o % Lp1;
  table_emit(a.fkeys, a.fvals, a.fmask, a.overflow, ((unsigned long long)(uint32_t)triple << 40) | (i0 << 20) | j0, val);
}

__global__ void resolve_candidates_kernel(ResolveArgs a) {
  // ends whose origin lies inside their own task (hash table) ...
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < a.tableSize; s += gridDim.x * blockDim.x) {
    const unsigned long long key = a.keys[s];
    if (key == kEmptyKey) continue;
    resolve_one(a, (int)(key >> 32), (uint32_t)key, a.vals[s]);
  }
  // ... and ends of paths that entered their task through its left edge (dense, one slot per boundary cell)
  for (int t = blockIdx.x; t < a.nTasks; t += gridDim.x) {
    const int L = a.slen[a.tasks[t].subject];
    const unsigned long long* pb = a.pbest + a.tasks[t].snapOff;
    for (int e = threadIdx.x; e < 2 * L; e += blockDim.x) {
      const unsigned long long val = pb[e];
      if (val == 0) continue;
      const uint32_t o = kPseudoBit | (e >= L ? kPseudoEBit : 0u) | (uint32_t)((e >= L ? e - L : e) + 1);
      resolve_one(a, t, o, val);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// K3: banded global traceback, one thread per candidate (latency-bound; a few % of a stage).
// Local coordinates u=1..n (query), v=1..m (subject); diagonals d=u-v in [dlo, dhi], dlo=min(0,n-m)-W,
// dhi=max(0,n-m)+W.  The alignment must open with a diagonal at (1,1) and close with one at (n,m).
// ops (written back to front): 0 = aligned pair, 1 = gap in subject row (query base), 2 = gap in query row.
// status: 1 ok (banded score == raw), 0 band too narrow -> the host doubles W and retries that candidate.
// ---------------------------------------------------------------------------------------------------------------
struct TraceCand {
  int32_t q, strand, subject;
  int32_t i0, j0, i1, j1;   // 1-based, strand coordinates
  int32_t raw;
  int32_t W;                // band half-width of the first attempt
  int32_t attempts;         // attempts made inside the kernel, W doubling each time (scratch is sized for the last)
  int32_t id;               // candidate number of the call: index into opsLen / status (a launch holds a subset)
  int32_t pad_;
  int64_t dirOff;           // offset into dir scratch
  int64_t rowOff;           // offset into row scratch (3 * bw ints)
  int64_t opsOff;           // offset into ops
};

struct TraceArgs {
  const TraceCand* cand;
  int nCand;
  const uint8_t* codes_fwd;
  const uint8_t* codes_rc;
  const int64_t* qoff;
  const uint8_t* scodes;
  const int64_t* soff;
  int32_t mat[16][16];
  int32_t gap_init, ins_ext, del_ext;
  uint8_t* dir;
  int32_t* rows;
  uint8_t* ops;
  int32_t* opsLen;   // [candidate id] number of ops, or -1
  int32_t* status;   // [candidate id] 1 + index of the attempt that succeeded, 0 if none did
};

// Band cell (u, k) (k = d - dlo, v = u - dlo - k) depends on (u-1, k-1) [E], (u-1, k) [D] and (u, k+1) [F]; all
// cells with the same w = 2u - k are independent, so a team of T threads (one warp for narrow bands, up to 512
// threads of a block for wide ones: one thread per cell of a wave) sweeps waves w, thread l owning k = 2l (+parity),
// 2l + 2T, ...  H/E/F live in place in shared memory (global scratch when the band is wider than the shared
// buffer), even and odd k in separate halves so that a wave's accesses are unit-stride; entries of parity w hold
// wave w-2 (the diagonal source), the others wave w-1.  The walk back is done by lane 0 of the team's first warp on
// a 32 x 32 window of direction bytes that the warp refills with 32 independent loads, instead of one dependent
// global load per step.
constexpr int kTraceWarpBand = 256;    // ints per array per warp in warp mode
constexpr int kTraceWarps = 4;         // warps (= candidates) per block in warp mode
constexpr int kTraceBlockThreads = 512;  // block mode: threads per candidate (the team is sized to the band)
constexpr int kTraceBlockBand = 2048;  // ints per array in block mode
constexpr int kTraceSeqBlock = 10240;  // query / subject bytes of a candidate staged in shared memory (block mode)
constexpr int kTraceSeqWarp = 1024;    // the same per warp in warp mode

// The wave loop of traceback_kernel when the band fits the shared buffer.  One cell of each parity per thread and step
// (J such pairs when the band is wider than the team): pair l owns k = 2l (wave 2s) and k = 2l + 1 (wave 2s + 1), which share their subject column
// v = s - l - dlo and whose query rows are u = s + l and s + l + 1.  Even k live in slots [0, half), slot `half` is
// a guard, odd k in [half + 1, half + 1 + bw/2), then a second guard: the neighbours outside the band read NEGI from
// the guards (never written), so a wave is 5 shared loads, the recurrence, 3 stores and one direction byte.  The
// substitution scores of step s + 1 are looked up during step s (they do not depend on the band), off the chain
// load -> recurrence -> store -> barrier that bounds a candidate.
template <int J, class Sync>
__device__ __forceinline__ void paired_waves(int32_t* Hb, int32_t* Eb, int32_t* Fb, const uint8_t* qc, const uint8_t* sc,
                                             const int32_t* smat, uint8_t* dir, int n, int m, int dlo, int bw, int tid, int T,
                                             int gi, int ie, int de, Sync team_sync) {
  constexpr int NEGI = -(1 << 28);
  const int half = (bw + 1) >> 1;
  const int sLo = (3 - bw) >> 1;  // floor: wave 2 sLo may be one before the first useful one
  auto qat = [&](int u) { return (int)qc[min(max(u, 1), n) - 1]; };       // residue of query row u (clamped)
  auto sat = [&](int v) { return (int)sc[min(max(v, 1), m) - 1] * 16; };  // score row of subject column v
  // pair j of this thread: slot l = tid + j T, cells k = 2l and 2l + 1
  int qo[J], se[J], so[J];
#pragma unroll
  for (int j = 0; j < J; j++) {
    const int l = tid + j * T, ue = sLo + l, v = sLo - l - dlo;
    qo[j] = qat(ue + 1);
    se[j] = smat[sat(v) + qat(ue)];
    so[j] = smat[sat(v) + qo[j]];
  }
  int64_t dE = (int64_t)(sLo + tid) * bw + 2 * tid;  // direction byte of pair 0's even cell
  const int64_t dJ = (int64_t)T * (bw + 2);          // from one pair of the thread to the next
  for (int s = sLo; s <= n; s++, dE += bw) {
    int sen[J], qn[J], srn[J];
#pragma unroll
    for (int j = 0; j < J; j++) {
      const int l = tid + j * T, ue = s + l, v = s - l - dlo;
      // next step: rows ue + 1 (even; the odd cell's residue now) and ue + 2, column v + 1
      qn[j] = qat(ue + 2);
      srn[j] = sat(v + 1);
      if (2 * l < bw && (unsigned)(ue - 1) < (unsigned)n) {
        int32_t *He = Hb + l, *Ee = Eb + l, *Fe = Fb + l;
        const int32_t *Ho = Hb + half + 1 + l, *Eo = Eb + half + 1 + l, *Fo = Fb + half + 1 + l;
        int hnew = NEGI, enew = NEGI, fnew = NEGI;
        if ((unsigned)(v - 1) < (unsigned)m) {
          const int el = Eo[-1], hl = Ho[-1], fr = Fo[0], hr = Ho[0], hd = He[0];
          const int e_ext = el + ie, e_open = hl + gi, f_ext = fr + de, f_open = hr + gi;
          uint32_t bits = (e_ext >= e_open ? 4u : 0u) | (f_ext >= f_open ? 8u : 0u);
          enew = max(max(e_ext, e_open), NEGI);
          fnew = max(max(f_ext, f_open), NEGI);
          // (1,1) needs no special case here: its E and F sources are row 0 / column 0, NEGI, and the clamp above
          // gives exactly the NEGI the generic loop below assigns
          const int D = hd + se[j];
          hnew = max(max(enew, fnew), max(D, NEGI));
          bits |= D > max(enew, fnew) ? 3u : (fnew > enew ? 2u : 1u);
          dir[dE + j * dJ] = (uint8_t)bits;
        }
        He[0] = hnew;
        Ee[0] = enew;
        Fe[0] = fnew;
      }
      sen[j] = smat[srn[j] + qo[j]];
    }
    team_sync();
#pragma unroll
    for (int j = 0; j < J; j++) {
      const int l = tid + j * T, ue = s + l, v = s - l - dlo;
      if (2 * l + 1 < bw && (unsigned)ue < (unsigned)n) {  // u = ue + 1
        const int32_t *He = Hb + l, *Ee = Eb + l, *Fe = Fb + l;
        int32_t *Ho = Hb + half + 1 + l, *Eo = Eb + half + 1 + l, *Fo = Fb + half + 1 + l;
        int hnew = NEGI, enew = NEGI, fnew = NEGI;
        if ((unsigned)(v - 1) < (unsigned)m) {
          const int el = Ee[0], hl = He[0], fr = Fe[1], hr = He[1], hd = Ho[0];
          const int e_ext = el + ie, e_open = hl + gi, f_ext = fr + de, f_open = hr + gi;
          uint32_t bits = (e_ext >= e_open ? 4u : 0u) | (f_ext >= f_open ? 8u : 0u);
          enew = max(max(e_ext, e_open), NEGI);
          fnew = max(max(f_ext, f_open), NEGI);
          const int D = hd + so[j];
          hnew = max(max(enew, fnew), max(D, NEGI));
          bits |= D > max(enew, fnew) ? 3u : (fnew > enew ? 2u : 1u);
          dir[dE + j * dJ + bw + 1] = (uint8_t)bits;
        }
        Ho[0] = hnew;
        Eo[0] = enew;
        Fo[0] = fnew;
      }
      so[j] = smat[srn[j] + qn[j]];
      se[j] = sen[j];
      qo[j] = qn[j];
    }
    team_sync();
  }
}

template <bool BLOCK>
__global__ void __launch_bounds__(BLOCK ? kTraceBlockThreads : kTraceWarps * 32) traceback_kernel(TraceArgs a) {
  constexpr int SB = BLOCK ? kTraceBlockBand : kTraceWarpBand;
  constexpr int SQ = BLOCK ? kTraceSeqBlock : kTraceSeqWarp;
  constexpr int NT = BLOCK ? 1 : kTraceWarps;  // candidates per block
  __shared__ int32_t smat[16 * 16];
  __shared__ int32_t sband[NT][3][SB];
  __shared__ uint8_t sseq[NT][2][SQ];
  __shared__ uint8_t swin[NT][32][32];
  __shared__ int s_len[NT];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) smat[i] = a.mat[i >> 4][i & 15];
  __syncthreads();
  const int tid = BLOCK ? threadIdx.x : (threadIdx.x & 31);
  const int lane = threadIdx.x & 31;
  const int wib = BLOCK ? 0 : (threadIdx.x >> 5);
  const int c = BLOCK ? blockIdx.x : blockIdx.x * kTraceWarps + wib;
  if (c >= a.nCand) return;  // whole block (block mode) or whole warp (warp mode)
  const int TALL = BLOCK ? kTraceBlockThreads : 32;
  const TraceCand tc = a.cand[c];
  const int NEGI = -(1 << 28);
  const int n = tc.i1 - tc.i0 + 1, m = tc.j1 - tc.j0 + 1;
  const uint8_t* qc = (tc.strand ? a.codes_rc : a.codes_fwd) + a.qoff[tc.q] + (tc.i0 - 1);
  const uint8_t* sc = a.scodes + a.soff[tc.subject] + (tc.j0 - 1);
  if (n <= SQ && m <= SQ) {  // stage both sequences: the waves then never wait on global memory for their residues
    for (int i = tid; i < n; i += TALL) sseq[wib][0][i] = qc[i];
    for (int i = tid; i < m; i += TALL) sseq[wib][1][i] = sc[i];
    qc = sseq[wib][0];
    sc = sseq[wib][1];
  }
  // every thread of the block (warp) staged residues, but a narrow band's team is only its first T threads and
  // team_sync() only covers those: the staged bytes must be visible to the team before its first wave
  if (BLOCK)
    __syncthreads();
  else
    __syncwarp();
  uint8_t* dir = a.dir + tc.dirOff;
  int len = -1, used = 0;
  for (int attempt = 0; attempt < tc.attempts && len <= 0; attempt++) {
    const int W = tc.W << attempt;
    const int dlo = min(0, n - m) - W, dhi = max(0, n - m) + W;
    const int bw = dhi - dlo + 1;
    const int half = (bw + 1) >> 1;
    // paired waves (paired_waves above) when the band plus two guard slots fits the shared buffer: then every thread has
    // pairsPer (1, 2 or 4) cells of each parity; values further than 2^27 from NEGI cannot occur below 2^20 rows
    const bool paired = bw + 2 <= SB && n < (1 << 20);
    const int pairsPer = !paired ? 1 : BLOCK ? (half + kTraceBlockThreads - 1) / kTraceBlockThreads : (half <= 32 ? 1 : half <= 64 ? 2 : 4);
    // team: one thread per cell (pairsPer cells) of a wave, whole warps
    const int T = BLOCK ? min(kTraceBlockThreads, max(32, ((half + pairsPer - 1) / pairsPer + 31) & ~31)) : 32;
    auto team_sync = [&]() {
      if (BLOCK)
        asm volatile("bar.sync 1, %0;" ::"r"(T) : "memory");
      else
        __syncwarp();
    };
    auto ix = [&](int k) { return (k >> 1) + ((k & 1) ? half + (paired ? 1 : 0) : 0); };
    int32_t *Hb, *Eb, *Fb;
    if (bw <= SB) {
      Hb = sband[wib][0];
      Eb = sband[wib][1];
      Fb = sband[wib][2];
    } else {
      Hb = a.rows + tc.rowOff;
      Eb = Hb + bw;
      Fb = Eb + bw;
    }
    if (tid < T) {
      // row u = 0: only (0,0) (d = 0) is reachable
      for (int k = tid; k < bw; k += T) {
        Hb[ix(k)] = (k == -dlo) ? 0 : NEGI;
        Eb[ix(k)] = NEGI;
        Fb[ix(k)] = NEGI;
      }
      if (paired && tid == 0) {  // the two guard slots: never written, read as the neighbour outside the band
        const int g0 = half, g1 = half + 1 + (bw >> 1);
        Hb[g0] = Eb[g0] = Fb[g0] = NEGI;
        Hb[g1] = Eb[g1] = Fb[g1] = NEGI;
      }
      team_sync();
      if (paired) {
        // the band is in shared memory here; calling with the arrays themselves (and with the staged residues when
        // they are) lets the compiler emit shared-memory loads instead of generic ones
        const bool staged = qc == sseq[wib][0];
#define RMGPU_PAIRED(JJ)                                                                                                   \
  if (staged)                                                                                                              \
    paired_waves<JJ>(sband[wib][0], sband[wib][1], sband[wib][2], sseq[wib][0], sseq[wib][1], smat, dir, n, m, dlo, bw, tid, \
                     T, a.gap_init, a.ins_ext, a.del_ext, team_sync);                                                      \
  else                                                                                                                     \
    paired_waves<JJ>(sband[wib][0], sband[wib][1], sband[wib][2], qc, sc, smat, dir, n, m, dlo, bw, tid, T, a.gap_init,      \
                     a.ins_ext, a.del_ext, team_sync);
        if (pairsPer == 1) {
          RMGPU_PAIRED(1)
        } else if (pairsPer == 2) {
          RMGPU_PAIRED(2)
        } else {
          RMGPU_PAIRED(4)
        }
#undef RMGPU_PAIRED
      } else
      // waves: u >= 1, 0 <= k < bw  ->  w = 2u - k in [3 - bw, 2n]
      for (int w = 3 - bw; w <= 2 * n; w++) {
        // only k with 1 <= (w + k) / 2 <= n matter: k in [2 - w, 2n - w]
        const int klo = max(0, 2 - w), khi = min(bw - 1, 2 * n - w);
        for (int k0 = klo & ~1; k0 <= khi; k0 += 2 * T) {
          const int k = k0 + 2 * tid + (w & 1);
          const int u = (w + k) >> 1;
          const int v = u - dlo - k;
          const bool inBand = k < bw && u >= 1 && u <= n;
          int hnew = NEGI, enew = NEGI, fnew = NEGI;
          uint8_t bits = 0;
          if (inBand && v >= 1 && v <= m) {
            const int il = ix(k - 1 < 0 ? 0 : k - 1), ir = ix(k + 1 < bw ? k + 1 : k);
            const int e_src_e = k > 0 ? Eb[il] : NEGI, e_src_h = k > 0 ? Hb[il] : NEGI;
            const int f_src_f = k + 1 < bw ? Fb[ir] : NEGI, f_src_h = k + 1 < bw ? Hb[ir] : NEGI;
            const int hd = Hb[ix(k)];
            int e_ext = e_src_e + a.ins_ext, e_open = e_src_h + a.gap_init;
            if (e_ext >= e_open) {
              enew = e_ext;
              bits |= 4;
            } else
              enew = e_open;
            int f_ext = f_src_f + a.del_ext, f_open = f_src_h + a.gap_init;
            if (f_ext >= f_open) {
              fnew = f_ext;
              bits |= 8;
            } else
              fnew = f_open;
            int D = hd > NEGI / 2 ? hd + smat[(int)sc[v - 1] * 16 + (int)qc[u - 1]] : NEGI;
            if (u == 1 && v == 1) {
              enew = NEGI;
              fnew = NEGI;
            }
            hnew = enew;
            uint8_t ch = 1;
            if (fnew > hnew) {
              hnew = fnew;
              ch = 2;
            }
            if (D > hnew) {
              hnew = D;
              ch = 3;
            }
            if (hnew < NEGI / 2) hnew = NEGI;
            if (enew < NEGI / 2) enew = NEGI;
            if (fnew < NEGI / 2) fnew = NEGI;
            dir[(int64_t)u * bw + k] = bits | ch;
          }
          if (inBand) {
            const int ik = ix(k);
            Hb[ik] = hnew;
            Eb[ik] = enew;
            Fb[ik] = fnew;
          }
        }
        team_sync();
      }
    }
    // after the last wave H(n, k) sits in Hb[k] for every k (each k's last in-range wave is row n)
    if (tid < 32) {
      const int kend = (n - m) - dlo;
      // the direction bytes were written by other threads of the team: make them visible to this warp's loads
      __threadfence_block();
      int uu = n, vv = m, state = 0, pos = 0;
      bool ok = Hb[ix(kend)] == tc.raw && (dir[(int64_t)n * bw + kend] & 3) == 3, done = false;
      uint8_t* ops = a.ops + tc.opsOff;
      for (;;) {
        const int uu0 = __shfl_sync(0xffffffffu, uu, 0), vv0 = __shfl_sync(0xffffffffu, vv, 0);
        if (__shfl_sync(0xffffffffu, (int)(done || !ok), 0)) break;
        const int kk0 = (uu0 - vv0) - dlo;
        const int kbase = max(0, min(kk0 - 16, bw - 32));
#pragma unroll 8
        for (int r = 0; r < 32; r++) {
          const int row = uu0 - r, col = kbase + lane;
          swin[wib][r][lane] = (row >= 1 && col < bw) ? dir[(int64_t)row * bw + col] : (uint8_t)0;
        }
        __syncwarp();
        if (lane == 0) {
          for (;;) {
            if (uu == 0 && vv == 0) {
              done = true;
              break;
            }
            if (uu < 1 || vv < 1) {
              ok = false;
              break;
            }
            const int kk = (uu - vv) - dlo;
            if (kk < 0 || kk >= bw) {
              ok = false;
              break;
            }
            const int r = uu0 - uu, cc = kk - kbase;
            if (r > 31 || cc < 0 || cc > 31) break;  // left the window: refill around the current cell
            const uint8_t bch = swin[wib][r][cc];
            if (state == 0) {
              const int ch = bch & 3;
              if (ch == 3) {
                ops[pos++] = 0;
                uu--;
                vv--;
              } else if (ch == 1)
                state = 1;
              else if (ch == 2)
                state = 2;
              else {
                ok = false;  // a cell the sweep never wrote
                break;
              }
            } else if (state == 1) {
              ops[pos++] = 1;
              state = (bch & 4) ? 1 : 0;
              uu--;
            } else {
              ops[pos++] = 2;
              state = (bch & 8) ? 2 : 0;
              vv--;
            }
          }
        }
        __syncwarp();
      }
      if (lane == 0) s_len[wib] = (ok && done) ? pos : -1;
    }
    if (BLOCK)
      __syncthreads();
    else
      __syncwarp();
    len = s_len[wib];
    used = attempt + 1;
    if (BLOCK)
      __syncthreads();  // s_len is rewritten by the next attempt
    else
      __syncwarp();
  }
  if (tid == 0) {
    a.opsLen[tc.id] = len;
    a.status[tc.id] = len > 0 ? used : 0;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// K4: alignment rows and everything that is counted over them, one warp per candidate (SearchResult.pm:2486-2834 prints
// these rows; ti/tv as CrossmatchSearchEngine.pm:671-674 reads them; SearchResult.pm:2111-2150 adjusts the score;
// SearchResult.pm:1607-1683 is the Kimura pass).
//   rows     gapped query and subject rows in cross_match's display orientation (strand C: the reverse-complemented query
//            ascending, subject forward), spelled from the canonical residue codes, NUL-terminated
//   counts   mismatches / gaps (the percentages' numerators), transitions / transversions / ambiguous, gap starts and
//            bases, aligned query letters per matrix column
//   score    complexity-adjusted: int(raw + (sum c ln f - (sum c ln c - N ln N)) / lambda + .999) in IEEE double with
//            round-to-nearest multiplies and adds in the host's order and the host's own ln() values (ln f and a table of
//            ln k), so the result is the host's bit for bit
//   Kimura   the sequential CpG-aware state machine, run by lane 0 over the finished rows; the logarithm of the final
//            formula stays on the host (three numbers per hit)
// ---------------------------------------------------------------------------------------------------------------
struct HitDev {
  int32_t raw, score;
  int32_t q, subject, strand;
  int32_t i0, j0, i1, j1;       // 1-based strand coordinates of origin and end cell
  int32_t len;                  // alignment columns (0: no traceback)
  int32_t nsub, ndel, nins, ti, tv, amb, gap_inits, gap_bases;
  int32_t ktv, wcb, cpg;        // Kimura: transversions, well-characterised pairs, CpG sites
  int32_t flags;                // kHitKeep | kHitStar | kHitDrop
  double ktrans;                // Kimura: transitions (fractional with the CpG rule)
  int64_t rowOff;               // query row at rows + rowOff (len + 1 bytes), subject row right behind it
  int64_t outOff;               // after the gather: offset of the rows in the packed buffer that crosses PCIe
};
constexpr int kHitKeep = 1;  // reported score >= minscore
constexpr int kHitStar = 2;  // a strictly higher-scoring kept hit of the same query overlaps its query domain
constexpr int kHitDrop = 4;  // removed by masklevel

struct RenderArgs {
  const TraceCand* cand;
  int nCand;
  const uint8_t* codes_fwd;
  const uint8_t* codes_rc;
  const int64_t* qoff;
  const int32_t* qlen;
  const uint8_t* scodes;
  const int64_t* soff;
  const uint8_t* ops;
  const int32_t* opsLen;
  uint8_t col_of_canon[16];
  int nCols;
  double lambda;
  double logfreq[16];     // host ln(freq[col]); NaN-free: see freqUse
  uint8_t freqUse[16];    // freq[col] > 0 && ln(freq[col]) != 0
  const double* logtab;   // host ln(k), k = 0 .. logtabN - 1
  int logtabN;
  int rawMode, minscore;
  char* rows;
  HitDev* out;
};

__device__ __forceinline__ int base_class_dev(int canon) {  // A G purine 0, C T pyrimidine 1, else -1 (canonical: A0 C1 G2 T3)
  return canon == 0 || canon == 2 ? 0 : (canon == 1 || canon == 3 ? 1 : -1);
}

constexpr int kRenderWarps = 4;

__global__ void __launch_bounds__(kRenderWarps * 32) render_alignment_kernel(RenderArgs a) {
  __shared__ int s_cnt[kRenderWarps][16];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int c = blockIdx.x * kRenderWarps + wib;
  if (c >= a.nCand) return;
  const TraceCand tc = a.cand[c];
  const int len = a.opsLen[tc.id];
  HitDev h;
  h.raw = tc.raw;
  h.score = 0;
  h.q = tc.q;
  h.subject = tc.subject;
  h.strand = tc.strand;
  h.i0 = tc.i0;
  h.j0 = tc.j0;
  h.i1 = tc.i1;
  h.j1 = tc.j1;
  h.len = len > 0 ? len : 0;
  h.nsub = h.ndel = h.nins = h.ti = h.tv = h.amb = h.gap_inits = h.gap_bases = 0;
  h.ktv = h.wcb = h.cpg = 0;
  h.flags = 0;
  h.ktrans = 0;
  h.rowOff = 2 * tc.opsOff + 2 * (int64_t)tc.id;  // rows need len + 1 bytes each, ops had n + m >= len
  h.outOff = -1;
  if (len <= 0) {
    if (lane == 0) a.out[tc.id] = h;
    return;
  }
  if (lane < 16) s_cnt[wib][lane] = 0;
  __syncwarp();
  const uint8_t* qc = (tc.strand ? a.codes_rc : a.codes_fwd) + a.qoff[tc.q];
  const uint8_t* sc = a.scodes + a.soff[tc.subject];
  const uint8_t* ops = a.ops + tc.opsOff;  // written back to front by the walk
  char* qrow = a.rows + h.rowOff;
  char* srow = qrow + len + 1;
  const char* L = "ACGTRYKMSWBDHVNX";
  int qi0 = tc.i0 - 1, si0 = tc.j0 - 1, carry = 0;
  int nsub = 0, ndel = 0, nins = 0, ti = 0, tv = 0, amb = 0, ginit = 0;
  const unsigned lt = (1u << lane) - 1u;
  for (int c0 = 0; c0 < len; c0 += 32) {
    const int col = c0 + lane;
    const bool valid = col < len;
    const int op = valid ? (int)ops[len - 1 - col] : 3;
    const unsigned mq = __ballot_sync(0xffffffffu, valid && op != 2);
    const unsigned ms = __ballot_sync(0xffffffffu, valid && op != 1);
    int prev = __shfl_up_sync(0xffffffffu, op, 1);
    if (lane == 0) prev = carry;
    if (valid) {
      const int qk = qc[qi0 + __popc(mq & lt)], sk = sc[si0 + __popc(ms & lt)];
      if (op == 0) {
        qrow[col] = L[qk];
        srow[col] = L[sk];
        atomicAdd(&s_cnt[wib][a.col_of_canon[qk]], 1);
        if (qk != sk) {
          nsub++;
          const int ca = base_class_dev(qk), cb = base_class_dev(sk);
          if (ca < 0 || cb < 0)
            amb++;
          else if (ca == cb)
            ti++;
          else
            tv++;
        }
      } else if (op == 1) {
        qrow[col] = L[qk];
        srow[col] = '-';
        nins++;
        ginit += prev != 1;
      } else {
        qrow[col] = '-';
        srow[col] = L[sk];
        ndel++;
        ginit += prev != 2;
      }
    }
    qi0 += __popc(mq);
    si0 += __popc(ms);
    carry = __shfl_sync(0xffffffffu, op, 31);
  }
  if (lane == 0) {
    qrow[len] = 0;
    srow[len] = 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    nsub += __shfl_xor_sync(0xffffffffu, nsub, o);
    ndel += __shfl_xor_sync(0xffffffffu, ndel, o);
    nins += __shfl_xor_sync(0xffffffffu, nins, o);
    ti += __shfl_xor_sync(0xffffffffu, ti, o);
    tv += __shfl_xor_sync(0xffffffffu, tv, o);
    amb += __shfl_xor_sync(0xffffffffu, amb, o);
    ginit += __shfl_xor_sync(0xffffffffu, ginit, o);
  }
  __syncwarp();  // rows and counts of every lane are visible to lane 0
  if (lane != 0) return;
  h.nsub = nsub;
  h.ndel = ndel;
  h.nins = nins;
  h.ti = ti;
  h.tv = tv;
  h.amb = amb;
  h.gap_inits = ginit;
  h.gap_bases = ndel + nins;
  // complexity-adjusted score (host_io.cpp:complexity_adjust, SearchResult.pm:2111-2150)
  int score = tc.raw;
  if (!a.rawMode) {
    double t_factor = 0, t_sum = 0, t_counts = 0;
    for (int k = 0; k < a.nCols; k++) {
      const int cnt = s_cnt[wib][k];
      if (cnt > 0 && a.freqUse[k]) {
        const double kd = (double)cnt;
        const double lk = cnt < a.logtabN ? a.logtab[cnt] : log(kd);
        t_factor = __dadd_rn(t_factor, __dmul_rn(kd, lk));
        t_sum = __dadd_rn(t_sum, __dmul_rn(kd, a.logfreq[k]));
        t_counts = __dadd_rn(t_counts, kd);
      }
    }
    if (t_counts != 0) {
      const int tc_i = (int)t_counts;
      const double lt_ = tc_i < a.logtabN ? a.logtab[tc_i] : log(t_counts);
      t_factor = __dsub_rn(t_factor, __dmul_rn(t_counts, lt_));
    }
    t_sum = __dsub_rn(t_sum, t_factor);
    const double adj = __dadd_rn(__dadd_rn((double)tc.raw, __ddiv_rn(t_sum, a.lambda)), .999);
    score = (int)adj;
    if (score < 0) score = 0;
  }
  h.score = score;
  if (score >= a.minscore) h.flags |= kHitKeep;
  // Kimura divergence counts (host_io.cpp:kimura_divergence): strand C reads the rows back to front through the
  // complement, i.e. as CrossmatchSearchEngine.pm:824-837 stores them
  {
    auto base = [](char ch) { return ch == 'A' ? 0 : ch == 'G' ? 1 : ch == 'C' ? 2 : ch == 'T' ? 3 : -1; };
    auto comp = [](char ch) {
      switch (ch) {
        case 'A': return 'T';
        case 'C': return 'G';
        case 'G': return 'C';
        case 'T': return 'A';
        case 'Y': return 'R';
        case 'R': return 'Y';
        case 'M': return 'K';
        case 'K': return 'M';
        case 'H': return 'D';
        case 'B': return 'V';
        case 'V': return 'B';
        case 'D': return 'H';
        default: return ch;
      }
    };
    double transitions = 0, prevTrans = 0;
    int transversions = 0, cpg = 0, wcb = 0;
    char pS = 0;
    for (int i = 0; i < len; i++) {
      const int at = tc.strand ? len - 1 - i : i;
      const char sb = tc.strand ? comp(srow[at]) : srow[at];
      if (sb == '-') continue;
      const char qb = tc.strand ? comp(qrow[at]) : qrow[at];
      const int bq = base(qb), bs = base(sb);
      const bool wc = bq >= 0 && bs >= 0;
      wcb += wc;
      const int mt = (!wc || bq == bs) ? 0 : ((bq >> 1) == (bs >> 1) ? 1 : 2);
      if (pS == 'C' && sb == 'G') {
        cpg++;
        if (mt == 1)
          prevTrans = __dadd_rn(prevTrans, 1.0);
        else if (mt == 2)
          transversions++;
        if (prevTrans == 2)
          prevTrans = 1;
        else if (prevTrans == 1)
          prevTrans = 1.0 / 10;
      } else {
        transitions = __dadd_rn(transitions, prevTrans);
        prevTrans = 0;
        if (mt == 1)
          prevTrans = 1;
        else if (mt == 2)
          transversions++;
      }
      pS = sb;
    }
    transitions = __dadd_rn(transitions, prevTrans);
    h.ktrans = transitions;
    h.ktv = transversions;
    h.wcb = wcb;
    h.cpg = cpg;
  }
  a.out[tc.id] = h;
}

// ---------------------------------------------------------------------------------------------------------------
// K5: what is decided over all hits of one query record, on the device so that only survivors cross PCIe.
//   star_rank_kernel   one thread per hit against every hit of its query: the '*' flag (a strictly higher-scoring kept hit
//                      overlaps the query domain, CrossmatchSearchEngine.pm:99-100) and the hit's rank by descending score
//   masklevel_kernel   cross_match's masklevel (repeatmasker.help:714-728; fixtures seq1-2-ml{97,98}.out): one warp per
//                      query walks its hits in rank order, equal scores as a group, against a bitmap of the query positions
//                      covered by kept, strictly higher-scoring hits; a hit stays when at least (100 - masklevel) % of its
//                      domain is uncovered
// Query coordinates here are forward-strand (qs, qe), 1-based inclusive.
// ---------------------------------------------------------------------------------------------------------------
struct FinalizeArgs {
  HitDev* hits;            // all candidates of the pass, grouped by query record
  const int32_t* qFirst;   // [nQueries] first candidate of the query
  const int32_t* qCount;   // [nQueries]
  const int32_t* qlen;
  int nQueries;
  int masklevel;
  int32_t* order;          // [nCand] candidates of a query by rank (only kept ones), at qFirst[q] + rank
  int32_t* nKept;          // [nQueries]
  uint32_t* cover;         // [coverOff[q] ...] one bit per query position
  const int64_t* coverOff; // [nQueries] word offset of the query's bitmap
};

__device__ __forceinline__ void fwd_domain(const HitDev& h, int Q, int& qs, int& qe) {
  if (!h.strand) {
    qs = h.i0;
    qe = h.i1;
  } else {
    qs = Q + 1 - h.i1;
    qe = Q + 1 - h.i0;
  }
}

__global__ void star_rank_kernel(FinalizeArgs a) {
  const int q = blockIdx.y;
  const int n = a.qCount[q], first = a.qFirst[q], Q = a.qlen[q];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ int s_kept;
  if (threadIdx.x == 0) s_kept = 0;
  __syncthreads();
  if (i < n) {
    HitDev& me = a.hits[first + i];
    if (me.flags & kHitKeep) {
      int qs, qe;
      fwd_domain(me, Q, qs, qe);
      bool star = false;
      int rank = 0;
      for (int j = 0; j < n; j++) {
        const HitDev& o = a.hits[first + j];
        if (!(o.flags & kHitKeep)) continue;
        int os, oe;
        fwd_domain(o, Q, os, oe);
        if (o.score > me.score) {
          rank++;
          star |= os <= qe && oe >= qs;
        } else if (o.score == me.score && j < i)
          rank++;
      }
      if (star) me.flags |= kHitStar;
      a.order[first + rank] = i;
      atomicAdd(&s_kept, 1);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && s_kept) atomicAdd(&a.nKept[q], s_kept);
}

__global__ void masklevel_kernel(FinalizeArgs a) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= a.nQueries) return;
  const int nk = a.nKept[q], first = a.qFirst[q], Q = a.qlen[q];
  uint32_t* cov = a.cover + a.coverOff[q];
  auto covered = [&](int qs, int qe) {  // covered positions of [qs, qe] (1-based); bit p-1 of the bitmap = position p
    const int w0 = (qs - 1) >> 5, w1 = (qe - 1) >> 5;
    int cnt = 0;
    for (int w = w0 + lane; w <= w1; w += 32) {
      uint32_t m = cov[w];
      if (w == w0) m &= ~0u << ((qs - 1) & 31);
      if (w == w1) m &= ~0u >> (31 - ((qe - 1) & 31));
      cnt += __popc(m);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    return cnt;
  };
  int g = 0;
  while (g < nk) {
    const int sc = a.hits[first + a.order[first + g]].score;
    int ge = g;
    while (ge < nk && a.hits[first + a.order[first + ge]].score == sc) ge++;
    for (int k = g; k < ge; k++) {
      HitDev& h = a.hits[first + a.order[first + k]];
      int qs, qe;
      fwd_domain(h, Q, qs, qe);
      const long lenq = qe - qs + 1, unc = lenq - covered(qs, qe);
      if (lane == 0 && unc * 100 < (long)(100 - a.masklevel) * lenq) h.flags |= kHitDrop;
    }
    __syncwarp();
    for (int k = g; k < ge; k++) {
      const HitDev& h = a.hits[first + a.order[first + k]];
      if (h.flags & kHitDrop) continue;
      int qs, qe;
      fwd_domain(h, Q, qs, qe);
      const int w0 = (qs - 1) >> 5, w1 = (qe - 1) >> 5;
      for (int w = w0 + lane; w <= w1; w += 32) {
        uint32_t m = ~0u;
        if (w == w0) m &= ~0u << ((qs - 1) & 31);
        if (w == w1) m &= ~0u >> (31 - ((qe - 1) & 31));
        cov[w] |= m;
      }
      __syncwarp();
    }
    g = ge;
  }
}

// Survivors' rows, packed: a single-block exclusive scan of the row bytes (2 * (len + 1) per surviving hit) followed by
// one warp per hit copying its two rows.
struct GatherArgs {
  HitDev* hits;
  int nCand;
  const char* rows;
  char* packed;
  long long* total;  // bytes of the packed buffer
};

__global__ void __launch_bounds__(1024) scan_rows_kernel(GatherArgs a) {
  __shared__ long long s_warp[32];
  __shared__ long long s_carry;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < a.nCand; base += 1024) {
    const int i = base + threadIdx.x;
    long long v = 0;
    if (i < a.nCand) {
      const HitDev& h = a.hits[i];
      if ((h.flags & kHitKeep) && !(h.flags & kHitDrop)) v = 2ll * (h.len + 1);
    }
    long long x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[wib] = x;
    __syncthreads();
    if (wib == 0) {
      long long w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      s_warp[lane] = w;
    }
    __syncthreads();
    const long long before = s_carry + (wib ? s_warp[wib - 1] : 0) + x - v;
    if (i < a.nCand && v) a.hits[i].outOff = before;
    __syncthreads();
    if (threadIdx.x == 0) s_carry += s_warp[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) *a.total = s_carry;
}

__global__ void gather_rows_kernel(GatherArgs a) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= a.nCand) return;
  const HitDev h = a.hits[i];
  if (h.outOff < 0) return;
  const char* src = a.rows + h.rowOff;
  char* dst = a.packed + h.outOff;
  for (int k = lane; k < 2 * (h.len + 1); k += 32) dst[k] = src[k];
}

// ---------------------------------------------------------------------------------------------------------------
// The two collection filters RepeatMasker applies after a search (RepeatMasker:2838-2987), query-domain flavour, with the
// reference's own arithmetic:
//   perl_masklevel_kernel   SearchResultCollection::maskLevelFilter (SearchResultCollection.pm:198-358): records sorted by
//                           (name, score descending, query span ascending), stable; a record is deleted when SOME kept
//                           record of the same name overlaps it and sprintf("%0.0f", 100 * outside / length) <= 100 -
//                           masklevel (pairwise, not against the union; round-half-even like printf).  One warp per name:
//                           the walk down the sorted list is sequential, the kept domains are checked 32 at a time.
//   contained_kernel        SearchResultCollection::filterContainedResults (:379-470): records sorted by (name, start,
//                           end descending); for every record i in that order the following records are scanned until
//                           the first LIVE one that starts beyond end_i, and a scanned record is deleted when it ends
//                           before end_i with a strictly lower score.  As in the reference the scan does not stop at
//                           the end of a name (so the order of the names matters), deleted records are skipped without
//                           ending the scan, and a deleted record i still deletes.
// ---------------------------------------------------------------------------------------------------------------
struct CollectionFilterArgs {
  const int32_t* start;   // [n] 1-based inclusive, in collection order
  const int32_t* end;
  const int32_t* score;
  const int32_t* order;   // [n] record numbers in the filter's sort order, grouped by name
  const int32_t* gFirst;  // [nGroups] first position of the group in `order`
  const int32_t* gCount;
  int nGroups, n, masklevel;
  int32_t* domains;       // [n] scratch: kept records of a group, at gFirst[g] + k
  uint8_t* deleted;       // [n] out, by record number
};

__global__ void perl_masklevel_kernel(CollectionFilterArgs a) {
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= a.nGroups) return;
  const int first = a.gFirst[g], cnt = a.gCount[g];
  const double thr = (double)(100 - a.masklevel);
  int nd = 0;
  for (int k = 0; k < cnt; k++) {
    const int rec = a.order[first + k];
    const int b1 = a.start[rec], e1 = a.end[rec];
    const double len = (double)(e1 - b1 + 1);
    bool del = false;
    for (int d = lane; d < nd; d += 32) {
      const int dom = a.domains[first + d];
      const int db = a.start[dom], de = a.end[dom];
      if (db > e1 || de < b1) continue;
      const int inter = min(e1, de) - max(b1, db) + 1;
      const double pct = rint(__dmul_rn(__ddiv_rn((double)((e1 - b1 + 1) - inter), len), 100.0));
      del |= pct <= thr;
    }
    del = __any_sync(0xffffffffu, del);
    if (lane == 0) {
      if (del)
        a.deleted[rec] = 1;
      else
        a.domains[first + nd] = rec;
    }
    nd += del ? 0 : 1;
    __syncwarp();
  }
}

__global__ void contained_kernel(CollectionFilterArgs a) {
  // One warp walks the whole sorted list (see the note above: the reference's inner loop runs across name boundaries).
  // For record i the lanes look at 32 successors at a time: already deleted ones are skipped (they neither stop the scan
  // nor are examined), the first live one that starts beyond end_i stops it, live ones before that are deleted when
  // they end before end_i with a lower score.
  const int lane = threadIdx.x & 31;
  if (blockIdx.x != 0 || threadIdx.x >= 32) return;
  for (int i = 0; i < a.n; i++) {
    const int r1 = a.order[i];
    const int e1 = a.end[r1], s1 = a.score[r1];
    for (int j0 = i + 1; j0 < a.n; j0 += 32) {
      const int j = j0 + lane;
      bool live = false, stop = false, kill = false;
      int r2 = 0;
      if (j < a.n) {
        r2 = a.order[j];
        live = !a.deleted[r2];
        stop = live && a.start[r2] > e1;
        kill = live && !stop && a.end[r2] < e1 && s1 > a.score[r2];
      }
      const unsigned stops = __ballot_sync(0xffffffffu, stop);
      const unsigned before = stops ? ((1u << (__ffs(stops) - 1)) - 1u) : 0xffffffffu;
      if (kill && ((before >> lane) & 1u)) a.deleted[r2] = 1;
      if (stops) break;
    }
    __syncwarp();
  }
}

}  // namespace rmgpu
