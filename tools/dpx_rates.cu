// dpx_rates.cu -- issue rates of the integer-SIMD instructions K1 is made of, and register-resident loops of the
// candidate cell recurrences, measured on the GPU at hand.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
// -o dpx_rates tools/dpx_rates.cu ; run with no arguments.  Prints warp-instructions per clock per SM (from the SM
// clock reported by the device) and, for the cell loops, GCUPS.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t prmt(uint32_t lo, uint32_t hi, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(lo), "r"(hi), "r"(sel));
  return d;
}

enum Op { VIADDMAX, VIADDMAX_RELU, VIMAX3, VIMAX3_RELU, VMAX2, VADD2, PRMT, IADD, IMAD, LOP, MIX_VIADDMAX_IMAD, MIX_VIADDMAX_PRMT,
          MIX_VIADDMAX_VIMAX3, MIX_VIMAX3_VMAX2, MIX_VIADDMAX_VADD2, MIX_VIADDMAX_IADD, VMAX2_RELU, MIX_VADD2_IMAD, MIX_VMAX2_VIADDMAX, MIX_VMAX2_VADD2, MIX3_VIADDMAX_VADD2_IMAD, MIX_VMAX2_IMAD, MIX_PRMT_VADD2, NOPS };
const char* kNames[] = {"VIADDMNMX.S16x2", "VIADDMNMX.S16x2.RELU", "VIMNMX3.S16x2", "VIMNMX3.S16x2.RELU", "VIMNMX.S16x2 (2-input)",
                        "VIADD.16x2", "PRMT", "IADD3 (32-bit add)", "IMAD (x*1+y)", "LOP3", "mix VIADDMNMX+IMAD 1:1", "mix VIADDMNMX+PRMT 1:1",
                        "mix VIADDMNMX+VIMNMX3 1:1", "mix VIMNMX3+VIMNMX 1:1", "mix VIADDMNMX+VIADD.16x2 1:1", "mix VIADDMNMX+IADD3 1:1", "VIMNMX.S16x2.RELU (2-input)", "mix VIADD.16x2+IMAD 1:1", "mix VIMNMX+VIADDMNMX 1:1", "mix VIMNMX+VIADD.16x2 1:1", "mix VIADDMNMX+VIADD.16x2+IMAD 1:1:1", "mix VIMNMX+IMAD 1:1", "mix PRMT+VIADD.16x2 1:1"};

template <int OP>
__global__ void rate_kernel(uint32_t* out, int iters, uint32_t c1, uint32_t c2, uint32_t one) {
  uint32_t x[8], y[8], z[4];
#pragma unroll
  for (int r = 0; r < 8; r++) {
    x[r] = threadIdx.x + r;
    y[r] = threadIdx.x * 3 + r;
    z[r & 3] = threadIdx.x * 5 + r;
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
      if (OP == VIADDMAX) x[r] = __viaddmax_s16x2(x[r], c1, c2);
      if (OP == VIADDMAX_RELU) x[r] = __viaddmax_s16x2_relu(x[r], c1, c2);
      if (OP == VIMAX3) x[r] = __vimax3_s16x2(x[r], c1, y[r]);
      if (OP == VIMAX3_RELU) x[r] = __vimax3_s16x2_relu(x[r], c1, y[r]);
      if (OP == VMAX2) x[r] = (it & 1) ? __vmaxs2(x[r], c1) : __vmins2(x[r], c2);
      if (OP == VADD2) x[r] = __vadd2(x[r], c1);
      if (OP == PRMT) x[r] = prmt(x[r], c1, c2);
      if (OP == IADD) {
        x[r] = x[r] + y[r];
        y[r] = y[r] + x[r];
      }
      if (OP == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[r]) : "r"(one), "r"(c1));
      if (OP == LOP) x[r] = (x[r] ^ c1) & c2;
      if (OP == MIX_VIADDMAX_IMAD) {
        x[r] = __viaddmax_s16x2(x[r], c1, c2);
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[r]) : "r"(one), "r"(c1));
      }
      if (OP == MIX_VIADDMAX_PRMT) {
        x[r] = __viaddmax_s16x2(x[r], c1, c2);
        y[r] = prmt(y[r], c1, c2);
      }
      if (OP == MIX_VIADDMAX_VIMAX3) {
        x[r] = __viaddmax_s16x2(x[r], c1, c2);
        y[r] = __vimax3_s16x2(y[r], c1, c2);
      }
      if (OP == MIX_VIMAX3_VMAX2) {
        x[r] = __vimax3_s16x2(x[r], c1, c2);
        asm volatile("max.s16x2 %0, %0, %1;" : "+r"(y[r]) : "r"(c1));
      }
      if (OP == MIX_VIADDMAX_VADD2) {
        x[r] = __viaddmax_s16x2(x[r], c1, c2);
        y[r] = __vadd2(y[r], c1);
      }
      if (OP == VMAX2_RELU) x[r] = __vimax_s16x2_relu(x[r], c1 + r);
      if (OP == MIX_VADD2_IMAD) {
        x[r] = __vadd2(x[r], c1);
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[r]) : "r"(one), "r"(c1));
      }
      if (OP == MIX_VMAX2_VIADDMAX) {
        x[r] = (it & 1) ? __vmaxs2(x[r], c1) : __vmins2(x[r], c2);
        y[r] = __viaddmax_s16x2(y[r], c1, c2);
      }
      if (OP == MIX_VMAX2_VADD2) {
        x[r] = (it & 1) ? __vmaxs2(x[r], c1) : __vmins2(x[r], c2);
        y[r] = __vadd2(y[r], c1);
      }
      if (OP == MIX3_VIADDMAX_VADD2_IMAD) {
        x[r] = __viaddmax_s16x2(x[r], c1, c2);
        y[r] = __vadd2(y[r], c1);
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(z[r & 3]) : "r"(one), "r"(c1));
      }
      if (OP == MIX_VMAX2_IMAD) {
        x[r] = (it & 1) ? __vmaxs2(x[r], c1) : __vmins2(x[r], c2);
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[r]) : "r"(one), "r"(c1));
      }
      if (OP == MIX_PRMT_VADD2) {
        x[r] = prmt(x[r], c1, c2);
        y[r] = __vadd2(y[r], c1);
      }
      if (OP == MIX_VIADDMAX_IADD) {
        x[r] = __viaddmax_s16x2(x[r], c1, c2);
        y[r] = y[r] + z[r & 3];
        if ((r & 1) == 1) z[r & 3] = z[r & 3] + y[r];
      }
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int r = 0; r < 8; r++) s ^= x[r] ^ y[r] ^ z[r & 3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- cell loops: 32 lanes x R rows x iters columns x 2 strands per warp, no shuffles, no memory ------------------
// A: the round-1 cell.  H >= 0, E' = E - gi, F' = F - gi.
template <int R>
__global__ void cellA(uint32_t* out, int iters, uint32_t seed, uint32_t gi2, uint32_t ie2, uint32_t de2) {
  uint32_t Plo[R], Phi[R], H[R], E[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    Plo[r] = 0x090af1f0u + r * 0x01010101u + threadIdx.x;
    Phi[r] = 0xf1f60a09u ^ (r * 0x00010001u);
    H[r] = 0;
    E[r] = 0;
  }
  uint32_t best = 0, sel = 0x9180u | (seed & 0x0303u), diag = 0, hu = 0, f = 0;
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
    }
    if ((it & 7) == 7) {
#pragma unroll
      for (int r = 0; r < R; r++) best = __vmaxs2(best, E[r]);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = best;
}

// A1: round-1 cell, E update split into VIADD.16x2 + VIMNMX.
template <int R>
__global__ void cellA1(uint32_t* out, int iters, uint32_t seed, uint32_t gi2, uint32_t ie2, uint32_t de2) {
  uint32_t Plo[R], Phi[R], H[R], E[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    Plo[r] = 0x090af1f0u + r * 0x01010101u + threadIdx.x;
    Phi[r] = 0xf1f60a09u ^ (r * 0x00010001u);
    H[r] = 0;
    E[r] = 0;
  }
  uint32_t best = 0, sel = 0x9180u | (seed & 0x0303u), diag = 0, hu = 0, f = 0;
  for (int it = 0; it < iters; it++) {
    uint32_t hd = diag;
    diag = hu;
    sel = ((sel + 0x0101u) & 0x0303u) | 0x9180u;
#pragma unroll
    for (int r = 0; r < R; r++) {
      uint32_t s = prmt(Plo[r], Phi[r], sel);
      uint32_t hl = H[r];
      E[r] = __vmaxs2(__vadd2(E[r], ie2), hl);
      uint32_t d = __vadd2(hd, s);
      uint32_t tt = __viaddmax_s16x2_relu(E[r], gi2, d);
      uint32_t h = __viaddmax_s16x2(f, gi2, tt);
      f = __viaddmax_s16x2(f, de2, tt);
      hd = hl;
      H[r] = h;
      hu = h;
    }
    if ((it & 7) == 7) {
#pragma unroll
      for (int r = 0; r < R; r++) best = __vmaxs2(best, E[r]);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = best;
}

// A2: E and H split.
template <int R>
__global__ void cellA2(uint32_t* out, int iters, uint32_t seed, uint32_t gi2, uint32_t ie2, uint32_t de2) {
  uint32_t Plo[R], Phi[R], H[R], E[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    Plo[r] = 0x090af1f0u + r * 0x01010101u + threadIdx.x;
    Phi[r] = 0xf1f60a09u ^ (r * 0x00010001u);
    H[r] = 0;
    E[r] = 0;
  }
  uint32_t best = 0, sel = 0x9180u | (seed & 0x0303u), diag = 0, hu = 0, f = 0;
  for (int it = 0; it < iters; it++) {
    uint32_t hd = diag;
    diag = hu;
    sel = ((sel + 0x0101u) & 0x0303u) | 0x9180u;
#pragma unroll
    for (int r = 0; r < R; r++) {
      uint32_t s = prmt(Plo[r], Phi[r], sel);
      uint32_t hl = H[r];
      E[r] = __vmaxs2(__vadd2(E[r], ie2), hl);
      uint32_t d = __vadd2(hd, s);
      uint32_t tt = __viaddmax_s16x2_relu(E[r], gi2, d);
      uint32_t h = __vmaxs2(__vadd2(f, gi2), tt);
      f = __viaddmax_s16x2(f, de2, tt);
      hd = hl;
      H[r] = h;
      hu = h;
    }
    if ((it & 7) == 7) {
#pragma unroll
      for (int r = 0; r < R; r++) best = __vmaxs2(best, E[r]);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = best;
}

// A3: E, H and T split.
template <int R>
__global__ void cellA3(uint32_t* out, int iters, uint32_t seed, uint32_t gi2, uint32_t ie2, uint32_t de2) {
  uint32_t Plo[R], Phi[R], H[R], E[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    Plo[r] = 0x090af1f0u + r * 0x01010101u + threadIdx.x;
    Phi[r] = 0xf1f60a09u ^ (r * 0x00010001u);
    H[r] = 0;
    E[r] = 0;
  }
  uint32_t best = 0, sel = 0x9180u | (seed & 0x0303u), diag = 0, hu = 0, f = 0;
  for (int it = 0; it < iters; it++) {
    uint32_t hd = diag;
    diag = hu;
    sel = ((sel + 0x0101u) & 0x0303u) | 0x9180u;
#pragma unroll
    for (int r = 0; r < R; r++) {
      uint32_t s = prmt(Plo[r], Phi[r], sel);
      uint32_t hl = H[r];
      E[r] = __vmaxs2(__vadd2(E[r], ie2), hl);
      uint32_t d = __vadd2(hd, s);
      uint32_t tt = __vimax_s16x2_relu(__vadd2(E[r], gi2), d);
      uint32_t h = __vmaxs2(__vadd2(f, gi2), tt);
      f = __viaddmax_s16x2(f, de2, tt);
      hd = hl;
      H[r] = h;
      hu = h;
    }
    if ((it & 7) == 7) {
#pragma unroll
      for (int r = 0; r < R; r++) best = __vmaxs2(best, E[r]);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = best;
}

// A2b: H and T split.
template <int R>
__global__ void cellA2b(uint32_t* out, int iters, uint32_t seed, uint32_t gi2, uint32_t ie2, uint32_t de2) {
  uint32_t Plo[R], Phi[R], H[R], E[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    Plo[r] = 0x090af1f0u + r * 0x01010101u + threadIdx.x;
    Phi[r] = 0xf1f60a09u ^ (r * 0x00010001u);
    H[r] = 0;
    E[r] = 0;
  }
  uint32_t best = 0, sel = 0x9180u | (seed & 0x0303u), diag = 0, hu = 0, f = 0;
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
      uint32_t tt = __vimax_s16x2_relu(__vadd2(E[r], gi2), d);
      uint32_t h = __vmaxs2(__vadd2(f, gi2), tt);
      f = __viaddmax_s16x2(f, de2, tt);
      hd = hl;
      H[r] = h;
      hu = h;
    }
    if ((it & 7) == 7) {
#pragma unroll
      for (int r = 0; r < R; r++) best = __vmaxs2(best, E[r]);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = best;
}

// A1h: H split.
template <int R>
__global__ void cellA1h(uint32_t* out, int iters, uint32_t seed, uint32_t gi2, uint32_t ie2, uint32_t de2) {
  uint32_t Plo[R], Phi[R], H[R], E[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    Plo[r] = 0x090af1f0u + r * 0x01010101u + threadIdx.x;
    Phi[r] = 0xf1f60a09u ^ (r * 0x00010001u);
    H[r] = 0;
    E[r] = 0;
  }
  uint32_t best = 0, sel = 0x9180u | (seed & 0x0303u), diag = 0, hu = 0, f = 0;
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
      uint32_t h = __vmaxs2(__vadd2(f, gi2), tt);
      f = __viaddmax_s16x2(f, de2, tt);
      hd = hl;
      H[r] = h;
      hu = h;
    }
    if ((it & 7) == 7) {
#pragma unroll
      for (int r = 0; r < R; r++) best = __vmaxs2(best, E[r]);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = best;
}

// L: long chain with VIMNMX3.  Hg = H + gi stored per row, true E and F, profile biased by -gi (s' = s - gi >= 0):
//   E  = max(E + ie, Hg_left)             VIADDMNMX
//   D  = Hg_diag + s'                     VIADD.16x2  (other pipe)
//   H  = max(0, D, E, F)                  VIMNMX3.RELU
//   Hg = H + gi                           VIADD.16x2  (other pipe)
//   F  = max(F + de, Hg)                  VIADDMNMX   -> next row
template <int R>
__global__ void cellL(uint32_t* out, int iters, uint32_t seed, uint32_t gi2, uint32_t ie2, uint32_t de2) {
  uint32_t Plo[R], Phi[R], Hg[R], E[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    Plo[r] = 0x090a0100u + r * 0x01010101u + (threadIdx.x & 3);
    Phi[r] = 0x01060a09u ^ (r * 0x00010001u);
    Hg[r] = gi2;
    E[r] = gi2;
  }
  uint32_t best = 0, sel = 0x9180u | (seed & 0x0303u), diag = gi2, hgu = gi2, f = gi2;
  for (int it = 0; it < iters; it++) {
    uint32_t hd = diag;
    diag = hgu;
    sel = ((sel + 0x0101u) & 0x0303u) | 0x9180u;
#pragma unroll
    for (int r = 0; r < R; r++) {
      uint32_t s = prmt(Plo[r], Phi[r], sel);
      uint32_t hl = Hg[r];
      E[r] = __viaddmax_s16x2(E[r], ie2, hl);
      uint32_t d = __vadd2(hd, s);
      uint32_t h = __vimax3_s16x2_relu(d, E[r], f);
      uint32_t hg = __vadd2(h, gi2);
      f = __viaddmax_s16x2(f, de2, hg);
      hd = hl;
      Hg[r] = hg;
      hgu = hg;
    }
    if ((it & 7) == 7) {
#pragma unroll
      for (int r = 0; r < R; r++) best = __vmaxs2(best, E[r]);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = best;
}

// Lp: the same with the diagonal sum passed through an IMAD (x*1) so that ptxas cannot fuse the add into a VIADDMNMX.
// L: long chain with VIMNMX3.  Hg = H + gi stored per row, true E and F, profile biased by -gi (s' = s - gi >= 0):
//   E  = max(E + ie, Hg_left)             VIADDMNMX
//   D  = Hg_diag + s'                     VIADD.16x2  (other pipe)
//   H  = max(0, D, E, F)                  VIMNMX3.RELU
//   Hg = H + gi                           VIADD.16x2  (other pipe)
//   F  = max(F + de, Hg)                  VIADDMNMX   -> next row
template <int R>
__global__ void cellLp(uint32_t* out, int iters, uint32_t seed, uint32_t gi2, uint32_t ie2, uint32_t de2, uint32_t one) {
  uint32_t Plo[R], Phi[R], Hg[R], E[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    Plo[r] = 0x090a0100u + r * 0x01010101u + (threadIdx.x & 3);
    Phi[r] = 0x01060a09u ^ (r * 0x00010001u);
    Hg[r] = gi2;
    E[r] = gi2;
  }
  uint32_t best = 0, sel = 0x9180u | (seed & 0x0303u), diag = gi2, hgu = gi2, f = gi2;
  for (int it = 0; it < iters; it++) {
    uint32_t hd = diag;
    diag = hgu;
    sel = ((sel + 0x0101u) & 0x0303u) | 0x9180u;
#pragma unroll
    for (int r = 0; r < R; r++) {
      uint32_t s = prmt(Plo[r], Phi[r], sel);
      uint32_t hl = Hg[r];
      E[r] = __viaddmax_s16x2(E[r], ie2, hl);
      uint32_t d = __vadd2(hd, s) * one;
      uint32_t h = __vimax3_s16x2_relu(d, E[r], f);
      uint32_t hg = __vadd2(h, gi2);
      f = __viaddmax_s16x2(f, de2, hg);
      hd = hl;
      Hg[r] = hg;
      hgu = hg;
    }
    if ((it & 7) == 7) {
#pragma unroll
      for (int r = 0; r < R; r++) best = __vmaxs2(best, E[r]);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = best;
}


// ---- the same cells with the gap constants as compile-time literals: VIADDMNMX / VIADD.16x2 take them as 32-bit immediates
// (two register operands instead of three).  Ai = cell A, Li / Lpi = the VIMNMX3 cells (4 alu-pipe instructions per packed cell).
template <int R>
__global__ void cellAi(uint32_t* out, int iters, uint32_t seed) {
  constexpr uint32_t gi2 = 0xffe2ffe2u, ie2 = 0xfffafffau, de2 = 0xfffbfffbu;
  uint32_t Plo[R], Phi[R], H[R], E[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    Plo[r] = 0x090af1f0u + r * 0x01010101u + threadIdx.x;
    Phi[r] = 0xf1f60a09u ^ (r * 0x00010001u);
    H[r] = 0;
    E[r] = 0;
  }
  uint32_t best = 0, sel = 0x9180u | (seed & 0x0303u), diag = 0, hu = 0, f = 0;
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
    }
    if ((it & 7) == 7) {
#pragma unroll
      for (int r = 0; r < R; r++) best = __vmaxs2(best, E[r]);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = best;
}

template <int R, bool LAUNDER>
__global__ void cellLi(uint32_t* out, int iters, uint32_t seed, uint32_t one) {
  constexpr uint32_t gi2 = 0xffe2ffe2u, ie2 = 0xfffafffau, de2 = 0xfffbfffbu;
  uint32_t Plo[R], Phi[R], Hg[R], E[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    Plo[r] = 0x090a0100u + r * 0x01010101u + (threadIdx.x & 3);
    Phi[r] = 0x01060a09u ^ (r * 0x00010001u);
    Hg[r] = gi2;
    E[r] = gi2;
  }
  uint32_t best = 0, sel = 0x9180u | (seed & 0x0303u), diag = gi2, hgu = gi2, f = gi2;
  for (int it = 0; it < iters; it++) {
    uint32_t hd = diag;
    diag = hgu;
    sel = ((sel + 0x0101u) & 0x0303u) | 0x9180u;
#pragma unroll
    for (int r = 0; r < R; r++) {
      uint32_t s = prmt(Plo[r], Phi[r], sel);
      uint32_t hl = Hg[r];
      E[r] = __viaddmax_s16x2(E[r], ie2, hl);
      uint32_t d = __vadd2(hd, s);
      if (LAUNDER) d *= one;
      uint32_t h = __vimax3_s16x2_relu(d, E[r], f);
      uint32_t hg = __vadd2(h, gi2);
      f = __viaddmax_s16x2(f, de2, hg);
      hd = hl;
      Hg[r] = hg;
      hgu = hg;
    }
    if ((it & 7) == 7) {
#pragma unroll
      for (int r = 0; r < R; r++) best = __vmaxs2(best, E[r]);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = best;
}

// S: round-1 recurrence with the score looked up in shared memory instead of PRMT: a table of packed (forward, reverse)
// scores per (subject residue, query class pair), replicated per lane (entry stride 128 B: bank = lane, no conflicts);
// address = row term (fixed per row and ticket) + column term (changes per column).  IMADADDR: the address add as an IMAD
// (fmaheavy pipe) instead of an IADD3.
template <int R, bool IMADADDR>
__global__ void cellS(uint32_t* out, int iters, uint32_t seed, uint32_t gi2, uint32_t ie2, uint32_t de2, uint32_t one) {
  extern __shared__ uint32_t s_tab[];  // [5 residues * 50 pairs][32 lanes]
  for (int i = threadIdx.x; i < 5 * 50 * 32; i += blockDim.x) s_tab[i] = 0x00090009u + (uint32_t)(i % 7) * 0x00010001u - ((i % 3) ? 0x00140014u : 0u);
  __syncthreads();
  uint32_t rowterm[R], H[R], E[R];
  const uint32_t lane4 = (threadIdx.x & 31) * 4;
#pragma unroll
  for (int r = 0; r < R; r++) {
    rowterm[r] = (uint32_t)((r * 7 + threadIdx.x) % 5) * 50u * 128u;
    H[r] = 0;
    E[r] = 0;
  }
  uint32_t best = 0, pair = seed % 49, diag = 0, hu = 0, f = 0;
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(s_tab);
  for (int it = 0; it < iters; it++) {
    uint32_t hd = diag;
    diag = hu;
    pair = (pair * 5 + 3) % 49;
    const uint32_t coloff = base + pair * 128u + lane4;
#pragma unroll
    for (int r = 0; r < R; r++) {
      uint32_t s;
      if (IMADADDR) {
        uint32_t addr = rowterm[r] * one + coloff;
        asm("ld.shared.u32 %0, [%1];" : "=r"(s) : "r"(addr));
      } else {
        s = s_tab[(rowterm[r] + coloff - base) >> 2];  // plain C load: the compiler picks the address arithmetic
      }
      uint32_t hl = H[r];
      E[r] = __viaddmax_s16x2(E[r], ie2, hl);
      uint32_t d = __vadd2(hd, s);
      uint32_t tt = __viaddmax_s16x2_relu(E[r], gi2, d);
      uint32_t h = __viaddmax_s16x2(f, gi2, tt);
      f = __viaddmax_s16x2(f, de2, tt);
      hd = hl;
      H[r] = h;
      hu = h;
    }
    if ((it & 7) == 7) {
#pragma unroll
      for (int r = 0; r < R; r++) best = __vmaxs2(best, E[r]);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = best;
}

// B: biased cell.  A = H + G, B = E + 2G, C = F + 2G (G = -gap_init), profile scores s' = s + G >= 0:
//   B  = max(B + ie, A_left)              VIADDMNMX
//   X  = A_diag + s'                      IMAD (no carry between halves: both operands >= 0)
//   T' = max3(X, B, 2G)                   VIMNMX3
//   T  = T' - G                           IMAD
//   A  = max(C - G, T)                    VIADDMNMX   (off the chain)
//   C  = max(C + de, T)                   VIADDMNMX   (the only chain)
template <int R>
__global__ void cellB(uint32_t* out, int iters, uint32_t seed, uint32_t mG2, uint32_t ie2, uint32_t de2, uint32_t twoG2, uint32_t one) {
  uint32_t Plo[R], Phi[R], A[R], B[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    Plo[r] = 0x090a0100u + r * 0x01010101u + (threadIdx.x & 3);
    Phi[r] = 0x01060a09u ^ (r * 0x00010001u);
    A[r] = 0x001e001eu;
    B[r] = 0x001e001eu;
  }
  uint32_t best = 0, sel = 0x9180u | (seed & 0x0303u), diag = 0x001e001eu, au = 0x001e001eu, c = 0x001e001eu;
  for (int it = 0; it < iters; it++) {
    uint32_t ad = diag;
    diag = au;
    sel = ((sel + 0x0101u) & 0x0303u) | 0x9180u;  // the upper byte of each half replicates the sign of a non-negative byte: 0
#pragma unroll
    for (int r = 0; r < R; r++) {
      uint32_t s = prmt(Plo[r], Phi[r], sel);
      uint32_t al = A[r];
      B[r] = __viaddmax_s16x2(B[r], ie2, al);
      uint32_t x = ad * one + s;
      uint32_t t1 = __vimax3_s16x2(x, B[r], twoG2);
      uint32_t t = t1 * one + mG2;
      uint32_t a = __viaddmax_s16x2(c, mG2, t);
      c = __viaddmax_s16x2(c, de2, t);
      ad = al;
      A[r] = a;
      au = a;
    }
    if ((it & 7) == 7) {
#pragma unroll
      for (int r = 0; r < R; r++) best = __vmaxs2(best, B[r]);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = best;
}

// C: the same with the long chain (one VIADDMNMX fewer):
//   B = max(B + ie, A_left); X = A_diag + s'; M = max3(X, B, C); M' = max(M, 2G); A = M' - G; C = max(C + de, A)
template <int R>
__global__ void cellC(uint32_t* out, int iters, uint32_t seed, uint32_t mG2, uint32_t ie2, uint32_t de2, uint32_t twoG2, uint32_t one) {
  uint32_t Plo[R], Phi[R], A[R], B[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    Plo[r] = 0x090a0100u + r * 0x01010101u + (threadIdx.x & 3);
    Phi[r] = 0x01060a09u ^ (r * 0x00010001u);
    A[r] = 0x001e001eu;
    B[r] = 0x001e001eu;
  }
  uint32_t best = 0, sel = 0x9180u | (seed & 0x0303u), diag = 0x001e001eu, au = 0x001e001eu, c = 0x001e001eu;
  for (int it = 0; it < iters; it++) {
    uint32_t ad = diag;
    diag = au;
    sel = ((sel + 0x0101u) & 0x0303u) | 0x9180u;
#pragma unroll
    for (int r = 0; r < R; r++) {
      uint32_t s = prmt(Plo[r], Phi[r], sel);
      uint32_t al = A[r];
      B[r] = __viaddmax_s16x2(B[r], ie2, al);
      uint32_t x = ad * one + s;
      uint32_t m = __vimax3_s16x2(x, B[r], c);
      uint32_t m2 = __vmaxs2(m, twoG2);
      uint32_t a = m2 * one + mG2;
      c = __viaddmax_s16x2(c, de2, a);
      ad = al;
      A[r] = a;
      au = a;
    }
    if ((it & 7) == 7) {
#pragma unroll
      for (int r = 0; r < R; r++) best = __vmaxs2(best, B[r]);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = best;
}

// D: round-1 cell with the diagonal add moved to the FMA pipe (profile biased by +G, T computed on biased values):
//   H' = H + G >= G, E'' = E + 2G, F'' = F + 2G
//   E'' = max(E'' + ie, H'_left)      VIADDMNMX
//   X   = H'_diag + s'                IMAD          (= H_diag + s + 2G)
//   T'  = max3(X, E'', 2G)            VIMNMX3       (= T + 2G)
//   H'  = max(F'' , T') - G  ...  as  VIADDMNMX(F'', 0?, ..)   -- needs an add after the max: IMAD
// (equivalent to B; kept out)

static double time_ms(cudaEvent_t a, cudaEvent_t b) {
  float f = 0;
  cudaEventSynchronize(b);
  cudaEventElapsedTime(&f, a, b);
  return f;
}

template <int OP>
static void run_rate(uint32_t* d, int sms, double clk_hz, int warpsPerSM) {
  const int threads = 256, blocks = sms * (warpsPerSM * 32 / threads), iters = 20000;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  double best = 1e30;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(a);
    rate_kernel<OP><<<blocks, threads>>>(d, iters, 0x00010003u, 0x00050007u, 1u);
    cudaEventRecord(b);
    best = fmin(best, time_ms(a, b));
  }
  const double per = OP == MIX3_VIADDMAX_VADD2_IMAD ? 3.0 : (OP >= MIX_VIADDMAX_IMAD && OP != VMAX2_RELU) ? 2.0 : 1.0;
  const double instr = (double)blocks * (threads / 32) * 8.0 * iters * per;
  printf("%-34s warps/SM %2d  %8.3f ms  %.3e warp-instr/s  = %.3f warp-instr/clk/SM (%.3f per sub-partition)\n", kNames[OP], warpsPerSM, best,
         instr / (best * 1e-3), instr / (best * 1e-3) / clk_hz / sms, instr / (best * 1e-3) / clk_hz / sms / 4);
}

int main(int argc, char** argv) {
  const bool full = argc > 1 && argv[1][0] == 'f';
  const bool immOnly = argc > 1 && argv[1][0] == 'i';     // only the cells with immediate gap constants
  const bool cellsOnly = (argc > 1 && argv[1][0] == 'c') || immOnly;  // for ncu: one pass of the cell loops at 24 warps/SM
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double clk_hz = clk_khz * 1e3;
  const int sms = prop.multiProcessorCount;
  printf("%s: %d SMs, clock %.0f MHz (rates per clock assume this clock)\n", prop.name, sms, clk_hz / 1e6);
  uint32_t* d;
  cudaMalloc(&d, (size_t)sms * 64 * 256 * 4);
  for (int w : {16, 32}) {
    if (cellsOnly) break;
    if (full) {
      run_rate<VIADDMAX>(d, sms, clk_hz, w);
      run_rate<VIADDMAX_RELU>(d, sms, clk_hz, w);
      run_rate<VIMAX3>(d, sms, clk_hz, w);
      run_rate<VIMAX3_RELU>(d, sms, clk_hz, w);
      run_rate<VMAX2>(d, sms, clk_hz, w);
      run_rate<VADD2>(d, sms, clk_hz, w);
      run_rate<PRMT>(d, sms, clk_hz, w);
      run_rate<IADD>(d, sms, clk_hz, w);
      run_rate<IMAD>(d, sms, clk_hz, w);
      run_rate<MIX_VIADDMAX_IMAD>(d, sms, clk_hz, w);
      run_rate<MIX_VIADDMAX_PRMT>(d, sms, clk_hz, w);
      run_rate<MIX_VIADDMAX_VIMAX3>(d, sms, clk_hz, w);
      run_rate<MIX_VIMAX3_VMAX2>(d, sms, clk_hz, w);
      run_rate<MIX_VIADDMAX_VADD2>(d, sms, clk_hz, w);
      run_rate<MIX_VIADDMAX_IADD>(d, sms, clk_hz, w);
    }
    run_rate<VMAX2_RELU>(d, sms, clk_hz, w);
    run_rate<MIX_VADD2_IMAD>(d, sms, clk_hz, w);
    run_rate<MIX_VMAX2_VIADDMAX>(d, sms, clk_hz, w);
    run_rate<MIX_VMAX2_VADD2>(d, sms, clk_hz, w);
    run_rate<MIX3_VIADDMAX_VADD2_IMAD>(d, sms, clk_hz, w);
    run_rate<MIX_VMAX2_IMAD>(d, sms, clk_hz, w);
    run_rate<MIX_PRMT_VADD2>(d, sms, clk_hz, w);
  }
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  const uint32_t gi2 = 0xffe2ffe2u, ie2 = 0xfffafffau, de2 = 0xfffbfffbu, twoG2 = 0x003c003cu;
  for (int warps : {16, 20, 24, 32}) {
    if (cellsOnly && !immOnly && warps != 24) continue;
    if (immOnly && warps != 16) continue;
    const int threads = 128, blocks = sms * (warps * 32 / threads), iters = 20000;
    auto report = [&](const char* name, int R, double ms) {
      printf("%-28s R=%2d warps/SM %2d  %8.3f ms  %8.1f GCUPS\n", name, R, warps, ms, (double)blocks * threads * R * (double)iters * 2.0 / (ms * 1e-3) / 1e9);
    };
    double best;
#define RUN(NAME, R, CALL)                 \
  best = 1e30;                             \
  for (int rep = 0; rep < 4; rep++) {      \
    cudaEventRecord(a);                    \
    CALL;                                  \
    cudaEventRecord(b);                    \
    best = fmin(best, time_ms(a, b));      \
  }                                        \
  report(NAME, R, best);
    RUN("cell Ai (A, immediates)", 16, (cellAi<16><<<blocks, threads>>>(d, iters, rep)));
    RUN("cell Li (VIMNMX3, immediates)", 16, (cellLi<16, false><<<blocks, threads>>>(d, iters, rep, 1u)));
    RUN("cell Lpi (VIMNMX3 forced, imm.)", 16, (cellLi<16, true><<<blocks, threads>>>(d, iters, rep, 1u)));
    RUN("cell Lpi (VIMNMX3 forced, imm.)", 12, (cellLi<12, true><<<blocks, threads>>>(d, iters, rep, 1u)));
    if (immOnly) continue;
    RUN("cell A (round 1)", 8, (cellA<8><<<blocks, threads>>>(d, iters, rep, gi2, ie2, de2)));
    RUN("cell A (round 1)", 16, (cellA<16><<<blocks, threads>>>(d, iters, rep, gi2, ie2, de2)));
    RUN("cell A1 (E split)", 16, (cellA1<16><<<blocks, threads>>>(d, iters, rep, gi2, ie2, de2)));
    RUN("cell A1h (H split)", 16, (cellA1h<16><<<blocks, threads>>>(d, iters, rep, gi2, ie2, de2)));
    RUN("cell A2 (E,H split)", 16, (cellA2<16><<<blocks, threads>>>(d, iters, rep, gi2, ie2, de2)));
    RUN("cell A2b (H,T split)", 16, (cellA2b<16><<<blocks, threads>>>(d, iters, rep, gi2, ie2, de2)));
    RUN("cell A3 (E,H,T split)", 16, (cellA3<16><<<blocks, threads>>>(d, iters, rep, gi2, ie2, de2)));
    RUN("cell A2 (E,H split)", 8, (cellA2<8><<<blocks, threads>>>(d, iters, rep, gi2, ie2, de2)));
    RUN("cell L (VIMNMX3, long chain)", 16, (cellL<16><<<blocks, threads>>>(d, iters, rep, gi2, ie2, de2)));
    RUN("cell Lp (VIMNMX3 forced)", 16, (cellLp<16><<<blocks, threads>>>(d, iters, rep, gi2, ie2, de2, 1u)));
    RUN("cell Lp (VIMNMX3 forced)", 12, (cellLp<12><<<blocks, threads>>>(d, iters, rep, gi2, ie2, de2, 1u)));
    RUN("cell L (VIMNMX3, long chain)", 8, (cellL<8><<<blocks, threads>>>(d, iters, rep, gi2, ie2, de2)));
    RUN("cell L (VIMNMX3, long chain)", 12, (cellL<12><<<blocks, threads>>>(d, iters, rep, gi2, ie2, de2)));
    RUN("cell S (LDS scores, IADD address)", 16, (cellS<16, false><<<blocks, threads, 32000>>>(d, iters, rep, gi2, ie2, de2, 1u)));
    RUN("cell S (LDS scores, IMAD address)", 16, (cellS<16, true><<<blocks, threads, 32000>>>(d, iters, rep, gi2, ie2, de2, 1u)));
    RUN("cell S (LDS scores, IADD address)", 12, (cellS<12, false><<<blocks, threads, 32000>>>(d, iters, rep, gi2, ie2, de2, 1u)));
    RUN("cell B (biased, short chain)", 8, (cellB<8><<<blocks, threads>>>(d, iters, rep, gi2, ie2, de2, twoG2, 1u)));
    RUN("cell B (biased, short chain)", 16, (cellB<16><<<blocks, threads>>>(d, iters, rep, gi2, ie2, de2, twoG2, 1u)));
    RUN("cell C (biased, long chain)", 8, (cellC<8><<<blocks, threads>>>(d, iters, rep, gi2, ie2, de2, twoG2, 1u)));
    RUN("cell C (biased, long chain)", 16, (cellC<16><<<blocks, threads>>>(d, iters, rep, gi2, ie2, de2, twoG2, 1u)));
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  cudaFree(d);
  return e != cudaSuccess;
}
