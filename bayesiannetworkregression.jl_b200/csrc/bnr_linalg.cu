// Dense FP64 linear algebra of the gamma draw (update_gamma!, src/gibbs.jl:420-438), batched over chains:
//   G_c = X diag(S_c) X' + I          -> k_gram_syrk : FP64 tensor-core (DMMA m8n8k4) SYRK, ring of TMA tensor-map boxes + mbarriers
//   G_c = L_c L_c', w = L_c^-1 rhs    -> blocked left-looking Cholesky: k_augment / k_chol_update (DMMA) / k_potf2_inv /
//                                        k_trsm_dmma (DMMA) / k_small_tile; the forward solve rides along as a bordering row
//   a4  = L_c^-T w                    -> k_bwd_stream (the factor streamed once through a ring of TMA boxes)
//   X v, X' a4 (all chains at once)   -> k_xmma (tall-skinny DMMA GEMM, deterministic split-K)
// tcgen05 has no FP64 kind, so the Blackwell tensor path for this contraction is the warp-level DMMA.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include "bnr_engine.cuh"
#include "bnr_kernels.h"

namespace bnr {

// ------------------------------------------------------------------------------------------------------------
// small PTX helpers
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// -x by flipping the sign bit (an integer op: the FP64 pipe is what the DMMA loop is bound by)
__device__ __forceinline__ double neg_bits(double x) {
  return __longlong_as_double(__double_as_longlong(x) ^ (long long)0x8000000000000000ull);
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}

// ------------------------------------------------------------------------------------------------------------
// mbarrier / bulk-copy (TMA engine) helpers
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 2-D tiled TMA load (SASS: UTMALDG): one request moves a whole (box_inner rows) x (16 k) operand tile; the box is
// 4 rows wider than the tile, which IS the shared-memory padding (row pitch == 4 mod 16 doubles: conflict-free fragments)
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, int x, int y, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
                   smem_u32(smem_dst)), "l"((unsigned long long)tm), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
// contiguous global -> shared copy by the TMA engine, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------------------------------------------------
// SYRK on FP64 tensor cores.
//   MODE 0:  C_c[i][j] = sum_k A[i][k] s_c[k] A[j][k] + (i==j)      A = X (shared by all chains), K = qp
//   MODE 1:  C_c[i][j] -= sum_{k < nk*16} L_c[i][k] L_c[j][k]        left-looking Cholesky update of block column J
//   MODE 2:  C_c[i][j]  = sum_{k < 128} C_c[i][k] Linv_c[j][k]       panel solve with the inverted diagonal block
// The operand is "k-major": element (row, k) at A[row + ld*k] (rows contiguous) - exactly how X (column-major
// n x q) and a column panel of the column-major G are stored, so a 128-row x 16-k tile is sixteen contiguous
// 1 KB rows = one box of a 2-D tensor map {rows, k}.  A dedicated producer warp loads one box per operand and stage
// into a 4-stage shared-memory ring (cp.async.bulk.tensor.2d, completion on "full" mbarriers); the 8 consumer warps
// never execute a block-wide barrier in the main loop - they wait on "full", issue DMMA m8n8k4, and release the
// stage on "empty".
// CTA tile 128 x 128, k-step 16.  The MMA "M" dimension runs along j (columns of C) and "N" along i (rows of C), so
// every accumulator pair is two consecutive rows of one column of the column-major C -> 16-byte stores.  Strictly
// lower tiles use a column-strip warp layout (syrk_strip_tile: 2 column fragments x all row fragments per warp),
// diagonal tiles compute only their lower triangle, balanced over the warps (syrk_diag_tile).  Shared rows are
// padded to 132 doubles (== 4 mod 16) which makes the (row, k) fragment loads bank-conflict free.
// The per-chain scale s_c[k] is applied to the operand with the fewest fragments per warp (2): DMUL shares the
// FP64 pipe with DMMA.  Work that cannot contribute is skipped: the symmetric half of diagonal tiles and row
// fragments that lie entirely in the zero padding beyond n.  The boxes are 132 rows wide (4 more than the tile):
// the box pitch is the padded row stride, at the price of 3 % more L2 traffic.
// NOTE the ring is written by the TMA engine behind the compiler's back: the consumer-side pointers into it must NOT
// be __restrict__ (with a compile-time trip count the compiler then reuses the fragments of a stage's previous
// occupant); the "memory" clobber of mbar_wait is what orders the fragment loads after the hand-shake.
// grid = (C, k-splits, tiles), block = 384 (2 consumer warpgroups + 1 producer warpgroup, registers re-balanced with
// setmaxnreg: 232 per consumer thread, 40 per producer thread); dynamic smem = SYRK_SMEM.
// ------------------------------------------------------------------------------------------------------------
constexpr int SY_BT = 128;        // tile edge
#ifndef BNR_SY_BK
#define BNR_SY_BK 16
#endif
#ifndef BNR_SY_STAGES
#define BNR_SY_STAGES 4
#endif
constexpr int SY_BK = BNR_SY_BK;  // k-step
constexpr int SY_LDS = 132;       // smem row stride in doubles (== 4 mod 16 -> conflict-free fragment loads)
constexpr int SY_STAGES = BNR_SY_STAGES;
constexpr int SYRK_MAX_SPLITS = 8;
constexpr int SY_STAGE_DBL = 2 * SY_BK * SY_LDS + SY_BK;   // two operand tiles + 16 scales
constexpr int SY_THREADS = 384;   // 2 consumer warpgroups + 1 producer warpgroup (one active lane)
constexpr int SY_PRODUCER_REGS = 40, SY_CONSUMER_REGS = 232;   // 128*40 + 256*232 = 64512 = 384*168
constexpr size_t SYRK_SMEM = (size_t)SY_STAGES * SY_STAGE_DBL * sizeof(double) + 2 * SY_STAGES * sizeof(unsigned long long);

// Diagonal tiles: only the lower triangle of the 128 x 128 tile is needed.  Seen as 16 x 16 fragments of 8 x 8,
// fragment column j holds 16 - j useful fragments; warp W takes columns W and 15 - W (17 fragments, the same for
// every warp), so the symmetric half is skipped AND the four SM sub-partitions stay evenly loaded.  W is a template
// parameter: every loop bound and register index is static, no predicated DMMA.
// acc[t], t < 16-W : fragment (row-fragment W+t, column-fragment W);  t >= 16-W : (row-fragment t-1, column 15-W).
template <int MODE, int W>
__device__ __forceinline__ void syrk_diag_frags(const double* sj, const double* ss, int k4,
                                                int lk, int lr, double (&a2)[2], double (&bfr)[16 - W]) {
  const int kr = k4 * 4 + lk;
  const double* row = sj + kr * SY_LDS + lr;
  // the scale rides on the two column-operand fragments (2 DMULs) rather than on the 16 - W row-operand fragments:
  // DMUL shares the FP64 pipe with DMMA, so every multiply saved is tensor throughput gained
  if (MODE == 0) {
    const double sk = ss[kr];
    a2[0] = row[W * 8] * sk;
    a2[1] = row[(15 - W) * 8] * sk;
  } else if (MODE == 1) {
    // C -= A A': the accumulators start from C (syrk_diag_tile) and the column operand enters negated
    a2[0] = neg_bits(row[W * 8]);
    a2[1] = neg_bits(row[(15 - W) * 8]);
  } else {
    a2[0] = row[W * 8];
    a2[1] = row[(15 - W) * 8];
  }
#pragma unroll
  for (int i = 0; i < 16 - W; ++i) bfr[i] = row[(W + i) * 8];
}

template <int MODE, int W>
__device__ __forceinline__ void syrk_diag_tile(const double* smem, unsigned long long* full,
                                               unsigned long long* empty, int nk, int lk, int lr, int lane,
                                               double* __restrict__ Cc, int np, int i0, double diag_add) {
  constexpr int NA = 16 - W;        // fragments of column W
  double acc[17][2];
  if (MODE == 1) {
    // MODE 1 updates C in place.  Its 17 fragments are loaded NOW, as independent 16-byte loads that overlap the
    // pipeline fill: the old load-subtract-store epilogue serialised them (a later load may alias an earlier store),
    // which cost ~25 us per CTA once G (512 MB) lives in HBM -- half of all stall samples of k_chol_update.
#pragma unroll
    for (int t = 0; t < 17; ++t) {
      const int jf = (t < NA) ? W : 15 - W;
      const int ifr = (t < NA) ? W + t : t - 1;
      const double2 v = *reinterpret_cast<const double2*>(Cc + (size_t)(i0 + jf * 8 + lr) * np + i0 + ifr * 8 + 2 * lk);
      acc[t][0] = v.x; acc[t][1] = v.y;
    }
  } else {
#pragma unroll
    for (int t = 0; t < 17; ++t) acc[t][0] = acc[t][1] = 0.0;
  }
  double a2[2][2], bfr[2][NA];
  mbar_wait(&full[0], 0);
  const double* sj = smem;
  syrk_diag_frags<MODE, W>(sj, sj + 2 * SY_BK * SY_LDS, 0, lk, lr, a2[0], bfr[0]);
  for (int kt = 0; kt < nk; ++kt) {
#pragma unroll
    for (int k4 = 0; k4 < SY_BK / 4; ++k4) {
      const int cur = k4 & 1, nxt = cur ^ 1;
      if (k4 + 1 < SY_BK / 4) {
        syrk_diag_frags<MODE, W>(sj, sj + 2 * SY_BK * SY_LDS, k4 + 1, lk, lr, a2[nxt], bfr[nxt]);
      } else if (kt + 1 < nk) {
        mbar_wait(&full[(kt + 1) % SY_STAGES], ((kt + 1) / SY_STAGES) & 1);
        const double* nj = smem + (size_t)((kt + 1) % SY_STAGES) * SY_STAGE_DBL;
        syrk_diag_frags<MODE, W>(nj, nj + 2 * SY_BK * SY_LDS, 0, lk, lr, a2[nxt], bfr[nxt]);
      }
#pragma unroll
      for (int t = 0; t < NA; ++t) dmma884(acc[t][0], acc[t][1], a2[cur][0], bfr[cur][t]);
#pragma unroll
      for (int t = NA; t < 17; ++t) dmma884(acc[t][0], acc[t][1], a2[cur][1], bfr[cur][t - 1 - W]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[kt % SY_STAGES]);
    sj = smem + (size_t)((kt + 1) % SY_STAGES) * SY_STAGE_DBL;
  }
#pragma unroll
  for (int t = 0; t < 17; ++t) {
    const int jf = (t < NA) ? W : 15 - W;
    const int ifr = (t < NA) ? W + t : t - 1;
    const int j = i0 + jf * 8 + lr;                   // diagonal tile: j0 == i0
    const int i = i0 + ifr * 8 + 2 * lk;
    double2* p = reinterpret_cast<double2*>(Cc + (size_t)j * np + i);
    double2 v;
    if (MODE == 0) {
      v.x = acc[t][0] + (i == j ? diag_add : 0.0);
      v.y = acc[t][1] + (i + 1 == j ? diag_add : 0.0);
    } else {
      v.x = acc[t][0];               // MODE 1: C - A A' accumulated in place
      v.y = acc[t][1];
    }
    *p = v;
  }
}

// Off-diagonal tiles, "column-strip" warp layout: warp w owns the two 8-column fragments 2w, 2w+1 of the tile
// (columns j0 + 16w .. j0 + 16w + 15 of C) and ALL NF 8-row fragments of the i operand, i.e. 2 x NF DMMAs per k4-step.
// Compared with a 64 x 32 warp tile this (a) halves the DMULs: the per-chain scale rides on the 2 column-operand
// fragments, (b) lets a tile of the last block row stop at the last fragment that holds a valid row (NF < 16):
// the zero padding of n up to a multiple of 128 costs no tensor work.
template <int MODE, int NF>
__device__ __forceinline__ void syrk_strip_frags(const double* sj, const double* si, const double* ss, int k4,
                                                 int f0, int f1, int lk, int lr, int ldi,
                                                 double (&af)[2], double (&bf)[NF]) {
  const int kr = k4 * 4 + lk;
  const double* rj = sj + kr * SY_LDS + lr;
  const double* ri = si + kr * ldi + lr;
  if (MODE == 0) {
    const double sk = ss[kr];
    af[0] = rj[f0 * 8] * sk;
    af[1] = rj[f1 * 8] * sk;
  } else if (MODE == 1) {
    af[0] = neg_bits(rj[f0 * 8]);    // C -= A_i A_j': accumulators start from C, the 2-fragment operand enters negated
    af[1] = neg_bits(rj[f1 * 8]);
  } else {
    af[0] = rj[f0 * 8];
    af[1] = rj[f1 * 8];
  }
#pragma unroll
  for (int nf = 0; nf < NF; ++nf) bf[nf] = ri[nf * 8];
}

// MODE 2 (panel solve, the j operand is the lower-triangular inverse of the diagonal block): output column j' only
// needs k <= j', so warp w takes the column fragments w and 15 - w (instead of 2w, 2w + 1) and stops feeding each of
// them once k passes its last column: 2w + 2 and 32 - 2w k4-steps -- 34 of 64 for every warp, balanced.
template <int MODE, int NF>
__device__ __forceinline__ void syrk_strip_tile(const double* smem, unsigned long long* full,
                                                unsigned long long* empty, int nk, int warp, int lk, int lr, int lane,
                                                double* __restrict__ Cc, int np, int i0, int j0, int ldi) {
  double acc[2][NF][2];
  const int f0 = (MODE == 2) ? warp : 2 * warp, f1 = (MODE == 2) ? 15 - warp : 2 * warp + 1;
  if (MODE == 1) {
    // in-place update: the tile's current values are the initial accumulators (see syrk_diag_tile)
#pragma unroll
    for (int mf = 0; mf < 2; ++mf) {
      const int j = j0 + (mf == 0 ? f0 : f1) * 8 + lr;
#pragma unroll
      for (int nf = 0; nf < NF; ++nf) {
        const double2 v = *reinterpret_cast<const double2*>(Cc + (size_t)j * np + i0 + nf * 8 + 2 * lk);
        acc[mf][nf][0] = v.x; acc[mf][nf][1] = v.y;
      }
    }
  } else {
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < NF; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
  }
  const int lim0 = 2 * f0 + 2, lim1 = 2 * f1 + 2;       // MODE 2: k4-steps that can still reach the fragment
  double af[2][2], bf[2][NF];
  mbar_wait(&full[0], 0);
  const double* sj = smem;
  syrk_strip_frags<MODE, NF>(sj, sj + SY_BK * SY_LDS, sj + 2 * SY_BK * SY_LDS, 0, f0, f1, lk, lr, ldi, af[0], bf[0]);
  for (int kt = 0; kt < nk; ++kt) {
#pragma unroll
    for (int k4 = 0; k4 < SY_BK / 4; ++k4) {
      const int cur = k4 & 1, nxt = cur ^ 1;
      if (k4 + 1 < SY_BK / 4) {
        syrk_strip_frags<MODE, NF>(sj, sj + SY_BK * SY_LDS, sj + 2 * SY_BK * SY_LDS, k4 + 1, f0, f1, lk, lr, ldi, af[nxt], bf[nxt]);
      } else if (kt + 1 < nk) {
        mbar_wait(&full[(kt + 1) % SY_STAGES], ((kt + 1) / SY_STAGES) & 1);
        const double* nj = smem + (size_t)((kt + 1) % SY_STAGES) * SY_STAGE_DBL;
        syrk_strip_frags<MODE, NF>(nj, nj + SY_BK * SY_LDS, nj + 2 * SY_BK * SY_LDS, 0, f0, f1, lk, lr, ldi, af[nxt], bf[nxt]);
      }
      if (MODE == 2) {
        const int k4g = kt * (SY_BK / 4) + k4;
        if (k4g < lim0) {
#pragma unroll
          for (int nf = 0; nf < NF; ++nf) dmma884(acc[0][nf][0], acc[0][nf][1], af[cur][0], bf[cur][nf]);
        }
        if (k4g < lim1) {
#pragma unroll
          for (int nf = 0; nf < NF; ++nf) dmma884(acc[1][nf][0], acc[1][nf][1], af[cur][1], bf[cur][nf]);
        }
      } else {
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) {
          dmma884(acc[0][nf][0], acc[0][nf][1], af[cur][0], bf[cur][nf]);
          dmma884(acc[1][nf][0], acc[1][nf][1], af[cur][1], bf[cur][nf]);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[kt % SY_STAGES]);
    sj = smem + (size_t)((kt + 1) % SY_STAGES) * SY_STAGE_DBL;
  }
#pragma unroll
  for (int mf = 0; mf < 2; ++mf) {
    const int j = j0 + (mf == 0 ? f0 : f1) * 8 + lr;
#pragma unroll
    for (int nf = 0; nf < NF; ++nf) {
      const int i = i0 + nf * 8 + 2 * lk;
      double2* p = reinterpret_cast<double2*>(Cc + (size_t)j * np + i);
      double2 v;
      v.x = acc[mf][nf][0];          // MODE 1: C - A_i A_j' accumulated in place
      v.y = acc[mf][nf][1];
      *p = v;
    }
  }
}

// Operands arrive as tensor-map boxes: the j operand through tmj at (row jx, k-row jy0 + 16 kt), the i operand through
// tmi at (row i0, k-row iy0 + 16 kt); jy0 / iy0 carry the chain, the first contributing column and the k-split offset.
template <int MODE>
__device__ __forceinline__ void
syrk_body(const CUtensorMap* tmj, const CUtensorMap* tmi, int jy0, int iy0, const double* __restrict__ scale,
          size_t scale_stride, double* __restrict__ Cm, size_t c_chain_stride, int np, int nvalid, int nk, int origin,
          double diag_add, int part = 0, int nstrip = 1, int tile_index = 0) {
  extern __shared__ __align__(128) double smem[];
  unsigned long long* full = reinterpret_cast<unsigned long long*>(smem + (size_t)SY_STAGES * SY_STAGE_DBL);
  unsigned long long* empty = full + SY_STAGES;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = blockIdx.x;          // chains vary fastest: the 64 CTAs that share one X tile run back to back (L2 reuse)
  int ib, jb;
  if (MODE == 0) {
    // tile_index enumerates the lower-triangular tiles by decreasing cost, so that the tail of the grid is made of
    // cheap CTAs: the strictly-lower (full-cost) tiles, the diagonal (half-cost) ones, and last the strictly-lower
    // tiles of a SHORT last row block (<= 32 valid rows: a quarter of the DMMAs)
    const int T = np / SY_BT;
    const bool short_last = T > 1 && nvalid - (T - 1) * SY_BT <= 32;
    const int noff = short_last ? (T - 1) * (T - 2) / 2 : T * (T - 1) / 2;
#ifdef SYRK_LAB_ONLY_DIAG
    const int t = tile_index + noff;
#else
    const int t = tile_index;
#endif
    if (t < noff) {
      ib = (int)((1.0 + sqrt(1.0 + 8.0 * t)) * 0.5);
      while (ib * (ib - 1) / 2 > t) --ib;
      while ((ib + 1) * ib / 2 <= t) ++ib;
      jb = t - ib * (ib - 1) / 2;
    } else if (t < noff + T) {
      ib = jb = t - noff;
    } else {
      ib = T - 1;
      jb = t - noff - T;
    }
  } else {
    // MODE 1 (Cholesky update of block column jb = origin) and MODE 2 (panel solve of block column jb): the CTAs of a
    // launch cover the row blocks part, part + 1, ... of that column (part >= jb; part == jb includes the diagonal tile)
    // When the launch has too few tiles to occupy the GPU (few chains: the latency-bound regime) every tile is cut
    // into nstrip (2 or 4) row strips of 128 / nstrip rows, one CTA each: the same fragments in the same k order
    // (bit-identical results), a quarter of the DMMA chain per CTA.  A diagonal tile is then computed like any other
    // (its upper triangle is written too; nobody reads it).
    jb = origin;
    ib = part + tile_index / nstrip;
  }
  const int strip = (MODE == 0) ? 0 : tile_index % nstrip;
  const int rows_i = SY_BT / nstrip;
  const int i0 = ib * SY_BT + strip * rows_i, j0 = jb * SY_BT;
  const bool diag = (ib == jb) && nstrip == 1;
  const double* sc = (MODE == 0) ? scale + (size_t)c * scale_stride : nullptr;
  // operand sources (k-major: element (row, k) at base[row + ld * k]).  MODE 0 / 1: both operands are row blocks of
  // the same matrix.  MODE 2: the i operand is block column J of C itself (rows i0.., k = the 128 columns of the
  // panel), the j operand is the 128 x 128 inverse of the diagonal block (ld = 128, rows 0..127).
  const int jx = (MODE == 2) ? 0 : j0;
  const int ldi = (MODE == 0) ? SY_LDS : rows_i + 4;   // row pitch of the i operand in shared memory = its box width

  if (tid == 0) {
    for (int s = 0; s < SY_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if (warp >= 8) {
    // ---------------- producer warpgroup: hands its registers to the consumers, one lane drives the TMA engine ----
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(SY_PRODUCER_REGS));
    if (warp == 8 && lane == 0) {
      // three requests per stage (two operand boxes + the 16 scales); the first version issued one 1 KB bulk copy per
      // k-row and operand -- 33 requests, ~1 us per stage at the TMA engine's ~30 ns per request, which bounded every
      // tile with less than ~1 us of DMMA work per stage (row strips, tiles of a short last row block)
      const unsigned bytes = SY_BK * SY_LDS * 8u + (diag ? 0u : SY_BK * (unsigned)ldi * 8u) + (MODE == 0 ? SY_BK * 8u : 0u);
      for (int kt = 0; kt < nk; ++kt) {
        const int stage = kt % SY_STAGES;
        const unsigned ph = (kt / SY_STAGES) & 1;
        mbar_wait(&empty[stage], ph ^ 1);
        double* sj = smem + (size_t)stage * SY_STAGE_DBL;
        double* si = sj + SY_BK * SY_LDS;
        mbar_expect_tx(&full[stage], bytes);
        tma_load_2d(sj, tmj, jx, jy0 + kt * SY_BK, &full[stage]);
        if (!diag) tma_load_2d(si, tmi, i0, iy0 + kt * SY_BK, &full[stage]);
        if (MODE == 0) bulk_g2s(si + SY_BK * SY_LDS, sc + kt * SY_BK, SY_BK * 8, &full[stage]);
      }
    }
    return;
  }

  // ---------------- consumers ----------------
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(SY_CONSUMER_REGS));
  const int lk = lane & 3, lr = lane >> 2;
  double* Cc = Cm + (size_t)c * c_chain_stride;

  if (diag) {
    switch (warp) {
      case 0: syrk_diag_tile<MODE, 0>(smem, full, empty, nk, lk, lr, lane, Cc, np, i0, diag_add); break;
      case 1: syrk_diag_tile<MODE, 1>(smem, full, empty, nk, lk, lr, lane, Cc, np, i0, diag_add); break;
      case 2: syrk_diag_tile<MODE, 2>(smem, full, empty, nk, lk, lr, lane, Cc, np, i0, diag_add); break;
      case 3: syrk_diag_tile<MODE, 3>(smem, full, empty, nk, lk, lr, lane, Cc, np, i0, diag_add); break;
      case 4: syrk_diag_tile<MODE, 4>(smem, full, empty, nk, lk, lr, lane, Cc, np, i0, diag_add); break;
      case 5: syrk_diag_tile<MODE, 5>(smem, full, empty, nk, lk, lr, lane, Cc, np, i0, diag_add); break;
      case 6: syrk_diag_tile<MODE, 6>(smem, full, empty, nk, lk, lr, lane, Cc, np, i0, diag_add); break;
      default: syrk_diag_tile<MODE, 7>(smem, full, empty, nk, lk, lr, lane, Cc, np, i0, diag_add); break;
    }
    return;
  }

  // strictly-lower tile: only the 8-row fragments of the i operand that hold a valid row are computed and stored.
  // Rows >= nvalid are zero padding: their entries of C are exactly zero, were zero-initialised at allocation and
  // are never written by any kernel, so skipping them changes nothing.
  int nfv = (nvalid - i0 + 7) / 8;
  nfv = nfv > rows_i / 8 ? rows_i / 8 : nfv;
  if (nfv <= 0) {
    // the whole row block is padding: keep the stage hand-shake alive and leave
    for (int kt = 0; kt < nk; ++kt) {
      mbar_wait(&full[kt % SY_STAGES], (kt / SY_STAGES) & 1);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[kt % SY_STAGES]);
    }
    return;
  }
  // (at least 4 row fragments: fewer template instances; the extra fragments are zero-padding rows of the last block)
  if (nfv < 4) nfv = 4;
  switch (nfv) {
#define BNR_STRIP_CASE(NF) case NF: syrk_strip_tile<MODE, NF>(smem, full, empty, nk, warp, lk, lr, lane, Cc, np, i0, j0, ldi); break;
    BNR_STRIP_CASE(4) BNR_STRIP_CASE(5) BNR_STRIP_CASE(6)
    BNR_STRIP_CASE(7) BNR_STRIP_CASE(8) BNR_STRIP_CASE(9) BNR_STRIP_CASE(10) BNR_STRIP_CASE(11) BNR_STRIP_CASE(12)
    BNR_STRIP_CASE(13) BNR_STRIP_CASE(14) BNR_STRIP_CASE(15)
#undef BNR_STRIP_CASE
    default: syrk_strip_tile<MODE, 16>(smem, full, empty, nk, warp, lk, lr, lane, Cc, np, i0, j0, ldi); break;
  }
}

// gridDim.z > 1 splits the contraction: split z handles k-steps [z * nk_split, ...) and writes its partial tile to
// ws[c][z - 1] (split 0 writes G itself, identity included); k_syrk_splitk_reduce adds them in a fixed order.  Used
// when chains x tiles cannot fill the SMs (few chains, large q: BASELINE config 4).
__global__ void __launch_bounds__(SY_THREADS, 1)
k_gram_syrk(const __grid_constant__ CUtensorMap tmX, const double* __restrict__ scale, size_t scale_stride,
            double* __restrict__ G, size_t g_chain_stride, int np, int nvalid, int nk, double diag_add,
            int nk_split, double* __restrict__ ws, int ws_cap) {
  // grid = (chains, k-splits, tiles): the hardware hands out CTAs x-fastest, z-slowest, so all splits of the expensive
  // tiles start before any cheap tile (with the split as the slowest index the last split's full-cost tiles started
  // last and the grid ended in a ~340 us tail of a few busy SMs)
  const int z = blockIdx.y, t = blockIdx.z;
  if (z == 0) {
    syrk_body<0>(&tmX, &tmX, 0, 0, scale, scale_stride, G, g_chain_stride, np, nvalid, nk < nk_split ? nk : nk_split, 0, diag_add,
                 0, 1, t);
  } else {
    const int kt0 = z * nk_split;
    const int nkl = (nk - kt0) < nk_split ? (nk - kt0) : nk_split;
    syrk_body<0>(&tmX, &tmX, kt0 * SY_BK, kt0 * SY_BK, scale + (size_t)kt0 * SY_BK, scale_stride,
                 ws + (size_t)(z - 1) * g_chain_stride, (size_t)ws_cap * g_chain_stride, np, nvalid, nkl, 0, 0.0, 0, 1, t);
  }
}

// G_c (lower tiles) += sum_z ws[c][z] (fixed order).  grid = (lower tiles * 8, C), block = 256: 2048 entries per block
__global__ void __launch_bounds__(256) k_syrk_splitk_reduce(double* __restrict__ G, size_t g_chain_stride, int np,
                                                            const double* __restrict__ ws, int ws_cap, int nz) {
  const int c = blockIdx.y, piece = blockIdx.x & 7;
  int t = blockIdx.x >> 3, ib = 0;
  while ((ib + 1) * (ib + 2) / 2 <= t) ++ib;
  const int jb = t - ib * (ib + 1) / 2;
  double* Gc = G + (size_t)c * g_chain_stride;
  const double* wc = ws + (size_t)c * ws_cap * g_chain_stride;
  for (int id = piece * 2048 + threadIdx.x; id < (piece + 1) * 2048; id += 256) {
    const int i = ib * SY_BT + (id & (SY_BT - 1)), j = jb * SY_BT + (id >> 7);
    const size_t o = (size_t)j * np + i;
    double v = Gc[o];
    for (int z = 0; z < nz; ++z) v += wc[(size_t)z * g_chain_stride + o];
    Gc[o] = v;
  }
}

// G[ib, jb] -= sum_k P[ib, k] P[jb, k]' for the row blocks ib = ib_first, ib_first + 1, ... (grid.y of them) of block
// column jb; P points at the first of the nk * 16 columns of the factor that contribute (the caller offsets it), so a
// launch applies any contiguous range of panels.
// (tmJ: the chain-stacked matrix {rows, columns x chains} with 132-row boxes; tmI: the same with boxes as wide as a row
//  strip + 4; kcol0: first contributing column)
__global__ void __launch_bounds__(SY_THREADS, 1)
k_chol_update(const __grid_constant__ CUtensorMap tmJ, const __grid_constant__ CUtensorMap tmI, int kcol0,
              double* __restrict__ G, size_t chain_stride, int np, int nvalid, int nk, int jb, int ib_first, int nstrip) {
  const int y0 = (int)blockIdx.x * np + kcol0;
  syrk_body<1>(&tmJ, &tmI, y0, y0, nullptr, 0, G, chain_stride, np, nvalid, nk, jb, 0.0, ib_first, nstrip, (int)blockIdx.y);
}

// panel solve on the tensor cores: G[I, J] <- G[I, J] Linv_J' for the row blocks I = ib_first, ... (in place: a CTA
// has consumed its whole tile through the ring before the first store).  tmL: the stacked inverses {128, 128 x panels x
// chains}; tmI as above.
__global__ void __launch_bounds__(SY_THREADS, 1)
k_trsm_dmma(const __grid_constant__ CUtensorMap tmL, const __grid_constant__ CUtensorMap tmI, double* __restrict__ G,
            size_t chain_stride, int np, int nvalid, int jb, int ib_first, int nstrip) {
  const int T = np / SY_BT;
  syrk_body<2>(&tmL, &tmI, ((int)blockIdx.x * T + jb) * SY_BT, (int)blockIdx.x * np + jb * SY_BT, nullptr, 0, G, chain_stride,
               np, nvalid, SY_BT / SY_BK, jb, 0.0, ib_first, nstrip, (int)blockIdx.y);
}

// ------------------------------------------------------------------------------------------------------------
// Blocked left-looking Cholesky, block size 128, of every chain's gdim x gdim matrix, with BOTH triangular solves
// organised around it:
//   * forward solve for free: the right-hand side r is stored as row m (the first padding row, m = n or q) of the
//     matrix itself, with the diagonal entry d = 1 + |r|^2 / lmin (k_augment; lmin <= lambda_min(G)).  The bordered
//     matrix [[G, r], [r', d]] is positive definite (r' G^-1 r <= |r|^2 / lambda_min < d), and row m of its Cholesky
//     factor IS w' = (L^-1 r)': the DMMA update / panel-solve kernels below carry the solve along as one more row of
//     the last row block.  (The corner entry sqrt(d - |w|^2) is never used.)
//   * for J = 0 .. T-1:
//       k_chol_update (DMMA): G[J.., J] -= L[J.., 0:J] L[J, 0:J]'
//       k_potf2_inv   one CTA per chain: the 128 x 128 diagonal block is factored in shared memory in 32-wide steps,
//                     software-pipelined around the pivot chain (one warp factors 32 x 32 in registers and publishes
//                     its columns; the rows below, the diagonal-block inverses, the trailing updates, the bordered
//                     inverse and the stores all run next to it, see the kernel) -> factor in place, Linv[c][J]
//       k_trsm_dmma   (DMMA): G[I, J] <- G[I, J] Linv_J' for I > J -- the panel solve is a GEMM
//   * backward solve L' x = w (+ z): k_bwd_stream, one CTA per chain streams the factor once (tensor-map boxes of
//     128 rows x 32 columns into a shared-memory ring), every step is a block mat-vec; the diagonal blocks use Linv.
// ------------------------------------------------------------------------------------------------------------
constexpr int PB = 128;            // panel / diagonal block size
constexpr int CHOL_MAX_DIM = 8192; // largest factored dimension (the back solve keeps the solution vector in shared memory)

__device__ __forceinline__ void bulk_s2g(void* gdst, const void* ssrc, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// row m of every G_c <- (r_c, 1 + |r_c|^2 / lmin_c), lmin_c a lower bound of the smallest eigenvalue of G_c, so that
// the bordered matrix stays positive definite (r' G^-1 r <= |r|^2 / lambda_min):  n-form G = X D X' + I: lmin = 1;
// q-form P = (X'X + D^-1)/tau2: lmin = 1 / (tau2 max_j S_j)  (S, tau2 given).  grid = C, block = 256.
__global__ void __launch_bounds__(256) k_augment(double* __restrict__ G, size_t chain_stride, int N, int m,
                                                 const double* __restrict__ r, int r_stride,
                                                 const double* __restrict__ S, const double* __restrict__ tau2) {
  __shared__ double red[8], redm[8];
  const int c = blockIdx.x, tid = threadIdx.x;
  double* Gc = G + (size_t)c * chain_stride;
  const double* rc = r + (size_t)c * r_stride;
  double s = 0.0, mx = 0.0;
  for (int k = tid; k < m; k += 256) {
    const double v = rc[k];
    Gc[(size_t)k * N + m] = v;
    s += v * v;
    if (S) mx = fmax(mx, S[(size_t)c * r_stride + k]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((tid & 31) == 0) { red[tid >> 5] = s; redm[tid >> 5] = mx; }
  __syncthreads();
  if (tid == 0) {
    double t = 0.0, mm = 0.0;
    for (int w = 0; w < 8; ++w) { t += red[w]; mm = fmax(mm, redm[w]); }
    const double inv_lmin = S ? tau2[c] * mm : 1.0;
    Gc[(size_t)m * N + m] = 1.0 + t * inv_lmin;
  }
}

// ---- warp-level 32 x 32 x 32 products on the FP64 tensor cores, operands column-major in shared memory ----
// One 8 x 8 output fragment of C = A B (A, B 32 x 32, leading dimensions lda / ldb == 4 mod 16 doubles -> conflict-free):
// the MMA runs on the transposed problem (M along the columns n of C, N along its rows m), so that the accumulator pair
// of a lane is rows m0 + 2 lk, + 1 of column n0 + lr -> one 16-byte store.   acc += A[m0.., :] B[:, n0..]
__device__ __forceinline__ void frag_mm32(double& c0, double& c1, const double* Aop, int lda, const double* Bop, int ldb,
                                          int m0, int n0, int lk, int lr) {
#pragma unroll
  for (int k4 = 0; k4 < 8; ++k4) {
    const int k = k4 * 4 + lk;
    const double bt = Bop[(n0 + lr) * ldb + k];       // "A" operand of the transposed problem: B'[n][k]
    const double at = Aop[k * lda + m0 + lr];         // "B" operand: A'[k][m]
    dmma884(c0, c1, bt, at);
  }
}

// NF independent fragments at once, k outermost: the NF accumulator chains interleave, so the warp is bound by the DMMA
// issue rate instead of the latency of one dependent chain (8 DMMAs back to back)
template <int NF>
__device__ __forceinline__ void frag_mm32_batch(double (&c)[NF][2], const double* const (&Aop)[NF], const int (&lda)[NF],
                                                const double* const (&Bop)[NF], const int (&ldb)[NF],
                                                const int (&m0)[NF], const int (&n0)[NF], int lk, int lr) {
#pragma unroll
  for (int k4 = 0; k4 < 8; ++k4) {
    const int k = k4 * 4 + lk;
#pragma unroll
    for (int i = 0; i < NF; ++i) {
      const double bt = Bop[i][(n0[i] + lr) * ldb[i] + k];
      const double at = Aop[i][k * lda[i] + m0[i] + lr];
      dmma884(c[i][0], c[i][1], bt, at);
    }
  }
}

constexpr int PLD = 132;           // column stride of the 128 x 128 working block (== 4 mod 16: conflict-free fragments)
constexpr int XD_LD = 36;          // column stride of the 32 x 32 scratch blocks (== 4 mod 16)
constexpr int XD_BLK = 32 * XD_LD;
// Optional in-kernel time stamps (lab builds only, -DBNR_POTF2_STAMPS): thread 0 of chain 0 records globaltimer at the
// phase boundaries of k_potf2_inv into bnr::g_potf2_stamps.
#ifdef BNR_POTF2_STAMPS
__device__ unsigned long long g_potf2_stamps[64];
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;\n" : "=l"(t));
  return t;
}
#define STAMP(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_potf2_stamps[i] = gtimer(); } while (0)
#else
#define STAMP(i) do { } while (0)
#endif

constexpr size_t POTF2_SMEM = sizeof(double) * ((size_t)PB * PLD + 8 * XD_BLK + PB) + 8 * (4 + PB);
#ifndef BNR_POTF2_PUBLISH
#define BNR_POTF2_PUBLISH 8
#endif
// the pivot warp publishes its columns in groups of this many: one release-arrive per column cost it ~0.25 us per 32
// columns more than one per 8 (measured: pivot block 4.2-4.9 / 4.0-4.1 / 3.9 us with 1 / 4 / 8)
constexpr int POTF2_PUBLISH = BNR_POTF2_PUBLISH;
constexpr int PU_ROWS = 32;        // k-rows of L(J, J-1) per staged chunk of the in-kernel diagonal update
static_assert(2 * PU_ROWS * PLD <= 8 * XD_BLK, "the update ring aliases the inverse scratch");

// One 32-k-row chunk of the diagonal-block update Delta += Lp Lp' (lower triangle by 8 x 8 fragments): warp W owns the
// fragment columns W and 15 - W (17 fragments, the same count for every warp) exactly like syrk_diag_tile.
template <int W>
__device__ __forceinline__ void potf2_update_chunk(const double* buf, int lk, int lr, double (&acc)[17][2]) {
  constexpr int NA = 16 - W;
#pragma unroll
  for (int k4 = 0; k4 < PU_ROWS / 4; ++k4) {
    const double* row = buf + (k4 * 4 + lk) * PLD + lr;
    const double a0 = row[W * 8], a1 = row[(15 - W) * 8];
    double bfr[NA];
#pragma unroll
    for (int i = 0; i < NA; ++i) bfr[i] = row[(W + i) * 8];
#pragma unroll
    for (int t = 0; t < NA; ++t) dmma884(acc[t][0], acc[t][1], a0, bfr[t]);
#pragma unroll
    for (int t = NA; t < 17; ++t) dmma884(acc[t][0], acc[t][1], a1, bfr[t - 1 - W]);
  }
}

template <int W>
__device__ __forceinline__ void potf2_update_apply(double* A, int lk, int lr, const double (&acc)[17][2]) {
  constexpr int NA = 16 - W;
#pragma unroll
  for (int t = 0; t < 17; ++t) {
    const int jf = (t < NA) ? W : 15 - W;
    const int ifr = (t < NA) ? W + t : t - 1;
    double2* p = reinterpret_cast<double2*>(A + (jf * 8 + lr) * PLD + ifr * 8 + 2 * lk);
    double2 v = *p;
    v.x -= acc[t][0];
    v.y -= acc[t][1];
    *p = v;
  }
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}

// Four 8 x 8 output fragments (rows m0.., columns 0..31) of a 32^3 block product on the FP64 tensor cores, operands
// column-major in shared memory:  c += A[m0.., :] B.   KIND 1: A is lower triangular (k <= m: the k-steps beyond the
// fragment's rows are skipped), KIND 2: B is lower triangular (k >= n: fragment i starts at k = 8 i), KIND 0: full.
// The skipped products are exact zeros, so the result does not depend on KIND.
template <int KIND>
__device__ __forceinline__ void quad_term(double (&c)[4][2], const double* Aop, int lda, const double* Bop, int ldb,
                                          int m0, int lk, int lr) {
  const int k4hi = (KIND == 1) ? (m0 >> 2) + 2 : 8;
#pragma unroll
  for (int k4 = 0; k4 < 8; ++k4) {
    if (k4 < k4hi) {
      const int k = k4 * 4 + lk;
      const double at = Aop[k * lda + m0 + lr];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (KIND != 2 || k4 >= 2 * i) dmma884(c[i][0], c[i][1], Bop[(8 * i + lr) * ldb + k], at);
    }
  }
}
__device__ __forceinline__ void quad_store(double* dst, int ldd, int m0, double sign, const double (&c)[4][2], int lk, int lr) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<double2*>(dst + (8 * i + lr) * ldd + m0 + 2 * lk) = make_double2(sign * c[i][0], sign * c[i][1]);
}

// Panel kernel of the blocked Cholesky, one CTA per chain:
//   (late update)  D = G[J, J] - L[J, J-1] L[J, J-1]'   when `late` (the contributions of the panels before J-1 were
//                  applied earlier by k_chol_update; L[J, J-1] streams through a two-buffer TMA ring into DMMA fragments)
//   (factor)       D = L_JJ L_JJ'                       in shared memory, 32-wide steps
//   (inverse)      Linv_J = L_JJ^-1                     in place, 32^3 DMMA block products
__global__ void __launch_bounds__(256) k_potf2_inv(double* __restrict__ G, size_t chain_stride, int N, int J,
                                                   double* __restrict__ Linv, int T, int* status, int late) {
  extern __shared__ __align__(16) double sm[];
  double* A = sm;                          // column-major 128 x 128 working block: A[col * PLD + row]
  double* Xd = sm + PB * PLD;              // [4] inverses of the 32 x 32 diagonal blocks, column stride XD_LD
  double* Tm = Xd + 4 * XD_BLK;            // [4] intermediate products of the inverse
  double* dall = Tm + 4 * XD_BLK;          // [128] reciprocal diagonal of L
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(dall + PB);   // [0] block load, [1..2] update ring
  unsigned long long* colbar = bar + 4;    // [128] "column j of L is published" (one phase each, one arrival)
  const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lk = lane & 3, lr = lane >> 2;
  double* Gc = G + (size_t)c * chain_stride;
  double* D = Gc + (size_t)J * PB * N + (size_t)J * PB;
  STAMP(0);
  if (tid < PB) mbar_init(&colbar[tid], 1);
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_init(&bar[2], 1); }
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  __syncthreads();
  // the 128 x 128 block: 16-byte LDGSTS copies by all threads (a 1 KB column = one coalesced request of 64 threads;
  // 128 one-kilobyte bulk copies took 4.4 us here -- the TMA engine accepts ~1 such request per 30 ns)
  {
    const int chunk = tid & 63, c0 = tid >> 6;
#pragma unroll 8
    for (int i = 0; i < PB / 4; ++i) {
      const int col = c0 + 4 * i;
      cp_async16(A + col * PLD + chunk * 2, D + (size_t)col * N + chunk * 2);
    }
    cp_async_commit();
  }
  if (late && J > 0) {
    // L[J, J-1]: rows J*128.., columns (J-1)*128..: element (row, k) at Lp[row + N * k]
    const double* Lp = Gc + (size_t)(J - 1) * PB * N + (size_t)J * PB;
    double* ring = Xd;
    auto issue = [&](int chunk) {            // by warp 1: lane l loads k-row chunk * 32 + l
      double* buf = ring + (chunk & 1) * PU_ROWS * PLD;
      if (lane == 0) mbar_expect_tx(&bar[1 + (chunk & 1)], PU_ROWS * PB * 8);
      __syncwarp();
      bulk_g2s(buf + lane * PLD, Lp + (size_t)(chunk * PU_ROWS + lane) * N, PB * 8, &bar[1 + (chunk & 1)]);
    };
    if (warp == 1) { issue(0); issue(1); }
    double acc[17][2];
#pragma unroll
    for (int t = 0; t < 17; ++t) acc[t][0] = acc[t][1] = 0.0;
    constexpr int NCH = PB / PU_ROWS;
    for (int chunk = 0; chunk < NCH; ++chunk) {
      mbar_wait(&bar[1 + (chunk & 1)], (chunk >> 1) & 1);
      const double* buf = ring + (chunk & 1) * PU_ROWS * PLD;
      switch (warp) {
        case 0: potf2_update_chunk<0>(buf, lk, lr, acc); break;
        case 1: potf2_update_chunk<1>(buf, lk, lr, acc); break;
        case 2: potf2_update_chunk<2>(buf, lk, lr, acc); break;
        case 3: potf2_update_chunk<3>(buf, lk, lr, acc); break;
        case 4: potf2_update_chunk<4>(buf, lk, lr, acc); break;
        case 5: potf2_update_chunk<5>(buf, lk, lr, acc); break;
        case 6: potf2_update_chunk<6>(buf, lk, lr, acc); break;
        default: potf2_update_chunk<7>(buf, lk, lr, acc); break;
      }
      if (chunk + 2 < NCH) {
        __syncthreads();                     // every warp is done with this buffer before the TMA engine refills it
        if (warp == 1) issue(chunk + 2);
      }
    }
    cp_async_wait<0>();
    __syncthreads();
    switch (warp) {
      case 0: potf2_update_apply<0>(A, lk, lr, acc); break;
      case 1: potf2_update_apply<1>(A, lk, lr, acc); break;
      case 2: potf2_update_apply<2>(A, lk, lr, acc); break;
      case 3: potf2_update_apply<3>(A, lk, lr, acc); break;
      case 4: potf2_update_apply<4>(A, lk, lr, acc); break;
      case 5: potf2_update_apply<5>(A, lk, lr, acc); break;
      case 6: potf2_update_apply<6>(A, lk, lr, acc); break;
      default: potf2_update_apply<7>(A, lk, lr, acc); break;
    }
    __syncthreads();
  } else {
    cp_async_wait<0>();
    __syncthreads();
  }
  STAMP(1);
  // ---- factor + inverse, software-pipelined over the four 32-wide steps ----
  // Only the 32 x 32 pivot chain (128 dependent rsqrt steps of ~110 ns) is serial; everything else hangs off it:
  //   warp 0        pivot warp: factors the 32 x 32 diagonal block of step s in registers and PUBLISHES every finished
  //                 column of L (written in place + one mbarrier arrival per column).
  //   warps 1-3     rows below the diagonal block, one thread per row, eliminating column j as soon as it is published
  //                 (right-looking: the only inputs of column step j are that column of L_ss and its reciprocal pivot), so
  //                 the rows are complete one column behind the pivot warp instead of a phase after it.
  //   warp 7        the same elimination on the 32 unit rows e_j: x = e_j L_ss^-T is column j of L_ss^-1 -> Xd[s].
  //   warp 4        (the pivot warp's scheduler partition: no FP64) stores finished block columns of L and rows of Linv.
  //   warps 5, 6    + the row warps that have run out of rows: the inverse, bordered by block rows, on the tensor cores:
  //                   Y(i,j) = sum_{j<=k<i} L(i,k) X(k,j),  X(i,j) = -Xd_i Y(i,j),  X(i,i) = Xd_i = L_ii^-1
  //                 -- every product except the last block row is ready before the pivot chain ends.
  // Between steps only the 10 fragments of the NEXT diagonal block are updated by everybody (Tdiag); the rest of the
  // trailing update runs under the next pivot block (the lock-step rows simply start late and catch up).
  auto blkA = [&](int i, int j) { return A + (32 * j) * PLD + 32 * i; };        // block (i, j) of the working matrix
  auto trail = [&](int s, int f0, int f1, int slot, int nslots) {
    // A[r][cc] -= sum_p L[r][o+p] L[cc][o+p] for the 8 x 8 fragments f0 <= f < f1 of the lower triangle below step s
    // (fragment f = fr (fr + 1) / 2 + fc, fc <= fr; diagonal fragments are computed whole: what lands above the
    // diagonal is never read).  f < 10 is the next diagonal block.
    const int o = s * 32, base = o + 32;
    const double* P = A + o * PLD + base;             // panel: P[p * PLD + i] = L[base + i][o + p]
    for (int f = f0 + slot; f < f1; f += nslots) {
      int fr = 0;
      while ((fr + 1) * (fr + 2) / 2 <= f) ++fr;
      const int fc = f - fr * (fr + 1) / 2;
      double c0 = 0.0, c1 = 0.0;
#pragma unroll
      for (int k4 = 0; k4 < 8; ++k4) {
        const int k = k4 * 4 + lk;
        dmma884(c0, c1, P[k * PLD + fc * 8 + lr], P[k * PLD + fr * 8 + lr]);     // M along the column index cc
      }
      double2* dst = reinterpret_cast<double2*>(A + (base + fc * 8 + lr) * PLD + base + fr * 8 + 2 * lk);
      double2 v = *dst;
      v.x -= c0; v.y -= c1;
      *dst = v;
    }
  };
  bool bad = false;
  for (int s = 0; s < PB / 32; ++s) {
    const int o = s * 32;
    if (warp == 0) {
      // 32 x 32 Cholesky in registers: lane r holds row o+r (columns o .. o+31); entries above the diagonal are
      // whatever the block held there and never reach a valid entry.  Column j is broadcast through its final place in
      // shared memory, and the next pivot is updated and fetched FIRST so its rsqrt chain runs under the remaining
      // updates of column j.
      double a[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) a[j] = A[(o + j) * PLD + o + lane];
      double djj = __shfl_sync(0xffffffffu, a[0], 0);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (!(djj > 0.0)) bad = true;
        const double inv = rsqrt(djj);
        const double lj = (lane == j) ? djj * inv : a[j] * inv;
        double* lc = A + (o + j) * PLD + o;
        if (lane >= j) lc[lane] = lj;
        if (lane == j) dall[o + j] = inv;
        if (j + 1 < 32) {
          // next pivot first, without the round trip through shared memory: for the lane that owns row j + 1 the
          // multiplier lc[j + 1] is its own lj (bit-identical to the generic update below)
          if (lane == j + 1) a[j + 1] = fma(-lj, lj, a[j + 1]);
          djj = __shfl_sync(0xffffffffu, a[j + 1], j + 1);
        }
        __syncwarp();
        if ((j % POTF2_PUBLISH) == POTF2_PUBLISH - 1 && lane == 0)
          mbar_arrive(&colbar[o + j]);                // release: columns <= j and their reciprocal pivots are visible
        if (j + 1 < 32 && lane != j + 1) a[j + 1] = fma(-lj, lc[j + 1], a[j + 1]);
#pragma unroll
        for (int cc = j + 2; cc < 32; ++cc) a[cc] -= lj * lc[cc];
      }
      STAMP(2 + 3 * s);
    } else {
      if (s > 0) {
        trail(s - 1, 10, (PB - o) / 8 * ((PB - o) / 8 + 1) / 2, warp - 1, 7);    // the rest of the previous step's update
        named_bar_sync(1, 224);
      }
      const int nrow_warps = 3 - s;                   // warps 1 .. nrow_warps own the rows below this step's block
      if (warp == 7 || warp <= nrow_warps) {
        const bool inv_row = (warp == 7);
        double* row = A + o * PLD + o + 32 + (warp - 1) * 32 + lane;
        double x[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = inv_row ? ((j == lane) ? 1.0 : 0.0) : row[j * PLD];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if ((j % POTF2_PUBLISH) == 0) mbar_wait(&colbar[o + j + POTF2_PUBLISH - 1], 0);
          const double xj = x[j] * dall[o + j];
          x[j] = xj;
#pragma unroll
          for (int k = j + 1; k < 32; ++k) x[k] -= xj * A[(o + j) * PLD + o + k];
        }
        if (inv_row) {
#pragma unroll
          for (int j = 0; j < 32; ++j) Xd[s * XD_BLK + lane * XD_LD + j] = (j >= lane) ? x[j] : 0.0;
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) row[j * PLD] = x[j];
        }
      } else if (warp == 4) {
        // finished pieces go to global memory through the TMA engine (one bulk store per lane and column): direct
        // stores from this warp clogged the memory pipe it shares with the pivot warp (the following diagonal-block
        // update took 5 us instead of 1)
        fence_async_smem();                           // the pieces were written through the generic proxy
        if (s > 0) {
          // block column s-1 of the factor is final (rows from its diagonal block down)
          const int ob = o - 32;
          bulk_s2g(D + (size_t)(ob + lane) * N + ob, A + (ob + lane) * PLD + ob, (unsigned)(PB - ob) * 8u);
        }
        if (s == 3) {
          // rows 0..63 of the inverse are final: X(0,0) = Xd0, X(1,0), X(1,1) = Xd1 (zeros above the diagonal of the
          // diagonal blocks come from Xd; the zero blocks above them are never written: the buffer is zeroed at creation)
          double* dst = Linv + ((size_t)c * T + J) * PB * PB;
          bulk_s2g(dst + (size_t)lane * PB, Xd + lane * XD_LD, 256u);
          bulk_s2g(dst + (size_t)lane * PB + 32, A + lane * PLD + 32, 256u);
          bulk_s2g(dst + (size_t)(32 + lane) * PB + 32, Xd + XD_BLK + lane * XD_LD, 256u);
        }
        bulk_commit();
        bulk_wait_read0();                            // the sources may be overwritten after this step's barrier
      } else if (s > 0) {
        // the inverse, one block row behind the factorisation.  Slots: warps 5, 6, then the idle row warps 3, 2, 1.
        const int slot = (warp >= 5) ? warp - 5 : 5 - warp, nslots = 2 + s;
        double* Tm0 = Tm; double* Tm1 = Tm + XD_BLK; double* Tm2 = Tm + 2 * XD_BLK; double* Tm3 = Tm + 3 * XD_BLK;
        const double* X0 = Xd; const double* X1 = Xd + XD_BLK; const double* X2 = Xd + 2 * XD_BLK;
        if (s == 1) {
          for (int q = slot; q < 4; q += nslots) {                               // Y(1,0) = L(1,0) Xd0
            double cq[4][2] = {};
            quad_term<2>(cq, blkA(1, 0), PLD, X0, XD_LD, 8 * q, lk, lr);
            quad_store(Tm0, XD_LD, 8 * q, 1.0, cq, lk, lr);
          }
        } else if (s == 2) {
          for (int q = slot; q < 8; q += nslots) {
            const int m0 = 8 * (q & 3);
            double cq[4][2] = {};
            if (q < 4) {                                                         // X(1,0) = -Xd1 Y(1,0)   (in place of L(1,0))
              quad_term<1>(cq, X1, XD_LD, Tm0, XD_LD, m0, lk, lr);
              quad_store(blkA(1, 0), PLD, m0, -1.0, cq, lk, lr);
            } else {                                                             // Y(2,1) = L(2,1) Xd1
              quad_term<2>(cq, blkA(2, 1), PLD, X1, XD_LD, m0, lk, lr);
              quad_store(Tm2, XD_LD, m0, 1.0, cq, lk, lr);
            }
          }
          named_bar_sync(2, 32 * nslots);
          for (int q = slot; q < 4; q += nslots) {                               // Y(2,0) = L(2,0) Xd0 + L(2,1) X(1,0)
            double cq[4][2] = {};
            quad_term<2>(cq, blkA(2, 0), PLD, X0, XD_LD, 8 * q, lk, lr);
            quad_term<0>(cq, blkA(2, 1), PLD, blkA(1, 0), PLD, 8 * q, lk, lr);
            quad_store(Tm1, XD_LD, 8 * q, 1.0, cq, lk, lr);
          }
        } else {
          for (int q = slot; q < 12; q += nslots) {
            const int m0 = 8 * (q & 3);
            double cq[4][2] = {};
            if (q < 8) {                                                         // X(2,j) = -Xd2 Y(2,j), j = 0, 1
              quad_term<1>(cq, X2, XD_LD, q < 4 ? Tm1 : Tm2, XD_LD, m0, lk, lr);
              quad_store(blkA(2, q >> 2), PLD, m0, -1.0, cq, lk, lr);
            } else {                                                             // Y(3,2) = L(3,2) Xd2
              quad_term<2>(cq, blkA(3, 2), PLD, X2, XD_LD, m0, lk, lr);
              quad_store(Tm3, XD_LD, m0, 1.0, cq, lk, lr);
            }
          }
          named_bar_sync(2, 32 * nslots);
          for (int q = slot; q < 8; q += nslots) {
            const int m0 = 8 * (q & 3);
            double cq[4][2] = {};
            if (q < 4) {                                                         // Y(3,0) = L(3,0) Xd0 + L(3,1) X(1,0) + L(3,2) X(2,0)
              quad_term<2>(cq, blkA(3, 0), PLD, X0, XD_LD, m0, lk, lr);
              quad_term<0>(cq, blkA(3, 1), PLD, blkA(1, 0), PLD, m0, lk, lr);
              quad_term<0>(cq, blkA(3, 2), PLD, blkA(2, 0), PLD, m0, lk, lr);
              quad_store(Tm0, XD_LD, m0, 1.0, cq, lk, lr);
            } else {                                                             // Y(3,1) = L(3,1) Xd1 + L(3,2) X(2,1)
              quad_term<2>(cq, blkA(3, 1), PLD, X1, XD_LD, m0, lk, lr);
              quad_term<0>(cq, blkA(3, 2), PLD, blkA(2, 1), PLD, m0, lk, lr);
              quad_store(Tm1, XD_LD, m0, 1.0, cq, lk, lr);
            }
          }
        }
      }
    }
    fence_async_smem();                               // what this step wrote will be read by warp 4's bulk stores
    __syncthreads();                                  // step s: pivot block, rows below, Xd[s], side work all done
    STAMP(3 + 3 * s);
    if (s + 1 < PB / 32) {
      trail(s, 0, 10, warp, 8);                       // the next diagonal block
      __syncthreads();
    }
    STAMP(4 + 3 * s);
  }
  if (bad && lane == 0) atomicOr(&status[c], BNR_ST_G_NOTPD_);
  // ---- tail: the last block row of the inverse, X(3,j) = -Xd3 Y(3,j) (in place of L(3,j), which warp 4 has stored) ----
  for (int q = warp; q < 12; q += 8) {
    const int m0 = 8 * (q & 3), j = q >> 2;
    double cq[4][2] = {};
    quad_term<1>(cq, Xd + 3 * XD_BLK, XD_LD, Tm + (j == 2 ? 3 : j) * XD_BLK, XD_LD, m0, lk, lr);
    quad_store(blkA(3, j), PLD, m0, -1.0, cq, lk, lr);
  }
  {
    // the last diagonal block of the factor (32 x 32; rows 96..127 of columns 96..127)
    const int col = 96 + (tid >> 3), r = 96 + 4 * (tid & 7);
    *reinterpret_cast<double2*>(D + (size_t)col * N + r) = *reinterpret_cast<const double2*>(A + col * PLD + r);
    *reinterpret_cast<double2*>(D + (size_t)col * N + r + 2) = *reinterpret_cast<const double2*>(A + col * PLD + r + 2);
  }
  __syncthreads();
  STAMP(14);
  {
    // rows 64..127 of Linv_J: X(2,0..1), Xd2 | X(3,0..2), Xd3; the zero blocks above the diagonal are never written
    double* dst = Linv + ((size_t)c * T + J) * PB * PB;
    const int r = 64 + 2 * (tid & 31), c0 = tid >> 5;
#pragma unroll 4
    for (int i = 0; i < PB / 8; ++i) {
      const int cc = c0 + 8 * i;
      if ((cc >> 5) > (r >> 5)) continue;
      const double* src = ((r >> 5) == (cc >> 5)) ? Xd + (r >> 5) * XD_BLK + (cc & 31) * XD_LD + (r & 31) : A + cc * PLD + r;
      *reinterpret_cast<double2*>(dst + (size_t)cc * PB + r) = *reinterpret_cast<const double2*>(src);
    }
  }
  if (warp == 4) bulk_wait0();                        // this warp's bulk stores have landed
  STAMP(15);
  STAMP(16);
  STAMP(17);
}

// ------------------------------------------------------------------------------------------------------------
// Depth-128 tile kernels of the latency schedule, operands resident in shared memory.
//   MODE 2  panel solve     C[strip] = G[strip, J] Linv_J'                  (T1 / T2 of launch_cholesky)
//   MODE 1  late update     C[strip] -= L[strip, J] L[J+1, J]'              (Ulate)
// One CTA per (chain, row strip of 32 or 64 rows).  The ring-fed kernels above spend ~8 us per tile on their eight
// pipeline stages (32 one-kilobyte TMA requests each, ~30 ns per request) around 2-4 us of DMMA work; here the two
// operands arrive as 16-byte cp.async copies of all 256 threads (1.5-2 us for 160-200 KB), one barrier, then the same
// fragments in the same k order as syrk_strip_tile (bit-identical results), and a direct store.
// grid = (C, tiles * nstrip), block = 256.
// ------------------------------------------------------------------------------------------------------------
constexpr size_t small_tile_smem(int rows_i) { return sizeof(double) * ((size_t)PB * SY_LDS + (size_t)PB * (rows_i + 4)); }

template <int MODE, int NF>
__global__ void __launch_bounds__(256) k_small_tile(double* __restrict__ G, size_t chain_stride, int N, int jb, int ib_first,
                                                    int nstrip, const double* __restrict__ Bj, size_t bj_chain_stride,
                                                    int ldj, int kcol0, int tri) {
  extern __shared__ __align__(16) double sm[];
  constexpr int ROWS = NF * 8, LDI = ROWS + 4;       // LDI == 4 mod 16: conflict-free fragment loads
  double* sJ = sm;                                   // [128 k][132]  j operand: sJ[k * SY_LDS + row]
  double* sI = sm + PB * SY_LDS;                     // [128 k][LDI]  i operand (this CTA's row strip)
  const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lk = lane & 3, lr = lane >> 2;
  const int ib = ib_first + (int)blockIdx.y / nstrip, strip = (int)blockIdx.y % nstrip;
  const int i0 = ib * PB + strip * ROWS;
  double* Gc = G + (size_t)c * chain_stride;
  // sources (k-major: element (row, k) at base[row + ld * k]); kcol0 = first column of the contributing panel
  const double* srcJ = Bj + (size_t)c * bj_chain_stride;
  const double* srcI = Gc + (size_t)kcol0 * N + i0;
  // Only the part of the j operand that can reach an output is loaded.  MODE 2: Linv_J is lower triangular (entry
  // (col, k) with k <= col): k-row k is needed from its own 8-column fragment on.  MODE 1 on the diagonal tile (tri):
  // only the lower triangle of the tile is ever read, so the strip needs the columns up to its own last row.
  const int jmax = (MODE == 1 && tri) ? (strip + 1) * ROWS : PB;       // columns of the tile this CTA touches
  {
    const int chunk = tid & 63, r0 = tid >> 6;       // j operand: 128 k-rows of 64 16-byte chunks
#pragma unroll 8
    for (int i = 0; i < PB / 4; ++i) {
      const int k = r0 + 4 * i;
      const bool need = (MODE == 2) ? (chunk >= 4 * (k >> 3)) : (2 * chunk < jmax);
      if (need) cp_async16(sJ + k * SY_LDS + chunk * 2, srcJ + (size_t)k * ldj + chunk * 2);
    }
    constexpr int CPR = ROWS / 2;                    // 16-byte chunks per k-row of the i operand
    for (int id = tid; id < PB * CPR; id += 256) {
      const int k = id / CPR, ch = id % CPR;
      cp_async16(sI + k * LDI + ch * 2, srcI + (size_t)k * N + ch * 2);
    }
    cp_async_commit();
  }
  const int f0 = (MODE == 2) ? warp : 2 * warp, f1 = (MODE == 2) ? 15 - warp : 2 * warp + 1;
  const int lim0 = 2 * f0 + 2, lim1 = 2 * f1 + 2;    // MODE 2: k4-steps that can still reach the column fragment
  const int j0 = jb * PB;
  double acc[2][NF][2];
  const bool on0 = f0 * 8 < jmax, on1 = f1 * 8 < jmax;   // column fragments inside the needed part (always, unless tri)
  if (MODE == 1) {                                   // in place: start from the tile's current values
#pragma unroll
    for (int mf = 0; mf < 2; ++mf) {
      const int j = j0 + (mf == 0 ? f0 : f1) * 8 + lr;
      if (mf == 0 ? on0 : on1) {
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) {
          const double2 v = *reinterpret_cast<const double2*>(Gc + (size_t)j * N + i0 + nf * 8 + 2 * lk);
          acc[mf][nf][0] = v.x; acc[mf][nf][1] = v.y;
        }
      }
    }
  } else {
#pragma unroll
    for (int mf = 0; mf < 2; ++mf)
#pragma unroll
      for (int nf = 0; nf < NF; ++nf) acc[mf][nf][0] = acc[mf][nf][1] = 0.0;
  }
  cp_async_wait<0>();
  __syncthreads();
  if (MODE == 1 && !on0) return;                     // (f0 < f1: nothing of this warp lies in the lower triangle)
#pragma unroll 4
  for (int k4 = 0; k4 < PB / 4; ++k4) {
    const int kr = k4 * 4 + lk;
    const double* rj = sJ + kr * SY_LDS + lr;
    const double* ri = sI + kr * LDI + lr;
    double af0 = 0.0, af1 = 0.0;
    if (MODE == 2) {                                 // (entries with k beyond the fragment's reach were not loaded)
      if (k4 < lim0) af0 = rj[f0 * 8];
      if (k4 < lim1) af1 = rj[f1 * 8];
    } else {
      af0 = neg_bits(rj[f0 * 8]);
      if (on1) af1 = neg_bits(rj[f1 * 8]);
    }
    double bf[NF];
#pragma unroll
    for (int nf = 0; nf < NF; ++nf) bf[nf] = ri[nf * 8];
    if (MODE == 2) {
      if (k4 < lim0) {
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) dmma884(acc[0][nf][0], acc[0][nf][1], af0, bf[nf]);
      }
      if (k4 < lim1) {
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) dmma884(acc[1][nf][0], acc[1][nf][1], af1, bf[nf]);
      }
    } else {
#pragma unroll
      for (int nf = 0; nf < NF; ++nf) {
        dmma884(acc[0][nf][0], acc[0][nf][1], af0, bf[nf]);
        if (on1) dmma884(acc[1][nf][0], acc[1][nf][1], af1, bf[nf]);
      }
    }
  }
#pragma unroll
  for (int mf = 0; mf < 2; ++mf) {
    const int j = j0 + (mf == 0 ? f0 : f1) * 8 + lr;
    if (mf == 1 && !on1) continue;
#pragma unroll
    for (int nf = 0; nf < NF; ++nf)
      *reinterpret_cast<double2*>(Gc + (size_t)j * N + i0 + nf * 8 + 2 * lk) = make_double2(acc[mf][nf][0], acc[mf][nf][1]);
  }
}

// ------------------------------------------------------------------------------------------------------------
// backward solve  L' x = b,  b = w (+ addz), w = row m of the factor (see above).  One CTA per chain streams the factor
// once, bottom-right to top-left, in parts of BW_COLS columns x 128 rows through a ring of TMA tensor-map boxes (one
// 32 KB request per part, full / empty mbarriers, a producer warp with one active lane).  History: 1 KB bulk copies from
// a producer warp (the TMA engine accepts ~one request per 30 ns: 157 us for a 1024 x 1024 factor), then 16-byte
// LDGSTS copies by all threads with a __syncthreads per part (118 us: ~500 of the ~1600 cycles per part were LDGSTS
// issue, the rest the per-part barrier chain), now box loads and barriers only at block-column boundaries.
// 8 consumer warps turn every part into BW_COLS dot products (warp per BW_COLS / 8 columns, lanes over the rows):
//   off-diagonal block (I, J): b_J -= L[I, J]' x_I          diagonal block J: x_J = Linv_J' b_J
// Within a block column every warp owns its entries of b_J, so the consumers only meet before and after the diagonal
// block.   grid = C, block = 288 (8 consumer warps + the producer warp).
// ------------------------------------------------------------------------------------------------------------
#ifndef BNR_BW_COLS
#define BNR_BW_COLS 32
#endif
constexpr int BW_COLS = BNR_BW_COLS;             // columns per staged part of a 128 x 128 block
constexpr int BW_NH = PB / BW_COLS;              // parts per block
constexpr int BW_CPW = BW_COLS / 8;              // columns per warp
constexpr int BW_MAX_STAGES = 6;
constexpr int BW_STAGE_DBL = BW_COLS * PB;
constexpr int BW_THREADS = 288;
static int bwd_stages(int N) {
  const size_t budget = 227 * 1024 - sizeof(double) * (size_t)N - 512;
  int ns = (int)(budget / (sizeof(double) * BW_STAGE_DBL));
  return ns > BW_MAX_STAGES ? BW_MAX_STAGES : ns;
}
static size_t bwd_smem(int N) { return sizeof(double) * ((size_t)bwd_stages(N) * BW_STAGE_DBL + N) + 256; }

// tmG: the chain-stacked factor {rows, columns x chains}, tmL: the stacked panel inverses {128, 128 x panels x chains};
// boxes of 128 rows x BW_COLS columns (dense in shared memory: part[col * 128 + row])
__global__ void __launch_bounds__(BW_THREADS) k_bwd_stream(const __grid_constant__ CUtensorMap tmG,
                                                           const __grid_constant__ CUtensorMap tmL,
                                                           const double* __restrict__ G, size_t chain_stride, int N, int m,
                                                           double* __restrict__ out, int out_stride,
                                                           const double* __restrict__ addz, int addz_stride, int nstages) {
  extern __shared__ __align__(128) double sm[];
  double* ring = sm;
  double* x = sm + (size_t)nstages * BW_STAGE_DBL;     // [N] right-hand side, overwritten block by block by x
  unsigned long long* full = reinterpret_cast<unsigned long long*>(x + N);
  unsigned long long* empty = full + BW_MAX_STAGES;
  const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = N / PB;
  const double* Gc = G + (size_t)c * chain_stride;
  const int total = T * (T + 1) / 2 * BW_NH;
  if (tid == 0) {
    for (int s = 0; s < nstages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if (warp == 8) {
    // ---- producer: the parts in consumption order (J down, I from T-1 down to J, h up) ----
    if (lane == 0) {
      int pJ = T - 1, pI = T - 1, ph = 0;
      for (int p = 0; p < total; ++p) {
        const int stage = p % nstages;
        mbar_wait(&empty[stage], ((p / nstages) & 1) ^ 1);
        mbar_expect_tx(&full[stage], BW_STAGE_DBL * 8u);
        double* dst = ring + (size_t)stage * BW_STAGE_DBL;
        if (pI == pJ) tma_load_2d(dst, &tmL, 0, (c * T + pJ) * PB + ph * BW_COLS, &full[stage]);
        else tma_load_2d(dst, &tmG, pI * PB, c * N + pJ * PB + ph * BW_COLS, &full[stage]);
        if (++ph == BW_NH) { ph = 0; if (--pI < pJ) { --pJ; pI = T - 1; } }
      }
    }
    return;
  }

  for (int k = tid; k < N; k += 256)
    x[k] = (k < m) ? Gc[(size_t)k * N + m] + (addz ? addz[(size_t)c * addz_stride + k] : 0.0) : 0.0;
  named_bar_sync(1, 256);

  int it = 0;
  for (int J = T - 1; J >= 0; --J) {
    for (int I = T - 1; I >= J; --I) {
      const bool dg = (I == J);
      if (dg) named_bar_sync(1, 256);                  // every warp's updates of b_J are in
      const double* v = x + I * PB;
      const double v0 = v[lane], v1 = v[lane + 32], v2 = v[lane + 64], v3 = v[lane + 96];
      double xn[BW_NH][BW_CPW];                        // diagonal block: this warp's entries of x_J
#pragma unroll
      for (int h = 0; h < BW_NH; ++h, ++it) {
        const int stage = it % nstages;
        mbar_wait(&full[stage], (it / nstages) & 1);
        const double* B = ring + (size_t)stage * BW_STAGE_DBL + (size_t)(warp * BW_CPW) * PB + lane;
        double acc[BW_CPW];
#pragma unroll
        for (int q = 0; q < BW_CPW; ++q)
          acc[q] = B[q * PB] * v0 + B[q * PB + 32] * v1 + B[q * PB + 64] * v2 + B[q * PB + 96] * v3;
#pragma unroll
        for (int q = 0; q < BW_CPW; ++q) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
        }
        if (dg) {
#pragma unroll
          for (int q = 0; q < BW_CPW; ++q) xn[h][q] = acc[q];
        } else if (lane == 0) {
          double* bj = x + J * PB + h * BW_COLS + warp * BW_CPW;
#pragma unroll
          for (int q = 0; q < BW_CPW; ++q) bj[q] -= acc[q];
        }
        // Release the part only HERE, after the shuffles (and the update of b_J) have consumed every lane's loads.
        // An arrive right after the loads were ISSUED is not enough: it is not ordered behind shared-memory loads
        // that still wait in the memory pipe, the producer's next box then overwrote data that had not been read yet
        // -- chains differed from run to run under load (tools/determinism_check.py found it; the SASS showed the
        // DFMAs that consume the loads scheduled after the SYNCS.ARRIVE).
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
      }
      if (dg) {
        named_bar_sync(1, 256);                        // everybody has read b_J
        if (lane == 0) {
#pragma unroll
          for (int h = 0; h < BW_NH; ++h)
#pragma unroll
            for (int q = 0; q < BW_CPW; ++q) x[J * PB + h * BW_COLS + warp * BW_CPW + q] = xn[h][q];
        }
        named_bar_sync(1, 256);                        // x_J is visible
      }
    }
  }
  for (int k = tid; k < N; k += 256) out[(size_t)c * out_stride + k] = x[k];
}

// symmetric copy of G (lower -> full) into the aux buffer, for the parity tests.  grid = (np, C)
__global__ void k_copy_sym(const double* __restrict__ G, size_t chain_stride, int np, double* __restrict__ out) {
  const int c = blockIdx.y, j = blockIdx.x;
  const double* Gc = G + (size_t)c * chain_stride;
  double* o = out + (size_t)c * np * np;
  for (int i = threadIdx.x; i < np; i += blockDim.x) {
    const double v = (i >= j) ? Gc[(size_t)j * np + i] : Gc[(size_t)i * np + j];
    o[(size_t)j * np + i] = v;
  }
}

// ------------------------------------------------------------------------------------------------------------
// tall-skinny GEMM with X for all chains at once on the FP64 tensor cores: out[c][m] = sum_k A(m,k) in[c][k]
//   TRANS = 0: A(m,k) = X[m + np*k]  (M = np, K = qp)      TRANS = 1: A(m,k) = X[k + np*m]  (M = qp, K = np)
// CTA tile 128 (m) x 64 (chains), k-step 16, 3-stage cp.async pipeline, 8 warps as 4 (m) x 2 (chains), warp tile
// 32 x 32 = 4 x 4 DMMA m8n8k4 per k4-step.  K is split over gridDim.z (deterministic: partial sums go to a
// workspace and are added in a fixed order by k_splitk_reduce) so that ~2 waves of CTAs stream X exactly once.
// Shared tiles: X as [k][m] (ld 132) for TRANS 0 / [m][k] (ld 20) for TRANS 1, the vectors as [chain][k] (ld 20);
// both strides are == 4 mod 16 doubles, which makes the (row, k) fragment loads bank-conflict free.
// ------------------------------------------------------------------------------------------------------------
constexpr int XM_BM = 128, XM_BN = 64, XM_BK = 16, XM_STAGES = 3;
constexpr int XM_LDM = 132, XM_LDK = 20;
constexpr int XM_A_DBL = (XM_BK * XM_LDM > XM_BM * XM_LDK) ? XM_BK * XM_LDM : XM_BM * XM_LDK;   // 2560
constexpr int XM_STAGE_DBL = XM_A_DBL + XM_BN * XM_LDK;
constexpr size_t XMMA_SMEM = sizeof(double) * XM_STAGES * XM_STAGE_DBL;

__device__ __forceinline__ void cp_async16_zfill(void* smem, const void* gmem, bool valid) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(sz));
}

template <int TRANS>
__global__ void __launch_bounds__(256) k_xmma(const double* __restrict__ X, int np, int M, int K, int N,
                                              const double* __restrict__ in, int ldin, double* __restrict__ out,
                                              int ldout, int k_per_split) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lk = lane & 3, lr = lane >> 2, wm = warp & 3, wn = warp >> 2;
  const int m0 = blockIdx.x * XM_BM, n0 = blockIdx.y * XM_BN;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);
  const int nsteps = (kend - kbeg + XM_BK - 1) / XM_BK;
  double* o = out + (size_t)blockIdx.z * N * ldout;

  auto load_stage = [&](int step) {
    double* As = sm + (size_t)(step % XM_STAGES) * XM_STAGE_DBL;
    double* Vs = As + XM_A_DBL;
    const int k0 = kbeg + step * XM_BK;
    if (TRANS == 0) {
      // 16 k-rows of 128 contiguous m: 64 x 16-byte pieces per row
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int id = tid + 256 * r, kk = id >> 6, mm = (id & 63) * 2;
        cp_async16_zfill(As + kk * XM_LDM + mm, X + (size_t)(k0 + kk) * np + m0 + mm, m0 + mm < M);
      }
    } else {
      // 128 m-rows of 16 contiguous k: 8 x 16-byte pieces per row
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int id = tid + 256 * r, mm = id >> 3, kk = (id & 7) * 2;
        const bool ok = m0 + mm < M;
        cp_async16_zfill(As + mm * XM_LDK + kk, X + (size_t)(ok ? m0 + mm : 0) * np + k0 + kk, ok);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int id = tid + 256 * r, nn = id >> 3, kk = (id & 7) * 2;
      const bool ok = n0 + nn < N;
      cp_async16_zfill(Vs + nn * XM_LDK + kk, in + (size_t)(ok ? n0 + nn : 0) * ldin + k0 + kk, ok);
    }
  };

  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
  // chain fragments of this warp that hold a real chain: with few chains (a chain group of 2-3, config 4's 8) the 64-chain
  // tile is mostly padding, and its DMMAs -- not the stream of X -- bounded the kernel (2 us per k-step for 16 KB)
  int nfv = (N - n0 - wn * 32 + 7) >> 3;
  nfv = nfv < 0 ? 0 : (nfv > 4 ? 4 : nfv);

  for (int s = 0; s < XM_STAGES - 1; ++s) {
    if (s < nsteps) load_stage(s);
    cp_async_commit();
  }
  for (int step = 0; step < nsteps; ++step) {
    cp_async_wait<XM_STAGES - 2>();
    __syncthreads();
    if (step + XM_STAGES - 1 < nsteps) load_stage(step + XM_STAGES - 1);
    cp_async_commit();
    const double* As = sm + (size_t)(step % XM_STAGES) * XM_STAGE_DBL;
    const double* Vs = As + XM_A_DBL;
    if (nfv == 0) continue;                          // (a warp without a real chain only keeps the barriers company)
#pragma unroll
    for (int k4 = 0; k4 < XM_BK / 4; ++k4) {
      const int kr = k4 * 4 + lk;
      double af[4], bf[4];
#pragma unroll
      for (int mf = 0; mf < 4; ++mf)
        af[mf] = (TRANS == 0) ? As[kr * XM_LDM + wm * 32 + mf * 8 + lr] : As[(wm * 32 + mf * 8 + lr) * XM_LDK + kr];
#pragma unroll
      for (int nf = 0; nf < 4; ++nf) bf[nf] = (nf < nfv) ? Vs[(wn * 32 + nf * 8 + lr) * XM_LDK + kr] : 0.0;
#pragma unroll
      for (int nf = 0; nf < 4; ++nf) {
        if (nf < nfv) {
#pragma unroll
          for (int mf = 0; mf < 4; ++mf) dmma884(acc[mf][nf][0], acc[mf][nf][1], af[mf], bf[nf]);
        }
      }
    }
  }
  cp_async_wait<0>();
#pragma unroll
  for (int mf = 0; mf < 4; ++mf) {
    const int mm = m0 + wm * 32 + mf * 8 + lr;
#pragma unroll
    for (int nf = 0; nf < 4; ++nf) {
      const int nn = n0 + wn * 32 + nf * 8 + 2 * lk;
      if (mm < M) {
        if (nn < N) o[(size_t)nn * ldout + mm] = acc[mf][nf][0];
        if (nn + 1 < N) o[(size_t)(nn + 1) * ldout + mm] = acc[mf][nf][1];
      }
    }
  }
}

__global__ void k_splitk_reduce(const double* __restrict__ ws, double* __restrict__ out, size_t count, int splits) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  double s = 0.0;
  for (int z = 0; z < splits; ++z) s += ws[(size_t)z * count + i];
  out[i] = s;
}

static int x_times_splits(const Dims& d, int trans) {
  const int M = trans ? d.qp : d.np, K = trans ? d.np : d.qp;
  const int tiles = ((M + XM_BM - 1) / XM_BM) * ((d.C + XM_BN - 1) / XM_BN);
  int ks = (2 * 148 + tiles - 1) / tiles;
  const int kmax = K / (4 * XM_BK) > 0 ? K / (4 * XM_BK) : 1;
  if (ks > kmax) ks = kmax;
  if (ks > 64) ks = 64;
  if (ks < 1) ks = 1;
  return ks;
}

size_t x_times_workspace_doubles(const Dims& d) {
  const int a = x_times_splits(d, 0), b = x_times_splits(d, 1);
  const size_t wa = (size_t)a * d.C * d.np, wb = (size_t)b * d.C * d.qp;
  return wa > wb ? wa : wb;
}

void launch_x_times(const Engine& e, int trans, const double* in, double* out, double* ws, cudaStream_t s) {
  const Dims& d = e.d;
  const int M = trans ? d.qp : d.np, K = trans ? d.np : d.qp, N = d.C;
  const int ldin = trans ? d.np : d.qp, ldout = trans ? d.qp : d.np;
  const int ks = x_times_splits(d, trans);
  const int kper = ((K + ks - 1) / ks + XM_BK - 1) / XM_BK * XM_BK;
  dim3 grid((M + XM_BM - 1) / XM_BM, (N + XM_BN - 1) / XM_BN, ks);
  double* dst = ks == 1 ? out : ws;
  ++g_launches;
  if (trans) k_xmma<1><<<grid, 256, XMMA_SMEM, s>>>(e.X, d.np, M, K, N, in, ldin, dst, ldout, kper);
  else k_xmma<0><<<grid, 256, XMMA_SMEM, s>>>(e.X, d.np, M, K, N, in, ldin, dst, ldout, kper);
  if (ks > 1) {
    const size_t count = (size_t)N * ldout;
    ++g_launches; k_splitk_reduce<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(ws, out, count, ks);
  }
}

void linalg_setup() {
  cudaFuncSetAttribute(k_xmma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)XMMA_SMEM);
  cudaFuncSetAttribute(k_xmma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)XMMA_SMEM);
  cudaFuncSetAttribute(k_gram_syrk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SYRK_SMEM);
  cudaFuncSetAttribute(k_chol_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SYRK_SMEM);
  cudaFuncSetAttribute(k_trsm_dmma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SYRK_SMEM);
  cudaFuncSetAttribute(k_potf2_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POTF2_SMEM);
  cudaFuncSetAttribute(k_bwd_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(k_small_tile<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_tile_smem(32));
  cudaFuncSetAttribute(k_small_tile<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_tile_smem(32));
  cudaFuncSetAttribute(k_small_tile<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_tile_smem(16));
  cudaFuncSetAttribute(k_small_tile<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_tile_smem(16));
  cudaFuncSetAttribute(k_small_tile<1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_tile_smem(64));
  cudaFuncSetAttribute(k_small_tile<2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_tile_smem(64));
}

// k-splits of the SYRK for a handle of C chains: when chains x tiles is below ~4 waves of CTAs, a CTA (one 128 x 128
// tile over the whole contraction: 0.7 ms at q = 5050, 2.7 ms at q = 20100) is too coarse a unit of work to balance over
// 148 SMs, so the contraction is cut into up to 8 pieces of at least 24 k-steps.  The count depends on the handle's chain
// count only (not on the chain groups, which each launch a part of the chains).
// ---- tensor maps (host): a k-major operand {rows (contiguous), k-rows} with boxes of box_rows x 16 ----
typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TmapEncodeFn g_tmap_encode = nullptr;
bool tmap_setup() {
  if (g_tmap_encode) return true;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess ||
      qr != cudaDriverEntryPointSuccess || !fn)
    return false;
  g_tmap_encode = (TmapEncodeFn)fn;
  return true;
}
// rows beyond `rows` (the box is 4 rows wider than a tile) and k-rows beyond `krows` read as zeros
static CUresult encode_map(CUtensorMap* m, const double* base, uint64_t rows, uint64_t krows, uint64_t ld, uint32_t box_rows,
                           uint32_t box_k) {
  memset(m, 0, sizeof(*m));
  const cuuint64_t dims[2] = {rows, krows};
  const cuuint64_t strides[1] = {ld * sizeof(double)};
  const cuuint32_t box[2] = {box_rows, box_k};
  const cuuint32_t estr[2] = {1, 1};
  if (!tmap_setup()) return CUDA_ERROR_NOT_SUPPORTED;
  return g_tmap_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}
static CUtensorMap make_map(const double* base, uint64_t rows, uint64_t krows, uint64_t ld, uint32_t box_rows,
                            uint32_t box_k = SY_BK) {
  CUtensorMap m;
  const CUresult r = encode_map(&m, base, rows, krows, ld, box_rows, box_k);
  if (r != CUDA_SUCCESS) {
    // (bnr_create has already encoded every map of this handle's geometry, tmaps_check(): unreachable from the ABI)
    fprintf(stderr, "[bnr] cuTensorMapEncodeTiled failed (%d) for a %llu x %llu operand, box %u x %d\n", (int)r,
            (unsigned long long)rows, (unsigned long long)krows, box_rows, (int)box_k);
    abort();
  }
  return m;
}
// every tensor map the launchers will build for this handle, encoded once at creation so that a driver that refuses one
// is reported through the ABI (bnr_create -> BNR_ECUDA) instead of at the first sweep
bool tmaps_check(const Engine& e) {
  const Dims& d = e.d;
  const uint64_t N = d.gdim, C = d.C, T = d.gdim / SY_BT;
  CUtensorMap m;
  bool ok = true;
  if (d.gmode != 2) ok = ok && encode_map(&m, e.X, d.np, d.qp, d.np, SY_LDS, SY_BK) == CUDA_SUCCESS;
  for (uint32_t box : {(uint32_t)SY_LDS, (uint32_t)(SY_BT / 2 + 4), (uint32_t)(SY_BT / 4 + 4)})
    ok = ok && encode_map(&m, e.G, N, N * C, N, box, SY_BK) == CUDA_SUCCESS;
  ok = ok && encode_map(&m, e.Linv, SY_BT, SY_BT * T * C, SY_BT, SY_LDS, SY_BK) == CUDA_SUCCESS;
  ok = ok && encode_map(&m, e.G, N, N * C, N, SY_BT, BW_COLS) == CUDA_SUCCESS;       // back solve
  ok = ok && encode_map(&m, e.Linv, SY_BT, SY_BT * T * C, SY_BT, SY_BT, BW_COLS) == CUDA_SUCCESS;
  return ok;
}

int syrk_splits(const Dims& d, int C) {
  const int T = d.np / SY_BT, tiles = T * (T + 1) / 2, nk = d.qp / SY_BK;
  if (const char* ov = getenv("BNR_SYRK_SPLITS")) {                       // tuning knob
    int s = atoi(ov);
    if (s > nk / 8) s = nk / 8;
    return s < 1 ? 1 : (s > SYRK_MAX_SPLITS ? SYRK_MAX_SPLITS : s);
  }
  // (measured: once chains x tiles reaches one wave, splitting only removes the natural stagger between chain groups)
  if (tiles * C >= 148) return 1;
  int s = (4 * 148 + tiles * C - 1) / (tiles * C);
  if (s > nk / 24) s = nk / 24;
  if (s > SYRK_MAX_SPLITS) s = SYRK_MAX_SPLITS;
  return s < 1 ? 1 : s;
}

void launch_syrk_G(const Engine& e, cudaStream_t s) {
  const Dims& d = e.d;
  const int T = d.np / SY_BT;
  const int nk = d.qp / SY_BK;
  const int ns = e.syrk_ws ? e.syrk_ws_cap + 1 : 1;   // fixed per handle: chain groups must not change the summation order
  const int nk_split = (nk + ns - 1) / ns;
#if defined(SYRK_LAB_ONLY_DIAG)
  dim3 grid(d.C, ns, T);
#elif defined(SYRK_LAB_ONLY_OFFDIAG)
  dim3 grid(d.C, ns, T * (T - 1) / 2);
#else
  dim3 grid(d.C, ns, T * (T + 1) / 2);
#endif
  const size_t gs = (size_t)d.np * d.np;
  const CUtensorMap tmX = make_map(e.X, d.np, d.qp, d.np, SY_LDS);
  ++g_launches; k_gram_syrk<<<grid, SY_THREADS, SYRK_SMEM, s>>>(tmX, e.S, (size_t)d.qp, e.G, gs, d.np, d.n, nk, 1.0,
                                                            nk_split, e.syrk_ws, e.syrk_ws_cap);
  if (ns > 1) {
    dim3 g2(T * (T + 1) / 2 * 8, d.C);
    ++g_launches; k_syrk_splitk_reduce<<<g2, 256, 0, s>>>(e.G, gs, d.np, e.syrk_ws, e.syrk_ws_cap, ns - 1);
  }
  if (e.aux.G_copy) {
    dim3 g2(d.np, d.C);
    ++g_launches; k_copy_sym<<<g2, 256, 0, s>>>(e.G, gs, d.np, e.aux.G_copy);
  }
}

// ------------------------------------------------------------------------------------------------------------
// q-form of the gamma draw: P_c = (X'X + diag(1/S_c)) / tau2_c  (q x q precision of gamma - W | rest).
// X'X is computed once per handle with the same DMMA SYRK (operand = X stored row-major, K = np, unit scales).
// ------------------------------------------------------------------------------------------------------------
// XtX (lower tiles of a qp x qp column-major matrix) from XT[i * qp + j] = X(i, j) (np rows, zero padded)
void launch_xtx(const Dims& d, const double* XT, const double* ones, double* XtX, cudaStream_t s) {
  const int T = d.qp / SY_BT;
  dim3 grid(1, 1, T * (T + 1) / 2);
  const CUtensorMap tmXT = make_map(XT, d.qp, d.np, d.qp, SY_LDS);
  ++g_launches; k_gram_syrk<<<grid, SY_THREADS, SYRK_SMEM, s>>>(tmXT, ones, 0, XtX, 0, d.qp, d.q, d.np / SY_BK, 0.0,
                                                            d.np / SY_BK, nullptr, 0);
}

// P_c lower triangle from XtX, S_c, tau2_c; padding rows/columns get the identity so the factorisation stays PD.
// grid = (qp/32, qp/32, C), block = (32, 8): tile (bi, bj) with bi >= bj only.
__global__ void __launch_bounds__(256) k_build_P(Engine e) {
  const Dims& d = e.d;
  if (blockIdx.x < blockIdx.y) return;
  const int c = blockIdx.z, N = d.qp;
  const int i = blockIdx.x * 32 + threadIdx.x;
  const double it2 = 1.0 / e.tau2[c];
  double* Gc = e.G + (size_t)c * N * N;
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    const int j = blockIdx.y * 32 + threadIdx.y * 4 + jj;
    if (i < j) continue;
    double v;
    if (i < d.q) {
      v = e.XtX[(size_t)j * N + i] * it2;
      if (i == j) v = (e.XtX[(size_t)j * N + i] + 1.0 / e.S[(size_t)c * d.qp + j]) * it2;
    } else {
      v = (i == j) ? 1.0 : 0.0;
    }
    Gc[(size_t)j * N + i] = v;
  }
}

void launch_build_P(const Engine& e, cudaStream_t s) {
  const Dims& d = e.d;
  dim3 grid(d.qp / 32, d.qp / 32, d.C), block(32, 8);
  ++g_launches; k_build_P<<<grid, block, 0, s>>>(e);
  if (e.aux.G_copy) {
    dim3 g2(d.qp, d.C);
    ++g_launches; k_copy_sym<<<g2, 256, 0, s>>>(e.G, (size_t)d.qp * d.qp, d.qp, e.aux.G_copy);
  }
}

// Factor every G_c (gdim x gdim) in place; the forward solve L w = rhs rides along as row m of the matrix.
// Two schedules of the same blocked algorithm, chosen from the handle's chain count (never from the chain grouping):
//
// (A) throughput schedule -- many chains, the SMs are busy anyway and total SM time is what counts.  Left-looking, one
//     deep update per block column; per panel J:
//       k_chol_update  tile (J, J) and, on `side`, tiles (i, J), i > J: all panels 0..J-1 at once (depth 128 J)
//       k_potf2_inv    factor + invert the diagonal block                                                  [main]
//       k_trsm_dmma    tiles (i, J), i > J                                                                 [main]
//
// (B) latency schedule -- few chains (chains x (T-1) <= 128; measured crossover between 16 and 32 chains at T = 8), the GPU is mostly idle and the length of the dependent
//     chain is what counts.  Left-looking with a one-panel look-ahead; per panel J, in program order (every tile sees its
//     updates in this order on any schedule):
//       PA(J)      k_potf2_inv    D_JJ -= L[J,J-1] L[J,J-1]' (late part, in the kernel), factor, invert       [main]
//       T1(J)      k_trsm_dmma    tile (J+1, J)                                                               [main]
//       T2(J)      k_trsm_dmma    tiles (i, J), i >= J+2                                                      [side]
//       Ulate(J)   k_chol_update  tiles (i, J+1), i >= J+2, contribution of panel J only (depth 128)          [side]
//       Uearly(J)  k_chol_update  tiles (i, J+2), i >= J+2 (diagonal included), panels 0..J (depth 128 (J+1)) [side]
//     i.e. block column j receives the panels 0..j-2 early (as soon as they exist, off the critical path) and panel j-1
//     late.  The critical path per panel is PA + T1; everything else runs on `side` next to the following panel
//     factorisation.  Tiles are cut into row strips (syrk_body) while the launch stays within one wave of CTAs.
//       dependencies across the two streams:  T2(J) after PA(J);  Ulate(J) after T1(J);
//                                             T1(J+1) after Ulate(J);  PA(J+2) after Uearly(J).
// `side` is a forked branch of the graph with the priority of the main stream.  Without a side stream the same kernels
// run in program order on `s` (identical results).
void launch_cholesky(const Engine& e, double* rhs, const ForkJoin& fj, cudaStream_t s) {
  const Dims& d = e.d;
  const int N = d.gdim;
  const int m = d.gmode == 2 ? d.q : d.n;         // first padding row = the bordering row
  const size_t cs = (size_t)N * N;
  const int T = N / PB;
  const size_t ls = (size_t)T * PB * PB;
  static const bool serial = getenv("BNR_CHOL_SERIAL") != nullptr;        // debugging knob: program order on one stream
  static const int forced = getenv("BNR_CHOL_SCHEDULE") ? atoi(getenv("BNR_CHOL_SCHEDULE")) : 0;   // 1 = (A), 2 = (B)
  const bool latency = forced ? forced == 2 : (long long)d.C_total * (T - 1) <= 128;
  ++g_launches; k_augment<<<d.C, 256, 0, s>>>(e.G, cs, N, m, rhs, d.gmode == 2 ? d.qp : d.np, d.gmode == 2 ? e.S : nullptr, e.tau2);
  // operand maps of the ring-fed kernels: the chains' matrices stacked along the k axis (chain c, column k -> k-row
  // c N + k), boxes as wide as a tile / a half / a quarter tile (+ 4 rows of padding); the stacked panel inverses
  const CUtensorMap mG = make_map(e.G, N, (uint64_t)N * d.C, N, SY_LDS);
  const CUtensorMap mG2 = make_map(e.G, N, (uint64_t)N * d.C, N, SY_BT / 2 + 4);
  const CUtensorMap mG4 = make_map(e.G, N, (uint64_t)N * d.C, N, SY_BT / 4 + 4);
  const CUtensorMap mL = make_map(e.Linv, PB, (uint64_t)PB * T * d.C, PB, SY_LDS);
  auto mI = [&](int ns) -> const CUtensorMap& { return ns == 4 ? mG4 : (ns == 2 ? mG2 : mG); };

  if (!latency) {
    // ---- (A) ----
    const bool can_fork = !serial && fj.side != nullptr;
    // The forked updates run at the panel chain's priority.  With another chain group's SYRK resident (600 us CTAs that
    // give an SM back every ~4 us) every dependent kernel has to re-acquire its SMs one by one; on the SYRK's own
    // priority level they queued behind ALL its pending CTAs and the factorisation only ran in the SYRK's last wave
    // (timeline in profiles/r02_timeline_c3.txt: 1 ms per sweep without any SYRK CTA resident).
    static const bool a_lo = getenv("BNR_CHOL_A_SIDE_LO") != nullptr;      // knob: the old low-priority branch
    cudaStream_t sideA = (!a_lo && fj.side_hi) ? fj.side_hi : fj.side;
    for (int J = 0; J < T; ++J) {
      const bool fork = can_fork && J > 0 && J + 1 < T;
      if (J > 0) {
        const int nk = J * PB / SY_BK;
        if (fork) {
          cudaEventRecord(fj.fork, s);
          cudaStreamWaitEvent(sideA, fj.fork, 0);
          dim3 g2(d.C, T - J - 1);
          ++g_launches; k_chol_update<<<g2, SY_THREADS, SYRK_SMEM, sideA>>>(mG, mG, 0, e.G, cs, N, m + 1, nk, J, J + 1, 1);
          cudaEventRecord(fj.join, sideA);
          dim3 g1(d.C, 1);
          ++g_launches; k_chol_update<<<g1, SY_THREADS, SYRK_SMEM, s>>>(mG, mG, 0, e.G, cs, N, m + 1, nk, J, J, 1);
        } else {
          dim3 g2(d.C, T - J);
          ++g_launches; k_chol_update<<<g2, SY_THREADS, SYRK_SMEM, s>>>(mG, mG, 0, e.G, cs, N, m + 1, nk, J, J, 1);
        }
      }
      ++g_launches; k_potf2_inv<<<d.C, 256, POTF2_SMEM, s>>>(e.G, cs, N, J, e.Linv, T, e.status, 0);
      if (fork) cudaStreamWaitEvent(s, fj.join, 0);
      if (J + 1 < T) {
        dim3 g1(d.C, T - J - 1);
        ++g_launches; k_trsm_dmma<<<g1, SY_THREADS, SYRK_SMEM, s>>>(mL, mG, e.G, cs, N, m + 1, J, J + 1, 1);
      }
    }
    return;
  }

  // ---- (B) ----
  const bool fork = !serial && fj.side_hi != nullptr && fj.pool != nullptr && fj.npool >= 4 * T;
  cudaStream_t side = fork ? fj.side_hi : s;
  cudaEvent_t* evP = fj.pool;                       // [T] each: after PA, after T1, after Ulate, after Uearly
  cudaEvent_t* evT = fj.pool + T;
  cudaEvent_t* evL = fj.pool + 2 * T;
  cudaEvent_t* evE = fj.pool + 3 * T;
  // row strips per tile (see syrk_body): as many as keep the launch within one wave of CTAs
  auto strips = [&](int tiles) {
    static const int fs = getenv("BNR_STRIPS") ? atoi(getenv("BNR_STRIPS")) : 0;      // tuning knob (1, 2 or 4)
    if (fs == 1 || fs == 2 || fs == 4) return fs;
    const int ctas = d.C_total * tiles;
    return ctas * 4 <= 148 ? 4 : (ctas * 2 <= 148 ? 2 : 1);
  };
  // panel solve of `rows` row blocks from ib_first of block column J: the shared-memory-resident kernel when the tiles
  // are cut into strips (few chains), else the ring-fed one
  static const bool small_tiles = getenv("BNR_NO_SMALL_TILES") == nullptr;
  auto launch_trsm = [&](int J, int ib_first, int rows, int ns, cudaStream_t st) {
    dim3 g(d.C, rows * ns);
    const double* Lj = e.Linv + (size_t)J * PB * PB;
    ++g_launches;
    if (small_tiles && ns == 8)
      k_small_tile<2, 2><<<g, 256, small_tile_smem(16), st>>>(e.G, cs, N, J, ib_first, 8, Lj, ls, PB, J * PB, 0);
    else if (small_tiles && ns == 4)
      k_small_tile<2, 4><<<g, 256, small_tile_smem(32), st>>>(e.G, cs, N, J, ib_first, 4, Lj, ls, PB, J * PB, 0);
    else if (small_tiles && ns == 2)
      k_small_tile<2, 8><<<g, 256, small_tile_smem(64), st>>>(e.G, cs, N, J, ib_first, 2, Lj, ls, PB, J * PB, 0);
    else
      k_trsm_dmma<<<g, SY_THREADS, SYRK_SMEM, st>>>(mL, mI(ns), e.G, cs, N, m + 1, J, ib_first, ns);
  };
  // The late part of the diagonal tile (J+1, J+1) -- the contribution of panel J -- either runs inside k_potf2_inv
  // (one SM: ~12 us of DMMA) or, when the chains are few enough for 4-8 CTAs per tile, as its own strip kernel Ud(J)
  // right after T1(J) (~5 us), with T1 itself cut into as many strips.
  static const int ud_knob = getenv("BNR_CHOL_UD") ? atoi(getenv("BNR_CHOL_UD")) : -1;     // tuning knob: 0 off, 4 / 8 strips
  // (measured: config 2 0.320 -> 0.300 ms, config 4 1.684 -> 1.666 ms; with a chain group's SYRK competing for the SMs
  //  -- 8 / 16 chains of config 3 -- the extra CTAs wait for SMs and the in-kernel update stays ahead: 1.770 vs 1.782 ms)
  const bool quiet = d.gmode == 2 || e.syrk_ws != nullptr;          // no full-size SYRK CTAs of other groups around
  int ud = (small_tiles && quiet) ? (d.C_total * 8 <= 148 ? 8 : (d.C_total * 4 <= 148 ? 4 : 0)) : 0;
  if (ud_knob == 0 || ud_knob == 4 || ud_knob == 8) ud = small_tiles ? ud_knob : 0;
  for (int J = 0; J < T; ++J) {
    const bool has_side = J + 2 < T;
    if (fork && J >= 2 && !ud) cudaStreamWaitEvent(s, evE[J - 2], 0);
    ++g_launches; k_potf2_inv<<<d.C, 256, POTF2_SMEM, s>>>(e.G, cs, N, J, e.Linv, T, e.status, (J > 0 && !ud) ? 1 : 0);
    if (fork && has_side) cudaEventRecord(evP[J], s);
    if (J + 1 < T) {
      if (fork && J >= 1 && J + 1 < T) cudaStreamWaitEvent(s, evL[J - 1], 0);
      const int ns1 = ud ? ud : strips(1);
      launch_trsm(J, J + 1, 1, ns1, s);
      if (fork && has_side) cudaEventRecord(evT[J], s);
      if (ud) {
        // Ud(J): tile (J+1, J+1) -= L[J+1, J] L[J+1, J]' (lower triangle), after the early panels 0..J-1 (Uearly(J-1))
        if (fork && J >= 1) cudaStreamWaitEvent(s, evE[J - 1], 0);
        dim3 gd(d.C, ud);
        const double* Lt = e.G + (size_t)J * PB * N + (size_t)(J + 1) * PB;
        ++g_launches;
        if (ud == 8) k_small_tile<1, 2><<<gd, 256, small_tile_smem(16), s>>>(e.G, cs, N, J + 1, J + 1, 8, Lt, cs, N, J * PB, 1);
        else k_small_tile<1, 4><<<gd, 256, small_tile_smem(32), s>>>(e.G, cs, N, J + 1, J + 1, 4, Lt, cs, N, J * PB, 1);
      }
    }
    if (has_side) {
      const int rows = T - J - 2;                   // row blocks J+2 .. T-1
      if (fork) cudaStreamWaitEvent(side, evP[J], 0);
      const int ns2 = strips(rows);
      dim3 g2(d.C, rows * ns2);
      launch_trsm(J, J + 2, rows, ns2, side);
      if (fork) cudaStreamWaitEvent(side, evT[J], 0);
      ++g_launches;
      if (small_tiles && ns2 == 4)
        k_small_tile<1, 4><<<g2, 256, small_tile_smem(32), side>>>(e.G, cs, N, J + 1, J + 2, 4, e.G + (size_t)J * PB * N + (size_t)(J + 1) * PB, cs, N, J * PB, 0);
      else if (small_tiles && ns2 == 2)
        k_small_tile<1, 8><<<g2, 256, small_tile_smem(64), side>>>(e.G, cs, N, J + 1, J + 2, 2, e.G + (size_t)J * PB * N + (size_t)(J + 1) * PB, cs, N, J * PB, 0);
      else
        k_chol_update<<<g2, SY_THREADS, SYRK_SMEM, side>>>(mG, mI(ns2), J * PB, e.G, cs, N, m + 1, PB / SY_BK, J + 1, J + 2, ns2);
      if (fork) cudaEventRecord(evL[J], side);
      ++g_launches; k_chol_update<<<g2, SY_THREADS, SYRK_SMEM, side>>>(mG, mI(ns2), 0, e.G, cs, N, m + 1, (J + 1) * PB / SY_BK,
                                                                   J + 2, J + 2, ns2);
      if (fork) cudaEventRecord(evE[J], side);
    }
  }
}

// out_c <- L_c^-T (w_c + addz_c), w = L^-1 rhs = row m of the factor (launch_cholesky)
void launch_chol_solve(const Engine& e, double* out, const double* addz, cudaStream_t s) {
  const Dims& d = e.d;
  const int N = d.gdim;
  const int m = d.gmode == 2 ? d.q : d.n;
  const int stride = d.gmode == 2 ? d.qp : d.np;
  const CUtensorMap tmG = make_map(e.G, N, (uint64_t)N * d.C, N, PB, BW_COLS);
  const CUtensorMap tmL = make_map(e.Linv, PB, (uint64_t)PB * (N / PB) * d.C, PB, PB, BW_COLS);
  ++g_launches; k_bwd_stream<<<d.C, BW_THREADS, bwd_smem(N), s>>>(tmG, tmL, e.G, (size_t)N * N, N, m, out, stride, addz, stride,
                                                               bwd_stages(N));
}

int chol_max_dim() { return CHOL_MAX_DIM; }

}  // namespace bnr
