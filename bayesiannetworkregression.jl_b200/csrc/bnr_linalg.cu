// Dense FP64 linear algebra of the gamma draw (update_gamma!, src/gibbs.jl:420-438), batched over chains:
//   G_c = X diag(S_c) X' + I          -> k_gram_syrk : FP64 tensor-core (DMMA m8n8k4) SYRK, cp.async 3-stage pipeline
//   G_c = L_c L_c', w = L_c^-1 rhs    -> blocked left-looking Cholesky: k_chol_update (DMMA) / k_potf2_128 / k_trsm_128
//   a4  = L_c^-T w                    -> k_trsv_bwd128
//   X v, X' a4 (all chains at once)   -> k_x_times (tall-skinny GEMM, deterministic split-K)
// tcgen05 has no FP64 kind, so the Blackwell tensor path for this contraction is the warp-level DMMA.
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
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

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
// contiguous global -> shared copy by the TMA engine, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------------------------------------------------
// SYRK on FP64 tensor cores.
//   MODE 0:  C_c[i][j] = sum_k A[i][k] s_c[k] A[j][k] + (i==j)      A = X (shared by all chains), K = qp
//   MODE 1:  C_c[i][j] -= sum_{k < nk*16} L_c[i][k] L_c[j][k]        left-looking Cholesky update of block column J
// The operand is "k-major": element (row, k) at A[row + ld*k] (rows contiguous) - exactly how X (column-major
// n x q) and a column panel of the column-major G are stored, so a 128-row x 16-k tile is sixteen contiguous
// 1 KB rows.  A dedicated producer warp streams them into a 4-stage shared-memory ring with TMA-engine bulk
// copies (cp.async.bulk, completion on "full" mbarriers); the 8 consumer warps never execute a block-wide
// barrier in the main loop - they wait on "full", issue DMMA m8n8k4, and release the stage on "empty".
// CTA tile 128 x 128, k-step 16.  The MMA "M" dimension runs along j (columns of C) and "N" along i (rows of C), so
// every accumulator pair is two consecutive rows of one column of the column-major C -> 16-byte stores.  Strictly
// lower tiles use a column-strip warp layout (syrk_strip_tile: 2 column fragments x all row fragments per warp),
// diagonal tiles compute only their lower triangle, balanced over the warps (syrk_diag_tile).  Shared rows are
// padded to 132 doubles (== 4 mod 16) which makes the (row, k) fragment loads bank-conflict free.
// The per-chain scale s_c[k] is applied to the operand with the fewest fragments per warp (2): DMUL shares the
// FP64 pipe with DMMA.  Work that cannot contribute is skipped: the symmetric half of diagonal tiles and row
// fragments that lie entirely in the zero padding beyond n.
// grid = (C, tiles), block = 384 (2 consumer warpgroups + 1 producer warpgroup, registers re-balanced with
// setmaxnreg: 232 per consumer thread, 40 per producer thread); dynamic smem = SYRK_SMEM.
// ------------------------------------------------------------------------------------------------------------
constexpr int SY_BT = 128;        // tile edge
constexpr int SY_BK = 16;         // k-step
constexpr int SY_LDS = 132;       // smem row stride in doubles (== 4 mod 16 -> conflict-free fragment loads)
constexpr int SY_STAGES = 4;
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
__device__ __forceinline__ void syrk_diag_frags(const double* __restrict__ sj, const double* __restrict__ ss, int k4,
                                                int lk, int lr, double (&a2)[2], double (&bfr)[16 - W]) {
  const int kr = k4 * 4 + lk;
  const double* row = sj + kr * SY_LDS + lr;
  // the scale rides on the two column-operand fragments (2 DMULs) rather than on the 16 - W row-operand fragments:
  // DMUL shares the FP64 pipe with DMMA, so every multiply saved is tensor throughput gained
  if (MODE == 0) {
    const double sk = ss[kr];
    a2[0] = row[W * 8] * sk;
    a2[1] = row[(15 - W) * 8] * sk;
  } else {
    a2[0] = row[W * 8];
    a2[1] = row[(15 - W) * 8];
  }
#pragma unroll
  for (int i = 0; i < 16 - W; ++i) bfr[i] = row[(W + i) * 8];
}

template <int MODE, int W>
__device__ __forceinline__ void syrk_diag_tile(const double* __restrict__ smem, unsigned long long* full,
                                               unsigned long long* empty, int nk, int lk, int lr, int lane,
                                               double* __restrict__ Cc, int np, int i0, double diag_add) {
  constexpr int NA = 16 - W;        // fragments of column W
  double acc[17][2];
#pragma unroll
  for (int t = 0; t < 17; ++t) acc[t][0] = acc[t][1] = 0.0;
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
      v = *p;
      v.x -= acc[t][0];
      v.y -= acc[t][1];
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
__device__ __forceinline__ void syrk_strip_frags(const double* __restrict__ sj, const double* __restrict__ si,
                                                 const double* __restrict__ ss, int k4, int warp, int lk, int lr,
                                                 double (&af)[2], double (&bf)[NF]) {
  const int kr = k4 * 4 + lk;
  const double* rj = sj + kr * SY_LDS + warp * 16 + lr;
  const double* ri = si + kr * SY_LDS + lr;
  if (MODE == 0) {
    const double sk = ss[kr];
    af[0] = rj[0] * sk;
    af[1] = rj[8] * sk;
  } else {
    af[0] = rj[0];
    af[1] = rj[8];
  }
#pragma unroll
  for (int nf = 0; nf < NF; ++nf) bf[nf] = ri[nf * 8];
}

template <int MODE, int NF>
__device__ __forceinline__ void syrk_strip_tile(const double* __restrict__ smem, unsigned long long* full,
                                                unsigned long long* empty, int nk, int warp, int lk, int lr, int lane,
                                                double* __restrict__ Cc, int np, int i0, int j0) {
  double acc[2][NF][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < NF; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
  double af[2][2], bf[2][NF];
  mbar_wait(&full[0], 0);
  const double* sj = smem;
  syrk_strip_frags<MODE, NF>(sj, sj + SY_BK * SY_LDS, sj + 2 * SY_BK * SY_LDS, 0, warp, lk, lr, af[0], bf[0]);
  for (int kt = 0; kt < nk; ++kt) {
#pragma unroll
    for (int k4 = 0; k4 < SY_BK / 4; ++k4) {
      const int cur = k4 & 1, nxt = cur ^ 1;
      if (k4 + 1 < SY_BK / 4) {
        syrk_strip_frags<MODE, NF>(sj, sj + SY_BK * SY_LDS, sj + 2 * SY_BK * SY_LDS, k4 + 1, warp, lk, lr, af[nxt], bf[nxt]);
      } else if (kt + 1 < nk) {
        mbar_wait(&full[(kt + 1) % SY_STAGES], ((kt + 1) / SY_STAGES) & 1);
        const double* nj = smem + (size_t)((kt + 1) % SY_STAGES) * SY_STAGE_DBL;
        syrk_strip_frags<MODE, NF>(nj, nj + SY_BK * SY_LDS, nj + 2 * SY_BK * SY_LDS, 0, warp, lk, lr, af[nxt], bf[nxt]);
      }
#pragma unroll
      for (int nf = 0; nf < NF; ++nf) {
        dmma884(acc[0][nf][0], acc[0][nf][1], af[cur][0], bf[cur][nf]);
        dmma884(acc[1][nf][0], acc[1][nf][1], af[cur][1], bf[cur][nf]);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[kt % SY_STAGES]);
    sj = smem + (size_t)((kt + 1) % SY_STAGES) * SY_STAGE_DBL;
  }
#pragma unroll
  for (int mf = 0; mf < 2; ++mf) {
    const int j = j0 + warp * 16 + mf * 8 + lr;
#pragma unroll
    for (int nf = 0; nf < NF; ++nf) {
      const int i = i0 + nf * 8 + 2 * lk;
      double2* p = reinterpret_cast<double2*>(Cc + (size_t)j * np + i);
      double2 v;
      if (MODE == 0) {
        v.x = acc[mf][nf][0];
        v.y = acc[mf][nf][1];
      } else {
        v = *p;
        v.x -= acc[mf][nf][0];
        v.y -= acc[mf][nf][1];
      }
      *p = v;
    }
  }
}

template <int MODE>
__device__ __forceinline__ void
syrk_body(const double* __restrict__ A, size_t a_chain_stride, int ld, const double* __restrict__ scale,
          size_t scale_stride, double* __restrict__ Cm, size_t c_chain_stride, int np, int nvalid, int nk, int origin,
          double diag_add) {
  extern __shared__ __align__(16) double smem[];
  unsigned long long* full = reinterpret_cast<unsigned long long*>(smem + (size_t)SY_STAGES * SY_STAGE_DBL);
  unsigned long long* empty = full + SY_STAGES;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = blockIdx.x;          // chains vary fastest: the 64 CTAs that share one X tile run back to back (L2 reuse)
  int ib, jb;
  if (MODE == 0) {
    // blockIdx.y enumerates the lower-triangular tiles, the strictly-lower (full-cost) ones first and the
    // diagonal (half-cost) ones last, so the tail of the grid is made of cheap CTAs
    const int T = np / SY_BT, noff = T * (T - 1) / 2;
#ifdef SYRK_LAB_ONLY_DIAG
    const int t = blockIdx.y + noff;
#else
    const int t = blockIdx.y;
#endif
    if (t < noff) {
      ib = (int)((1.0 + sqrt(1.0 + 8.0 * t)) * 0.5);
      while (ib * (ib - 1) / 2 > t) --ib;
      while ((ib + 1) * ib / 2 <= t) ++ib;
      jb = t - ib * (ib - 1) / 2;
    } else {
      ib = jb = t - noff;
    }
  } else {
    // left-looking update of block column J = origin: tiles (J+1 .. T-1, J) first, the diagonal tile (J, J) last
    const int T = np / SY_BT, nt = T - origin;
    jb = origin;
    ib = (blockIdx.y == nt - 1) ? origin : origin + 1 + blockIdx.y;
  }
  const int i0 = ib * SY_BT, j0 = jb * SY_BT;
  const bool diag = (ib == jb);
  const double* Ac = A + (size_t)c * a_chain_stride;
  const double* sc = (MODE == 0) ? scale + (size_t)c * scale_stride : nullptr;

  if (tid == 0) {
    for (int s = 0; s < SY_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if (warp >= 8) {
    // ---------------- producer warpgroup: hands its registers to the consumers, one lane drives the TMA engine ----
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(SY_PRODUCER_REGS));
    if (warp == 8 && lane == 0) {
      const unsigned bytes = (diag ? 1u : 2u) * SY_BK * SY_BT * 8u + (MODE == 0 ? SY_BK * 8u : 0u);
      for (int kt = 0; kt < nk; ++kt) {
        const int stage = kt % SY_STAGES;
        const unsigned ph = (kt / SY_STAGES) & 1;
        mbar_wait(&empty[stage], ph ^ 1);
        double* sj = smem + (size_t)stage * SY_STAGE_DBL;
        double* si = sj + SY_BK * SY_LDS;
        mbar_expect_tx(&full[stage], bytes);
        const double* g = Ac + (size_t)kt * SY_BK * ld;
#pragma unroll 4
        for (int kr = 0; kr < SY_BK; ++kr) {
          bulk_g2s(sj + kr * SY_LDS, g + (size_t)kr * ld + j0, SY_BT * 8, &full[stage]);
          if (!diag) bulk_g2s(si + kr * SY_LDS, g + (size_t)kr * ld + i0, SY_BT * 8, &full[stage]);
        }
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
  nfv = nfv > 16 ? 16 : nfv;
  switch (nfv) {
#define BNR_STRIP_CASE(NF) case NF: syrk_strip_tile<MODE, NF>(smem, full, empty, nk, warp, lk, lr, lane, Cc, np, i0, j0); break;
    BNR_STRIP_CASE(1) BNR_STRIP_CASE(2) BNR_STRIP_CASE(3) BNR_STRIP_CASE(4) BNR_STRIP_CASE(5) BNR_STRIP_CASE(6)
    BNR_STRIP_CASE(7) BNR_STRIP_CASE(8) BNR_STRIP_CASE(9) BNR_STRIP_CASE(10) BNR_STRIP_CASE(11) BNR_STRIP_CASE(12)
    BNR_STRIP_CASE(13) BNR_STRIP_CASE(14) BNR_STRIP_CASE(15)
#undef BNR_STRIP_CASE
    default: syrk_strip_tile<MODE, 16>(smem, full, empty, nk, warp, lk, lr, lane, Cc, np, i0, j0); break;
  }
}

__global__ void __launch_bounds__(SY_THREADS, 1)
k_gram_syrk(const double* __restrict__ X, int ld, const double* __restrict__ scale, size_t scale_stride,
            double* __restrict__ G, size_t g_chain_stride, int np, int nvalid, int nk, double diag_add) {
  syrk_body<0>(X, 0, ld, scale, scale_stride, G, g_chain_stride, np, nvalid, nk, 0, diag_add);
}

__global__ void __launch_bounds__(SY_THREADS, 1)
k_chol_update(const double* __restrict__ P, size_t chain_stride, int ld, double* __restrict__ G, int np, int nvalid,
              int nk, int origin) {
  syrk_body<1>(P, chain_stride, ld, nullptr, 0, G, chain_stride, np, nvalid, nk, origin, 0.0);
}

// ------------------------------------------------------------------------------------------------------------
// Blocked left-looking Cholesky, block size 128, with the forward solve L w = rhs folded in.
//   for J = 0 .. np/128-1:
//     k_chol_update   (DMMA, above):  G[J.., J] -= L[J.., 0:J] L[J, 0:J]'
//     k_potf2_128     one CTA per chain: factor the 128 x 128 diagonal block in shared
//                     memory (4 sub-blocks of 32: a warp factors 32 x 32 in registers with shuffles, threads eliminate
//                     the rows below, everybody updates the trailing part in 2 x 2 register tiles); rhs_J rides along as
//                     row 128, so w_J = L_JJ^-1 rhs_J comes out of the same elimination.  1/L_jj goes to `dinv`.
//     k_trsm_128      rows below the diagonal block: L[i, J] = G[i, J] L_JJ^-T, one thread per row, two 64-column
//                     halves, right-looking inside the thread (independent FMAs, no divisions); the same thread then
//                     subtracts its row's share of the forward solve: rhs[i] -= L[i, J] w_J.
// All inner loops are arranged so that consecutive FP64 FMAs are independent: these kernels are latency-bound.
// ------------------------------------------------------------------------------------------------------------
constexpr int PB = 128;            // panel / diagonal block size
constexpr int PB_LD = PB + 1;      // shared-memory row stride of the (PB+1) x PB working block
constexpr size_t POTF2_SMEM = sizeof(double) * ((size_t)(PB + 1) * PB_LD + 32);

__global__ void __launch_bounds__(256) k_potf2_128(double* __restrict__ G, size_t chain_stride, int np, int J,
                                                   double* __restrict__ rhs, double* __restrict__ dinv_out,
                                                   int* status) {
  extern __shared__ double sm[];
  double* A = sm;                          // A[r][c] at A[r * PB_LD + c], rows 0..128 (row 128 = rhs), cols 0..127
  double* dinv = sm + (PB + 1) * PB_LD;    // [32] reciprocal diagonal of the current 32-block
  const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* Gc = G + (size_t)c * chain_stride;
  double* D = Gc + (size_t)J * PB * np + (size_t)J * PB;
  double* rc = rhs + (size_t)c * np;
  const int kprev = J * PB;
  for (int id = tid; id < PB * PB; id += 256) {
    const int r = id & (PB - 1), cc = id >> 7;
    A[r * PB_LD + cc] = (r >= cc) ? D[(size_t)cc * np + r] : 0.0;
  }
  // rhs rides along as row 128 of the working block; contributions of earlier panels were already subtracted
  // by k_trsm_128 (right-looking forward solve), so w_J = L_JJ^-1 rhs_J falls out of the elimination below
  if (tid < PB) A[PB * PB_LD + tid] = rc[kprev + tid];
  __syncthreads();
  bool bad = false;
  for (int s = 0; s < PB / 32; ++s) {
    const int o = s * 32;
    if (warp == 0) {
      // 32 x 32 Cholesky in registers: lane r holds row o+r (columns o .. o+31)
      double a[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) a[j] = A[(o + lane) * PB_LD + o + j];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const double djj = __shfl_sync(0xffffffffu, a[j], j);
        if (!(djj > 0.0)) bad = true;
        const double inv = rsqrt(djj);
        const double lj = (lane == j) ? djj * inv : a[j] * inv;
        a[j] = lj;
        if (lane == j) dinv[j] = inv;
#pragma unroll
        for (int cc = j + 1; cc < 32; ++cc) {
          const double lcj = __shfl_sync(0xffffffffu, lj, cc);
          a[cc] -= lj * lcj;                 // lanes < cc compute garbage in their (unused) upper part
        }
      }
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (lane >= j) A[(o + lane) * PB_LD + o + j] = a[j];
    }
    __syncthreads();
    // rows below (incl. the rhs row 128): x L_ss' = a, thread per row, right-looking (independent FMAs)
    const int nbelow = PB + 1 - (o + 32);
    if (tid < nbelow) {
      double* row = A + (o + 32 + tid) * PB_LD + o;
      double x[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = row[j];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const double xj = x[j] * dinv[j];
        x[j] = xj;
#pragma unroll
        for (int k = j + 1; k < 32; ++k) x[k] -= xj * A[(o + k) * PB_LD + o + j];
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) row[j] = x[j];
    }
    __syncthreads();
    // trailing update inside the block in 2 x 2 register tiles:
    //   A[r][cc] -= sum_p L[r][o+p] L[cc][o+p],  o+32 <= cc <= r <= 128 (cc < 128)
    const int base = o + 32;
    const int nr2 = (PB + 1 - base + 1) / 2;     // row pairs (rows base .. 128)
    const int nc2 = (PB - base) / 2;             // column pairs (cols base .. 127), even count
    for (int id = tid; id < nr2 * nc2; id += 256) {
      const int r = base + 2 * (id / nc2), cc = base + 2 * (id % nc2);
      if (cc > r + 1) continue;
      const int r1 = (r + 1 <= PB) ? r + 1 : r;  // clamp the phantom row 129
      const double* La = A + r * PB_LD + o;
      const double* Lb = A + r1 * PB_LD + o;
      const double* Lc = A + cc * PB_LD + o;
      const double* Ld = A + (cc + 1) * PB_LD + o;
      double s00 = 0.0, s01 = 0.0, s10 = 0.0, s11 = 0.0;
#pragma unroll
      for (int p = 0; p < 32; ++p) {
        const double la = La[p], lb = Lb[p], lc = Lc[p], ld = Ld[p];
        s00 += la * lc; s01 += la * ld; s10 += lb * lc; s11 += lb * ld;
      }
      A[r * PB_LD + cc] -= s00;
      if (cc + 1 <= r) A[r * PB_LD + cc + 1] -= s01;
      if (r1 != r) {
        A[r1 * PB_LD + cc] -= s10;
        A[r1 * PB_LD + cc + 1] -= s11;
      }
    }
    __syncthreads();
    if (tid < 32) dinv_out[(size_t)c * np + kprev + o + tid] = dinv[tid];
  }
  if (bad) atomicOr(&status[c], BNR_ST_G_NOTPD_);
  for (int id = tid; id < PB * PB; id += 256) {
    const int r = id & (PB - 1), cc = id >> 7;
    if (r >= cc) D[(size_t)cc * np + r] = A[r * PB_LD + cc];
  }
  if (tid < PB) rc[kprev + tid] = A[PB * PB_LD + tid];
}

// rows below the diagonal block.  grid = (rows_below / 128, C), block = 128.
// shared: Lt11, Lt21, Lt22 as [p][j] (column p of the 64 x 64 sub-block contiguous in j, read as 16-byte broadcasts),
// 128 reciprocal diagonals and w_J.
constexpr size_t TRSM_SMEM = sizeof(double) * (3 * 64 * 64 + 256);

// x[k] -= xj * l[k] for k = K0 .. 63 with 16-byte shared loads (K0 is a compile-time constant)
template <int K0>
__device__ __forceinline__ void axpy_tail(double (&x)[64], double xj, const double* __restrict__ l) {
  constexpr int KE = (K0 + 1) & ~1;                    // first even index >= K0
  if (K0 & 1) x[K0] -= xj * l[K0];
  const double2* lv = reinterpret_cast<const double2*>(l);
#pragma unroll
  for (int k2 = KE / 2; k2 < 32; ++k2) {
    const double2 v = lv[k2];
    x[2 * k2] -= xj * v.x;
    x[2 * k2 + 1] -= xj * v.y;
  }
}

template <int J0>
__device__ __forceinline__ void trsm_solve64(double (&x)[64], const double* __restrict__ Lt, const double* __restrict__ dinv) {
  if constexpr (J0 < 64) {
    const double xj = x[J0] * dinv[J0];
    x[J0] = xj;
    if constexpr (J0 + 1 < 64) axpy_tail<J0 + 1>(x, xj, Lt + J0 * 64);
    trsm_solve64<J0 + 1>(x, Lt, dinv);
  }
}

__global__ void __launch_bounds__(128) k_trsm_128(double* __restrict__ G, size_t chain_stride, int np, int J,
                                                  const double* __restrict__ dinv_g, double* __restrict__ rhs) {
  extern __shared__ __align__(16) double sm[];
  double* L11 = sm;                 // [p][j] = L11[j][p]
  double* L21 = sm + 64 * 64;       // [p][j] = L21[j][p]   (j: second-half column, p: first-half column)
  double* L22 = sm + 2 * 64 * 64;
  double* dinv = sm + 3 * 64 * 64;  // [128]
  double* wJ = dinv + 128;          // [128]
  const int c = blockIdx.y, tid = threadIdx.x;
  double* Gc = G + (size_t)c * chain_stride;
  const double* D = Gc + (size_t)J * PB * np + (size_t)J * PB;
  for (int id = tid; id < 64 * 64; id += 128) {
    const int r = id & 63, cc = id >> 6;          // D(r, cc) column-major: coalesced in r
    L11[cc * 64 + r] = D[(size_t)cc * np + r];
    L21[cc * 64 + r] = D[(size_t)cc * np + 64 + r];
    L22[cc * 64 + r] = D[(size_t)(64 + cc) * np + 64 + r];
  }
  dinv[tid] = dinv_g[(size_t)c * np + J * PB + tid];
  wJ[tid] = rhs[(size_t)c * np + J * PB + tid];
  __syncthreads();
  const int row = (J + 1) * PB + blockIdx.x * 128 + tid;
  if (row >= np) return;
  double* prow = Gc + (size_t)J * PB * np + row;
  double* prow2 = prow + (size_t)64 * np;
  double x[64];
#pragma unroll
  for (int j = 0; j < 64; ++j) x[j] = prow[(size_t)j * np];
  trsm_solve64<0>(x, L11, dinv);            // first half: x_j = a_j / L11[j][j]; a_k -= x_j L11[k][j] (k > j)
  double racc = 0.0;
#pragma unroll
  for (int j = 0; j < 64; ++j) { prow[(size_t)j * np] = x[j]; racc += x[j] * wJ[j]; }
  // second half, 16 columns at a time: y_k = A[row][64 + k] - sum_p x_p L21[k][p]
#pragma unroll 1
  for (int k0 = 0; k0 < 64; k0 += 16) {
    double y[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) y[k] = prow2[(size_t)(k0 + k) * np];
#pragma unroll
    for (int p = 0; p < 64; ++p) {
      const double2* l2 = reinterpret_cast<const double2*>(L21 + p * 64 + k0);
#pragma unroll
      for (int k2 = 0; k2 < 8; ++k2) {
        const double2 v = l2[k2];
        y[2 * k2] -= x[p] * v.x;
        y[2 * k2 + 1] -= x[p] * v.y;
      }
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) prow2[(size_t)(k0 + k) * np] = y[k];
  }
#pragma unroll
  for (int j = 0; j < 64; ++j) x[j] = prow2[(size_t)j * np];
  trsm_solve64<0>(x, L22, dinv + 64);
#pragma unroll
  for (int j = 0; j < 64; ++j) { prow2[(size_t)j * np] = x[j]; racc += x[j] * wJ[64 + j]; }
  // right-looking forward solve: this row's right-hand side loses the contribution of panel J
  rhs[(size_t)c * np + row] -= racc;
}

// ------------------------------------------------------------------------------------------------------------
// backward solve  L' x = w  (w in rhs, overwritten by x), left-looking over 128-blocks from the bottom.
// grid = C, block = 256 (8 warps): warp per column group for the matvec with the rows below (coalesced, 8 columns in
// flight per warp), then a 128 x 128 transposed triangular solve by one warp (shuffles, reciprocal diagonals).
// ------------------------------------------------------------------------------------------------------------
static size_t trsvb_smem(int np) { return sizeof(double) * ((size_t)PB * PB_LD + np + 2 * PB); }
constexpr int CHOL_MAX_DIM = 4096;   // largest factored dimension (shared-memory solution vector of the back solve)

// addz (optional, [C][np]): added to the right-hand side before the solve - the q-form draws L' beta = w + z
__global__ void __launch_bounds__(256) k_trsv_bwd128(const double* __restrict__ G, size_t chain_stride, int np,
                                                     double* __restrict__ rhs, const double* __restrict__ dinv_g,
                                                     const double* __restrict__ addz) {
  extern __shared__ double sm[];
  double* Ls = sm;                     // [r][c] lower block
  double* x = sm + PB * PB_LD;         // [np] solution so far (entries >= (J+1)*128 valid)
  double* b = x + np;                  // [128] current right-hand side
  double* dinv = b + PB;               // [128]
  const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* Gc = G + (size_t)c * chain_stride;
  double* rc = rhs + (size_t)c * np;
  const int T = np / PB;
  for (int J = T - 1; J >= 0; --J) {
    const int r0 = (J + 1) * PB, nrow = np - r0;
    const double* D = Gc + (size_t)J * PB * np + (size_t)J * PB;
    for (int id = tid; id < PB * PB; id += 256) {
      const int r = id & (PB - 1), cc = id >> 7;
      Ls[r * PB_LD + cc] = (r >= cc) ? D[(size_t)cc * np + r] : 0.0;
    }
    if (tid < PB) dinv[tid] = dinv_g[(size_t)c * np + J * PB + tid];
    // b_j = w_j - sum_{i >= r0} L[i][J*128 + j] x_i : warp w owns columns 16w .. 16w+15, all 16 in flight
    {
      const int jj = warp * 16;
      double acc[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) acc[u] = 0.0;
      const double* col = Gc + (size_t)(J * PB + jj) * np + r0;
      for (int i = lane; i < nrow; i += 32) {
        const double xi = x[r0 + i];
#pragma unroll
        for (int u = 0; u < 16; ++u) acc[u] += col[(size_t)u * np + i] * xi;
      }
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        double v = acc[u];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) {
          const int idx = J * PB + jj + u;
          b[jj + u] = rc[idx] + (addz ? addz[(size_t)c * np + idx] : 0.0) - v;
        }
      }
    }
    __syncthreads();
    // L_JJ' x_J = b : columns from the right; one warp, lane owns entries lane, lane+32, lane+64, lane+96
    if (warp == 0) {
      double v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = b[lane + 32 * u];
#pragma unroll
      for (int u = 3; u >= 0; --u) {
        for (int src = 31; src >= 0; --src) {
          const int j = 32 * u + src;
          const double xj = __shfl_sync(0xffffffffu, v[u], src) * dinv[j];
          const double* Lj = Ls + j * PB_LD;
          if (lane == src) v[u] = xj;
          else if (lane < src) v[u] -= Lj[lane + 32 * u] * xj;
#pragma unroll
          for (int uu = 0; uu < 4; ++uu)
            if (uu < u) v[uu] -= Lj[lane + 32 * uu] * xj;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { x[J * PB + lane + 32 * u] = v[u]; rc[J * PB + lane + 32 * u] = v[u]; }
    }
    __syncthreads();
  }
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
// tall-skinny GEMM with X for all chains at once: out[c][m] = sum_k A(m,k) in[c][k]
//   TRANS = 0: A(m,k) = X[m + np*k]  (M = np, K = qp)      TRANS = 1: A(m,k) = X[k + np*m]  (M = qp, K = np)
// 64 x 32 x 16 tiles, 256 threads, 4 x 2 per thread, deterministic split-K through a workspace.
// ------------------------------------------------------------------------------------------------------------
constexpr int XT_BM = 64, XT_BN = 32, XT_BK = 16;

template <int TRANS>
__global__ void __launch_bounds__(256) k_x_times(const double* __restrict__ X, int np, int M, int K, int N,
                                                 const double* __restrict__ in, int ldin, double* __restrict__ out,
                                                 int ldout, int k_per_split) {
  __shared__ double As[XT_BK][XT_BM + 4];
  __shared__ double Bs[XT_BK][XT_BN + 2];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * XT_BM, n0 = blockIdx.y * XT_BN;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);
  double* o = out + (size_t)blockIdx.z * N * ldout;
  const int tx = tid & 15, ty = tid >> 4;
  double acc[4][2] = {};
  for (int k0 = kbeg; k0 < kend; k0 += XT_BK) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int id = tid + 256 * r;
      if (TRANS == 0) {
        const int kk = id >> 6, mm = id & 63;
        As[kk][mm] = (m0 + mm < M) ? X[(size_t)(k0 + kk) * np + m0 + mm] : 0.0;
      } else {
        const int mm = id >> 4, kk = id & 15;
        As[kk][mm] = (m0 + mm < M) ? X[(size_t)(m0 + mm) * np + k0 + kk] : 0.0;
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int id = tid + 256 * r;
      const int nn = id >> 4, kk = id & 15;
      Bs[kk][nn] = (n0 + nn < N) ? in[(size_t)(n0 + nn) * ldin + k0 + kk] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < XT_BK; ++kk) {
      double a[4], b[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][tx * 4 + i];
      b[0] = Bs[kk][ty * 2]; b[1] = Bs[kk][ty * 2 + 1];
#pragma unroll
      for (int i = 0; i < 4; ++i) { acc[i][0] += a[i] * b[0]; acc[i][1] += a[i] * b[1]; }
    }
    __syncthreads();
  }
#pragma unroll
  for (int jn = 0; jn < 2; ++jn) {
    const int nn = n0 + ty * 2 + jn;
    if (nn >= N) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int mm = m0 + tx * 4 + i;
      if (mm < M) o[(size_t)nn * ldout + mm] = acc[i][jn];
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
  const int tiles = ((M + XT_BM - 1) / XT_BM) * ((d.C + XT_BN - 1) / XT_BN);
  int ks = (2 * 148 + tiles - 1) / tiles;
  const int kmax = K / (4 * XT_BK) > 0 ? K / (4 * XT_BK) : 1;
  if (ks > kmax) ks = kmax;
  if (ks > 32) ks = 32;
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
  int kper = ((K + ks - 1) / ks + XT_BK - 1) / XT_BK * XT_BK;
  dim3 grid((M + XT_BM - 1) / XT_BM, (N + XT_BN - 1) / XT_BN, ks);
  double* dst = ks == 1 ? out : ws;
  if (trans) k_x_times<1><<<grid, 256, 0, s>>>(e.X, d.np, M, K, N, in, ldin, dst, ldout, kper);
  else k_x_times<0><<<grid, 256, 0, s>>>(e.X, d.np, M, K, N, in, ldin, dst, ldout, kper);
  if (ks > 1) {
    const size_t count = (size_t)N * ldout;
    ++g_launches; k_splitk_reduce<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(ws, out, count, ks);
  }
}

void linalg_setup() {
  cudaFuncSetAttribute(k_gram_syrk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SYRK_SMEM);
  cudaFuncSetAttribute(k_chol_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SYRK_SMEM);
  cudaFuncSetAttribute(k_potf2_128, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POTF2_SMEM);
  cudaFuncSetAttribute(k_trsm_128, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRSM_SMEM);
  cudaFuncSetAttribute(k_trsv_bwd128, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)trsvb_smem(CHOL_MAX_DIM));
}

void launch_syrk_G(const Engine& e, cudaStream_t s) {
  const Dims& d = e.d;
  const int T = d.np / SY_BT;
#if defined(SYRK_LAB_ONLY_DIAG)
  dim3 grid(d.C, T);
#elif defined(SYRK_LAB_ONLY_OFFDIAG)
  dim3 grid(d.C, T * (T - 1) / 2);
#else
  dim3 grid(d.C, T * (T + 1) / 2);
#endif
  ++g_launches; k_gram_syrk<<<grid, SY_THREADS, SYRK_SMEM, s>>>(e.X, d.np, e.S, (size_t)d.qp, e.G, (size_t)d.np * d.np, d.np, d.n,
                                                            d.qp / SY_BK, 1.0);
  if (e.aux.G_copy) {
    dim3 g2(d.np, d.C);
    ++g_launches; k_copy_sym<<<g2, 256, 0, s>>>(e.G, (size_t)d.np * d.np, d.np, e.aux.G_copy);
  }
}

// ------------------------------------------------------------------------------------------------------------
// q-form of the gamma draw: P_c = (X'X + diag(1/S_c)) / tau2_c  (q x q precision of gamma - W | rest).
// X'X is computed once per handle with the same DMMA SYRK (operand = X stored row-major, K = np, unit scales).
// ------------------------------------------------------------------------------------------------------------
// XtX (lower tiles of a qp x qp column-major matrix) from XT[i * qp + j] = X(i, j) (np rows, zero padded)
void launch_xtx(const Dims& d, const double* XT, const double* ones, double* XtX, cudaStream_t s) {
  const int T = d.qp / SY_BT;
  dim3 grid(1, T * (T + 1) / 2);
  ++g_launches; k_gram_syrk<<<grid, SY_THREADS, SYRK_SMEM, s>>>(XT, d.qp, ones, 0, XtX, 0, d.qp, d.q, d.np / SY_BK, 0.0);
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

// factor every G_c (gdim x gdim) in place AND forward-solve: rhs_c <- L_c^-1 rhs_c
void launch_cholesky(const Engine& e, double* rhs, cudaStream_t s) {
  const Dims& d = e.d;
  const int N = d.gdim;
  const int nvalid = d.gmode == 2 ? d.q : d.n;    // rows beyond are identity padding: exact zeros off the diagonal
  const size_t cs = (size_t)N * N;
  const int T = N / PB;
  for (int J = 0; J < T; ++J) {
    if (J > 0) {
      dim3 g2(d.C, T - J);
      ++g_launches; k_chol_update<<<g2, SY_THREADS, SYRK_SMEM, s>>>(e.G, cs, N, e.G, N, nvalid, J * PB / SY_BK, J);
    }
    ++g_launches; k_potf2_128<<<d.C, 256, POTF2_SMEM, s>>>(e.G, cs, N, J, rhs, e.dinv, e.status);
    if (J + 1 < T) {
      dim3 g1(T - J - 1, d.C);
      ++g_launches; k_trsm_128<<<g1, 128, TRSM_SMEM, s>>>(e.G, cs, N, J, e.dinv, rhs);
    }
  }
}

// rhs_c <- L_c^-T (rhs_c + addz_c)  (the forward half already happened inside launch_cholesky)
void launch_chol_solve(const Engine& e, double* rhs, const double* addz, cudaStream_t s) {
  const Dims& d = e.d;
  const int N = d.gdim;
  ++g_launches; k_trsv_bwd128<<<d.C, 256, trsvb_smem(N), s>>>(e.G, (size_t)N * N, N, rhs, e.dinv, addz);
}

int chol_max_dim() { return CHOL_MAX_DIM; }

}  // namespace bnr
