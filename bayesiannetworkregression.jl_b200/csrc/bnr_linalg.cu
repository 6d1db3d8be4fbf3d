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
// SYRK on FP64 tensor cores.
//   MODE 0:  C_c[i][j] = sum_k A[i][k] s_c[k] A[j][k] + (i==j)      A = X (shared by all chains), K = qp
//   MODE 1:  C_c[i][j] -= sum_{k < nk*16} L_c[i][k] L_c[j][k]        left-looking Cholesky update of block column J
// The operand is "k-major": element (row, k) at A[row + ld*k] (rows contiguous) - exactly how X (column-major
// n x q) and a column panel of the column-major G are stored, so tiles are staged with 16-byte cp.async and no
// transposition.  CTA tile 128 x 128, k-step 16, 8 warps (2 along j x 4 along i), warp tile 64(j) x 32(i):
// the MMA "M" dimension runs along j (columns of C) and "N" along i (rows of C), so every accumulator pair is two
// consecutive rows of one column of the column-major C -> 16-byte stores.
// grid = (lower-triangular tiles, C); dynamic smem = SYRK_SMEM.
// ------------------------------------------------------------------------------------------------------------
constexpr int SY_BT = 128;        // tile edge
constexpr int SY_BK = 16;         // k-step
constexpr int SY_LDS = 132;       // smem row stride in doubles (== 4 mod 16 -> conflict-free fragment loads)
constexpr int SY_STAGES = 3;
constexpr int SY_STAGE_DBL = 2 * SY_BK * SY_LDS + SY_BK;   // two operand tiles + 16 scales
constexpr size_t SYRK_SMEM = (size_t)SY_STAGES * SY_STAGE_DBL * sizeof(double);

template <int MODE>
__device__ __forceinline__ void
syrk_body(const double* __restrict__ A, size_t a_chain_stride, int ld, const double* __restrict__ scale,
       size_t scale_stride, double* __restrict__ Cm, size_t c_chain_stride, int np, int nk, int origin) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = blockIdx.y;
  int ib, jb;
  if (MODE == 0) {
    // lower-triangular tile enumeration: t -> (ib >= jb)
    const int t = blockIdx.x;
    ib = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
    while ((ib + 1) * (ib + 2) / 2 <= t) ++ib;
    while (ib * (ib + 1) / 2 > t) --ib;
    jb = t - ib * (ib + 1) / 2;
  } else {
    // left-looking update of block column J = origin: tiles (J + blockIdx.x, J)
    jb = origin;
    ib = origin + blockIdx.x;
  }
  const int i0 = ib * SY_BT, j0 = jb * SY_BT;
  const double* Ac = A + (size_t)c * a_chain_stride;
  const double* sc = (MODE == 0) ? scale + (size_t)c * scale_stride : nullptr;

  auto load_stage = [&](int stage, int kt) {
    double* sj = smem + (size_t)stage * SY_STAGE_DBL;
    double* si = sj + SY_BK * SY_LDS;
    const int kbase = kt * SY_BK;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int id = tid + 256 * r;
      const int krow = id >> 6, c16 = id & 63;
      const double* gsrc = Ac + (size_t)(kbase + krow) * ld;
      cp_async16(sj + krow * SY_LDS + 2 * c16, gsrc + j0 + 2 * c16);
      cp_async16(si + krow * SY_LDS + 2 * c16, gsrc + i0 + 2 * c16);
    }
    if (MODE == 0 && tid < 8) cp_async16(si + SY_BK * SY_LDS + 2 * tid, sc + kbase + 2 * tid);
  };

  double acc[8][4][2];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

  const int wj = warp >> 2, wi = warp & 3;
  const int lk = lane & 3, lr = lane >> 2;

#pragma unroll
  for (int s = 0; s < SY_STAGES - 1; ++s) {
    if (s < nk) load_stage(s, s);
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<SY_STAGES - 2>();
    __syncthreads();
    const int pre = kt + SY_STAGES - 1;
    if (pre < nk) load_stage(pre % SY_STAGES, pre);
    cp_async_commit();
    const double* sj = smem + (size_t)(kt % SY_STAGES) * SY_STAGE_DBL;
    const double* si = sj + SY_BK * SY_LDS;
    const double* ss = si + SY_BK * SY_LDS;
#pragma unroll
    for (int k4 = 0; k4 < SY_BK / 4; ++k4) {
      const int kr = k4 * 4 + lk;
      double af[8], bf[4];
#pragma unroll
      for (int mf = 0; mf < 8; ++mf) af[mf] = sj[kr * SY_LDS + wj * 64 + mf * 8 + lr];
      const double sk = (MODE == 0) ? ss[kr] : 1.0;
#pragma unroll
      for (int nf = 0; nf < 4; ++nf) bf[nf] = si[kr * SY_LDS + wi * 32 + nf * 8 + lr] * sk;
#pragma unroll
      for (int mf = 0; mf < 8; ++mf)
#pragma unroll
        for (int nf = 0; nf < 4; ++nf) dmma884(acc[mf][nf][0], acc[mf][nf][1], af[mf], bf[nf]);
    }
  }
  cp_async_wait<0>();

  double* Cc = Cm + (size_t)c * c_chain_stride;
#pragma unroll
  for (int mf = 0; mf < 8; ++mf) {
    const int j = j0 + wj * 64 + mf * 8 + lr;
#pragma unroll
    for (int nf = 0; nf < 4; ++nf) {
      const int i = i0 + wi * 32 + nf * 8 + 2 * lk;
      if (i < np && j < np) {
        double2* p = reinterpret_cast<double2*>(Cc + (size_t)j * np + i);
        double2 v;
        if (MODE == 0) {
          v.x = acc[mf][nf][0] + (i == j ? 1.0 : 0.0);
          v.y = acc[mf][nf][1] + (i + 1 == j ? 1.0 : 0.0);
        } else {
          v = *p;
          v.x -= acc[mf][nf][0];
          v.y -= acc[mf][nf][1];
        }
        *p = v;
      }
    }
  }
}

__global__ void __launch_bounds__(256, 1)
k_gram_syrk(const double* __restrict__ X, int ld, const double* __restrict__ scale, size_t scale_stride,
            double* __restrict__ G, size_t g_chain_stride, int np, int nk) {
  syrk_body<0>(X, 0, ld, scale, scale_stride, G, g_chain_stride, np, nk, 0);
}

__global__ void __launch_bounds__(256, 1)
k_chol_update(const double* __restrict__ P, size_t chain_stride, int ld, double* __restrict__ G, int np, int nk,
              int origin) {
  syrk_body<1>(P, chain_stride, ld, nullptr, 0, G, chain_stride, np, nk, origin);
}

// ------------------------------------------------------------------------------------------------------------
// Blocked left-looking Cholesky, block size 128, with the forward solve L w = rhs folded in.
//   for J = 0 .. np/128-1:
//     k_chol_update   (DMMA, above):  G[J.., J] -= L[J.., 0:J] L[J, 0:J]'
//     k_potf2_128     one CTA per chain: rhs_J -= L[J, 0:J] w[0:J]; factor the 128 x 128 diagonal block in shared
//                     memory (4 sub-blocks of 32: a warp factors 32 x 32 in registers with shuffles, threads eliminate
//                     the rows below, everybody updates the trailing part in 2 x 2 register tiles); rhs_J rides along as
//                     row 128, so w_J = L_JJ^-1 rhs_J comes out of the same elimination.  1/L_jj goes to `dinv`.
//     k_trsm_128      rows below the diagonal block: L[i, J] = G[i, J] L_JJ^-T, one thread per row, two 64-column
//                     halves, right-looking inside the thread (independent FMAs, no divisions).
// All inner loops are arranged so that consecutive FP64 FMAs are independent: these kernels are latency-bound.
// ------------------------------------------------------------------------------------------------------------
constexpr int PB = 128;            // panel / diagonal block size
constexpr int PB_LD = PB + 1;      // shared-memory row stride of the (PB+1) x PB working block
constexpr size_t POTF2_SMEM = sizeof(double) * ((size_t)(PB + 1) * PB_LD + 1024 + 32);

__global__ void __launch_bounds__(256) k_potf2_128(double* __restrict__ G, size_t chain_stride, int np, int J,
                                                   double* __restrict__ rhs, double* __restrict__ dinv_out,
                                                   int* status) {
  extern __shared__ double sm[];
  double* A = sm;                          // A[r][c] at A[r * PB_LD + c], rows 0..128 (row 128 = rhs), cols 0..127
  double* wprev = sm + (PB + 1) * PB_LD;   // [<= 1024] previously solved w (J*128 entries used)
  double* dinv = wprev + 1024;             // [32] reciprocal diagonal of the current 32-block
  const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* Gc = G + (size_t)c * chain_stride;
  double* D = Gc + (size_t)J * PB * np + (size_t)J * PB;
  double* rc = rhs + (size_t)c * np;
  const int kprev = J * PB;
  for (int id = tid; id < PB * PB; id += 256) {
    const int r = id & (PB - 1), cc = id >> 7;
    A[r * PB_LD + cc] = (r >= cc) ? D[(size_t)cc * np + r] : 0.0;
  }
  for (int k = tid; k < kprev; k += 256) wprev[k] = rc[k];
  __syncthreads();
  // rhs_J -= L[J-block rows, 0:kprev] w[0:kprev]  (thread (jj, half) walks half of the columns; coalesced in jj)
  {
    const int jj = tid & (PB - 1), half = tid >> 7;
    const double* Lrow = Gc + (size_t)J * PB + jj;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    const int k0 = half * (kprev / 2), k1 = half ? kprev : kprev / 2;
    for (int k = k0; k < k1; k += 4) {      // kprev is a multiple of 128
      a0 += Lrow[(size_t)k * np] * wprev[k];
      a1 += Lrow[(size_t)(k + 1) * np] * wprev[k + 1];
      a2 += Lrow[(size_t)(k + 2) * np] * wprev[k + 2];
      a3 += Lrow[(size_t)(k + 3) * np] * wprev[k + 3];
    }
    const double acc = (a0 + a1) + (a2 + a3);
    double* part = A + PB * PB_LD;          // row 128 of the working block
    if (half == 0) part[jj] = rc[kprev + jj] - acc;
    __syncthreads();
    if (half == 1) part[jj] -= acc;
  }
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
// shared: Lt11, Lt21, Lt22 as [p][j] (column p of the 64 x 64 sub-block contiguous in j) + 128 reciprocal diagonals
constexpr size_t TRSM_SMEM = sizeof(double) * (3 * 64 * 64 + 128);

__global__ void __launch_bounds__(128) k_trsm_128(double* __restrict__ G, size_t chain_stride, int np, int J,
                                                  const double* __restrict__ dinv_g) {
  extern __shared__ double sm[];
  double* L11 = sm;                 // [p][j] = L11[j][p]
  double* L21 = sm + 64 * 64;       // [p][j] = L21[j][p]   (j: second-half column, p: first-half column)
  double* L22 = sm + 2 * 64 * 64;
  double* dinv = sm + 3 * 64 * 64;
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
  __syncthreads();
  const int row = (J + 1) * PB + blockIdx.x * 128 + tid;
  if (row >= np) return;
  double* prow = Gc + (size_t)J * PB * np + row;
  double* prow2 = prow + (size_t)64 * np;
  double x[64];
#pragma unroll
  for (int j = 0; j < 64; ++j) x[j] = prow[(size_t)j * np];
  // first half: x_j = a_j / L11[j][j]; a_k -= x_j L11[k][j] (k > j)
#pragma unroll
  for (int j = 0; j < 64; ++j) {
    const double xj = x[j] * dinv[j];
    x[j] = xj;
    const double* l1 = L11 + j * 64;
#pragma unroll
    for (int k = j + 1; k < 64; ++k) x[k] -= xj * l1[k];
  }
#pragma unroll
  for (int j = 0; j < 64; ++j) prow[(size_t)j * np] = x[j];
  // second half, 16 columns at a time: y_k = A[row][64 + k] - sum_p x_p L21[k][p]
#pragma unroll 1
  for (int k0 = 0; k0 < 64; k0 += 16) {
    double y[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) y[k] = prow2[(size_t)(k0 + k) * np];
#pragma unroll
    for (int p = 0; p < 64; ++p) {
      const double* l2 = L21 + p * 64 + k0;
#pragma unroll
      for (int k = 0; k < 16; ++k) y[k] -= x[p] * l2[k];
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) prow2[(size_t)(k0 + k) * np] = y[k];
  }
#pragma unroll
  for (int j = 0; j < 64; ++j) x[j] = prow2[(size_t)j * np];
#pragma unroll
  for (int j = 0; j < 64; ++j) {
    const double yj = x[j] * dinv[64 + j];
    x[j] = yj;
    const double* l2 = L22 + j * 64;
#pragma unroll
    for (int k = j + 1; k < 64; ++k) x[k] -= yj * l2[k];
  }
#pragma unroll
  for (int j = 0; j < 64; ++j) prow2[(size_t)j * np] = x[j];
}

// ------------------------------------------------------------------------------------------------------------
// backward solve  L' x = w  (w in rhs, overwritten by x), left-looking over 128-blocks from the bottom.
// grid = C, block = 256 (8 warps): warp per column group for the matvec with the rows below (coalesced, 8 columns in
// flight per warp), then a 128 x 128 transposed triangular solve by one warp (shuffles, reciprocal diagonals).
// ------------------------------------------------------------------------------------------------------------
constexpr size_t TRSVB_SMEM = sizeof(double) * ((size_t)PB * PB_LD + 1024 + 2 * PB);

__global__ void __launch_bounds__(256) k_trsv_bwd128(const double* __restrict__ G, size_t chain_stride, int np,
                                                     double* __restrict__ rhs, const double* __restrict__ dinv_g) {
  extern __shared__ double sm[];
  double* Ls = sm;                     // [r][c] lower block
  double* x = sm + PB * PB_LD;         // [np] solution so far (entries >= (J+1)*128 valid)
  double* b = x + 1024;                // [128] current right-hand side
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
    // b_j = w_j - sum_{i >= r0} L[i][J*128 + j] x_i : warp w owns columns 16w .. 16w+15, 8 at a time
    for (int jj = warp * 16; jj < warp * 16 + 16; jj += 8) {
      double acc[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] = 0.0;
      const double* col = Gc + (size_t)(J * PB + jj) * np + r0;
      for (int i = lane; i < nrow; i += 32) {
        const double xi = x[r0 + i];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] += col[(size_t)u * np + i] * xi;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        double v = acc[u];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) b[jj + u] = rc[J * PB + jj + u] - v;
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
  cudaFuncSetAttribute(k_trsv_bwd128, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRSVB_SMEM);
}

void launch_syrk_G(const Engine& e, cudaStream_t s) {
  const Dims& d = e.d;
  const int T = d.np / SY_BT;
  dim3 grid(T * (T + 1) / 2, d.C);
  ++g_launches; k_gram_syrk<<<grid, 256, SYRK_SMEM, s>>>(e.X, d.np, e.S, (size_t)d.qp, e.G, (size_t)d.np * d.np, d.np, d.qp / SY_BK);
  if (e.aux.G_copy) {
    dim3 g2(d.np, d.C);
    ++g_launches; k_copy_sym<<<g2, 256, 0, s>>>(e.G, (size_t)d.np * d.np, d.np, e.aux.G_copy);
  }
}

// factor every G_c in place AND forward-solve: rhs_c <- L_c^-1 rhs_c
void launch_cholesky(const Engine& e, cudaStream_t s) {
  const Dims& d = e.d;
  const size_t cs = (size_t)d.np * d.np;
  const int T = d.np / PB;
  for (int J = 0; J < T; ++J) {
    if (J > 0) {
      dim3 g2(T - J, d.C);
      ++g_launches; k_chol_update<<<g2, 256, SYRK_SMEM, s>>>(e.G, cs, d.np, e.G, d.np, J * PB / SY_BK, J);
    }
    ++g_launches; k_potf2_128<<<d.C, 256, POTF2_SMEM, s>>>(e.G, cs, d.np, J, e.rhs, e.dinv, e.status);
    if (J + 1 < T) {
      dim3 g1(T - J - 1, d.C);
      ++g_launches; k_trsm_128<<<g1, 128, TRSM_SMEM, s>>>(e.G, cs, d.np, J, e.dinv);
    }
  }
}

// rhs_c <- L_c^-T rhs_c  (the forward half already happened inside launch_cholesky)
void launch_chol_solve(const Engine& e, cudaStream_t s) {
  const Dims& d = e.d;
  ++g_launches; k_trsv_bwd128<<<d.C, 256, TRSVB_SMEM, s>>>(e.G, (size_t)d.np * d.np, d.np, e.rhs, e.dinv);
}

}  // namespace bnr
