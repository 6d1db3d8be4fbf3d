// Dense FP64 linear algebra of the gamma draw (update_gamma!, src/gibbs.jl:420-438), batched over chains:
//   G_c = X diag(S_c) X' + I          -> k_syrk<0>  : FP64 tensor-core (DMMA m8n8k4) SYRK, cp.async 3-stage pipeline
//   G_c = L_c L_c'                    -> blocked right-looking Cholesky: k_potf2 / k_trsm_panel / k_syrk<1>
//   a4  = L_c^-T L_c^-1 rhs           -> k_trsv_fwd / k_trsv_bwd
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
//   MODE 1:  C_c[i][j] -= sum_k P_c[i][k] P_c[j][k]                  P_c = a 64-column panel of C_c itself
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
__global__ void __launch_bounds__(256, 1)
k_syrk(const double* __restrict__ A, size_t a_chain_stride, int ld, const double* __restrict__ scale,
       size_t scale_stride, double* __restrict__ Cm, size_t c_chain_stride, int np, int nk, int origin) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = blockIdx.y;
  // lower-triangular tile enumeration: t -> (ib >= jb)
  const int t = blockIdx.x;
  int ib = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
  while ((ib + 1) * (ib + 2) / 2 <= t) ++ib;
  while (ib * (ib + 1) / 2 > t) --ib;
  const int jb = t - ib * (ib + 1) / 2;
  const int i0 = origin + ib * SY_BT, j0 = origin + jb * SY_BT;
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

// ------------------------------------------------------------------------------------------------------------
// Cholesky panel kernels (NB = 64)
// ------------------------------------------------------------------------------------------------------------
// factor the 64 x 64 diagonal block kb of every chain in shared memory.  grid = C, block = 256.
__global__ void __launch_bounds__(256) k_potf2(double* __restrict__ G, size_t chain_stride, int np, int kb, int* status) {
  __shared__ double Ls[CHOL_NB][CHOL_NB + 1];
  const int c = blockIdx.x, tid = threadIdx.x;
  double* Gc = G + (size_t)c * chain_stride + (size_t)kb * CHOL_NB * np + (size_t)kb * CHOL_NB;
  for (int id = tid; id < CHOL_NB * CHOL_NB; id += 256) {
    const int r = id & 63, cc = id >> 6;
    Ls[r][cc] = (r >= cc) ? Gc[(size_t)cc * np + r] : 0.0;
  }
  __syncthreads();
  // right-looking elimination with the column scaling deferred to the write-back: one barrier per column.
  // Ls[r][cc] -= a_rj a_ccj / d_jj only touches columns > j, so column j and d_jj are read-only in step j.
  const int r = tid & 63, cg = tid >> 6;
  for (int j = 0; j < CHOL_NB; ++j) {
    __syncthreads();
    const double djj = Ls[j][j];
    if (tid == 0 && !(djj > 0.0)) atomicOr(&status[c], BNR_ST_G_NOTPD_);
    const double lrj = Ls[r][j] / djj;
    for (int cc = j + 1 + cg; cc <= r; cc += 4) Ls[r][cc] -= lrj * Ls[cc][j];
  }
  __syncthreads();
  for (int id = tid; id < CHOL_NB * CHOL_NB; id += 256) {
    const int rr = id & 63, cc = id >> 6;
    if (rr >= cc) {
      const double sd = sqrt(Ls[cc][cc]);
      Gc[(size_t)cc * np + rr] = (rr == cc) ? sd : Ls[rr][cc] / sd;
    }
  }
}

// panel solve below the diagonal block: L21 = A21 L11^-T, one thread per row.  grid = (ceil(rows/128), C), block 128
__global__ void __launch_bounds__(128) k_trsm_panel(double* __restrict__ G, size_t chain_stride, int np, int kb) {
  __shared__ double Ls[CHOL_NB * CHOL_NB];   // L11 row-major: Ls[j*64 + p] = L11[j][p]
  __shared__ double dinv[CHOL_NB];
  const int c = blockIdx.y, tid = threadIdx.x;
  double* Gc = G + (size_t)c * chain_stride;
  const double* D = Gc + (size_t)kb * CHOL_NB * np + (size_t)kb * CHOL_NB;
  for (int id = tid; id < CHOL_NB * CHOL_NB; id += 128) {
    const int r = id & 63, cc = id >> 6;
    Ls[r * CHOL_NB + cc] = D[(size_t)cc * np + r];
  }
  __syncthreads();
  if (tid < CHOL_NB) dinv[tid] = 1.0 / Ls[tid * CHOL_NB + tid];
  __syncthreads();
  const int row = (kb + 1) * CHOL_NB + blockIdx.x * 128 + tid;
  if (row >= np) return;
  double* prow = Gc + (size_t)kb * CHOL_NB * np + row;
  double x[CHOL_NB];
#pragma unroll
  for (int j = 0; j < CHOL_NB; ++j) x[j] = prow[(size_t)j * np];
#pragma unroll
  for (int j = 0; j < CHOL_NB; ++j) {
    double s = x[j];
#pragma unroll
    for (int p = 0; p < j; ++p) s -= x[p] * Ls[j * CHOL_NB + p];
    x[j] = s * dinv[j];
  }
#pragma unroll
  for (int j = 0; j < CHOL_NB; ++j) prow[(size_t)j * np] = x[j];
}

// ------------------------------------------------------------------------------------------------------------
// triangular solves with the factor: rhs <- L^-1 rhs (fwd), rhs <- L^-T rhs (bwd).  grid = C, block = 256.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_trsv_fwd(const double* __restrict__ G, size_t chain_stride, int np,
                                                  double* __restrict__ rhs) {
  extern __shared__ double sm[];
  double* r = sm;                       // [np]
  double* Ls = r + np;                  // [64*65]
  const int c = blockIdx.x, tid = threadIdx.x;
  const double* Gc = G + (size_t)c * chain_stride;
  for (int i = tid; i < np; i += 256) r[i] = rhs[(size_t)c * np + i];
  const int T = np / CHOL_NB;
  for (int kb = 0; kb < T; ++kb) {
    const double* D = Gc + (size_t)kb * CHOL_NB * np + (size_t)kb * CHOL_NB;
    __syncthreads();
    for (int id = tid; id < CHOL_NB * CHOL_NB; id += 256) {
      const int rr = id & 63, cc = id >> 6;
      Ls[rr * 65 + cc] = D[(size_t)cc * np + rr];
    }
    __syncthreads();
    double* x = r + kb * CHOL_NB;
    if (tid < 32) {   // one warp solves the 64 x 64 lower system; lane owns rows lane and lane+32
      double x0 = x[tid], x1 = x[tid + 32];
      for (int j = 0; j < CHOL_NB; ++j) {
        double xj;
        if (j < 32) { xj = __shfl_sync(0xffffffffu, x0, j) / Ls[j * 65 + j]; if (tid == j) x0 = xj; }
        else { xj = __shfl_sync(0xffffffffu, x1, j - 32) / Ls[j * 65 + j]; if (tid == j - 32) x1 = xj; }
        if (tid > j) x0 -= Ls[tid * 65 + j] * xj;
        if (tid + 32 > j) x1 -= Ls[(tid + 32) * 65 + j] * xj;
      }
      x[tid] = x0; x[tid + 32] = x1;
    }
    __syncthreads();
    const double* P = Gc + (size_t)kb * CHOL_NB * np;
    for (int i = (kb + 1) * CHOL_NB + tid; i < np; i += 256) {
      double s = 0.0;
#pragma unroll 8
      for (int j = 0; j < CHOL_NB; ++j) s += P[(size_t)j * np + i] * x[j];
      r[i] -= s;
    }
  }
  __syncthreads();
  for (int i = tid; i < np; i += 256) rhs[(size_t)c * np + i] = r[i];
}

__global__ void __launch_bounds__(256) k_trsv_bwd(const double* __restrict__ G, size_t chain_stride, int np,
                                                  double* __restrict__ rhs) {
  extern __shared__ double sm[];
  double* r = sm;
  double* Ls = r + np;
  const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* Gc = G + (size_t)c * chain_stride;
  for (int i = tid; i < np; i += 256) r[i] = rhs[(size_t)c * np + i];
  const int T = np / CHOL_NB;
  for (int kb = T - 1; kb >= 0; --kb) {
    const double* D = Gc + (size_t)kb * CHOL_NB * np + (size_t)kb * CHOL_NB;
    __syncthreads();
    for (int id = tid; id < CHOL_NB * CHOL_NB; id += 256) {
      const int rr = id & 63, cc = id >> 6;
      Ls[rr * 65 + cc] = D[(size_t)cc * np + rr];
    }
    __syncthreads();
    double* x = r + kb * CHOL_NB;
    if (tid < 32) {   // solve L11' x = b : x_j = (b_j - sum_{i>j} L[i][j] x_i) / L[j][j], j descending
      double x0 = x[tid], x1 = x[tid + 32];
      for (int j = CHOL_NB - 1; j >= 0; --j) {
        double xj;
        if (j < 32) { xj = __shfl_sync(0xffffffffu, x0, j) / Ls[j * 65 + j]; if (tid == j) x0 = xj; }
        else { xj = __shfl_sync(0xffffffffu, x1, j - 32) / Ls[j * 65 + j]; if (tid == j - 32) x1 = xj; }
        if (tid < j) x0 -= Ls[j * 65 + tid] * xj;
        if (tid + 32 < j) x1 -= Ls[j * 65 + tid + 32] * xj;
      }
      x[tid] = x0; x[tid + 32] = x1;
    }
    __syncthreads();
    // r[j] -= sum_{ii<64} L[kb*64+ii][j] x[ii] for every earlier column j: one warp per column, lanes over ii
    const double xa = x[lane], xb = x[lane + 32];
    for (int j = warp; j < kb * CHOL_NB; j += 8) {
      const double* col = Gc + (size_t)j * np + (size_t)kb * CHOL_NB;
      double s = col[lane] * xa + col[lane + 32] * xb;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) r[j] -= s;
    }
  }
  __syncthreads();
  for (int i = tid; i < np; i += 256) rhs[(size_t)c * np + i] = r[i];
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
  cudaFuncSetAttribute(k_syrk<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SYRK_SMEM);
  cudaFuncSetAttribute(k_syrk<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SYRK_SMEM);
  cudaFuncSetAttribute(k_trsv_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  cudaFuncSetAttribute(k_trsv_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
}

void launch_syrk_G(const Engine& e, cudaStream_t s) {
  const Dims& d = e.d;
  const int T = d.np / SY_BT;
  dim3 grid(T * (T + 1) / 2, d.C);
  ++g_launches; k_syrk<0><<<grid, 256, SYRK_SMEM, s>>>(e.X, 0, d.np, e.S, (size_t)d.qp, e.G, (size_t)d.np * d.np, d.np,
                                         d.qp / SY_BK, 0);
  if (e.aux.G_copy) {
    dim3 g2(d.np, d.C);
    ++g_launches; k_copy_sym<<<g2, 256, 0, s>>>(e.G, (size_t)d.np * d.np, d.np, e.aux.G_copy);
  }
}

void launch_cholesky(const Engine& e, cudaStream_t s) {
  const Dims& d = e.d;
  const size_t cs = (size_t)d.np * d.np;
  const int T = d.np / CHOL_NB;
  for (int kb = 0; kb < T; ++kb) {
    ++g_launches; k_potf2<<<d.C, 256, 0, s>>>(e.G, cs, d.np, kb, e.status);
    const int rows = d.np - (kb + 1) * CHOL_NB;
    if (rows <= 0) break;
    dim3 g1((rows + 127) / 128, d.C);
    ++g_launches; k_trsm_panel<<<g1, 128, 0, s>>>(e.G, cs, d.np, kb);
    const int Tt = (rows + SY_BT - 1) / SY_BT;
    dim3 g2(Tt * (Tt + 1) / 2, d.C);
    ++g_launches; k_syrk<1><<<g2, 256, SYRK_SMEM, s>>>(e.G + (size_t)kb * CHOL_NB * d.np, cs, d.np, nullptr, 0, e.G, cs, d.np,
                                         CHOL_NB / SY_BK, (kb + 1) * CHOL_NB);
  }
}

void launch_chol_solve(const Engine& e, cudaStream_t s) {
  const Dims& d = e.d;
  const size_t sm = sizeof(double) * ((size_t)d.np + CHOL_NB * 65);
  ++g_launches; k_trsv_fwd<<<d.C, 256, sm, s>>>(e.G, (size_t)d.np * d.np, d.np, e.rhs);
  ++g_launches; k_trsv_bwd<<<d.C, 256, sm, s>>>(e.G, (size_t)d.np * d.np, d.np, e.rhs);
}

}  // namespace bnr
