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
  double x[64], y[64];
#pragma unroll
  for (int j = 0; j < 64; ++j) { x[j] = prow[(size_t)j * np]; y[j] = prow2[(size_t)j * np]; }
  // first half: x_j = a_j / L11[j][j]; a_k -= x_j L11[k][j] (k > j); y_k -= x_j L21[k][j] (all k)
#pragma unroll
  for (int j = 0; j < 64; ++j) {
    const double xj = x[j] * dinv[j];
    x[j] = xj;
    const double* l1 = L11 + j * 64;
    const double* l2 = L21 + j * 64;
#pragma unroll
    for (int k = j + 1; k < 64; ++k) x[k] -= xj * l1[k];
#pragma unroll
    for (int k = 0; k < 64; ++k) y[k] -= xj * l2[k];
  }
#pragma unroll
  for (int j = 0; j < 64; ++j) prow[(size_t)j * np] = x[j];
#pragma unroll
  for (int j = 0; j < 64; ++j) {
    const double yj = y[j] * dinv[64 + j];
    y[j] = yj;
    const double* l2 = L22 + j * 64;
#pragma unroll
    for (int k = j + 1; k < 64; ++k) y[k] -= yj * l2[k];
  }
#pragma unroll
  for (int j = 0; j < 64; ++j) prow2[(size_t)j * np] = y[j];
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

