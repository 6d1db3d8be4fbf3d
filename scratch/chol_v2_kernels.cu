// ------------------------------------------------------------------------------------------------------------
// Blocked left-looking Cholesky, block size 128, with the forward solve L w = rhs folded in.
//   for J = 0 .. np/128-1:
//     k_chol_update   (DMMA, above):  G[J.., J] -= L[J.., 0:J] L[J, 0:J]'
//     k_potf2_128     one CTA per chain: rhs_J -= L[J, 0:J] w[0:J]; factor the 128 x 128 diagonal block in shared
//                     memory (4 sub-blocks of 32: a warp factors 32 x 32 in registers with shuffles, threads solve the
//                     rows below, everybody updates the trailing part); rhs_J rides along as row 128, so w_J = L_JJ^-1 rhs_J
//                     comes out of the same elimination.
//     k_trsm_128      rows below the diagonal block: L[i, J] = G[i, J] L_JJ^-T, one thread per row, two 64-column halves.
// ------------------------------------------------------------------------------------------------------------
constexpr int PB = 128;            // panel / diagonal block size
constexpr int PB_LD = PB + 1;      // shared-memory row stride of the (PB+1) x PB working block
constexpr size_t POTF2_SMEM = sizeof(double) * ((size_t)(PB + 1) * PB_LD + 1024);

__global__ void __launch_bounds__(256) k_potf2_128(double* __restrict__ G, size_t chain_stride, int np, int J,
                                                   double* __restrict__ rhs, int* status) {
  extern __shared__ double sm[];
  double* A = sm;                          // A[r][c] at A[r * PB_LD + c], rows 0..128 (row 128 = rhs), cols 0..127
  double* wprev = sm + (PB + 1) * PB_LD;   // [<= 1024] previously solved w (J*128 entries used)
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
    double acc = 0.0;
    const int k0 = half * (kprev / 2), k1 = half ? kprev : kprev / 2;
#pragma unroll 8
    for (int k = k0; k < k1; ++k) acc += Lrow[(size_t)k * np] * wprev[k];
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
        const double d = sqrt(djj);
        const double lj = (lane == j) ? d : a[j] / d;
        a[j] = lj;
#pragma unroll
        for (int cc = j + 1; cc < 32; ++cc) {
          const double lcj = __shfl_sync(0xffffffffu, lj, cc);
          if (lane >= cc) a[cc] -= lj * lcj;
        }
      }
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (lane >= j) A[(o + lane) * PB_LD + o + j] = a[j];
    }
    __syncthreads();
    // rows below (incl. the rhs row 128): x L_ss' = a, thread per row
    const int nbelow = PB + 1 - (o + 32);
    if (tid < nbelow) {
      double* row = A + (o + 32 + tid) * PB_LD + o;
      double x[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = row[j];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        double sacc = x[j];
        const double* Lj = A + (o + j) * PB_LD + o;
#pragma unroll
        for (int p = 0; p < j; ++p) sacc -= x[p] * Lj[p];
        x[j] = sacc / Lj[j];
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) row[j] = x[j];
    }
    __syncthreads();
    // trailing update inside the block: A[r][cc] -= sum_p L[r][o+p] L[cc][o+p], o+32 <= cc <= r <= 128 (cc < 128)
    const int base = o + 32;
    const int nr = PB + 1 - base;          // rows base .. 128
    const int ncol = PB - base;            // cols base .. 127
    for (int id = tid; id < nr * ncol; id += 256) {
      const int r = base + id / ncol, cc = base + id % ncol;
      if (cc > r) continue;
      const double* Lr = A + r * PB_LD + o;
      const double* Lc = A + cc * PB_LD + o;
      double sacc = 0.0;
#pragma unroll
      for (int p = 0; p < 32; ++p) sacc += Lr[p] * Lc[p];
      A[r * PB_LD + cc] -= sacc;
    }
    __syncthreads();
  }
  if (bad) atomicOr(&status[c], BNR_ST_G_NOTPD_);
  for (int id = tid; id < PB * PB; id += 256) {
    const int r = id & (PB - 1), cc = id >> 7;
    if (r >= cc) D[(size_t)cc * np + r] = A[r * PB_LD + cc];
  }
  if (tid < PB) rc[kprev + tid] = A[PB * PB_LD + tid];
}

// rows below the diagonal block.  grid = (rows_below / 128, C), block = 128, dynamic smem = 3 * 64*64 doubles
constexpr size_t TRSM_SMEM = sizeof(double) * 3 * 64 * 64;

__global__ void __launch_bounds__(128) k_trsm_128(double* __restrict__ G, size_t chain_stride, int np, int J) {
  extern __shared__ double sm[];
  double* L11 = sm;                 // [j][p] row-major 64 x 64
  double* L21 = sm + 64 * 64;
  double* L22 = sm + 2 * 64 * 64;
  const int c = blockIdx.y, tid = threadIdx.x;
  double* Gc = G + (size_t)c * chain_stride;
  const double* D = Gc + (size_t)J * PB * np + (size_t)J * PB;
  for (int id = tid; id < 64 * 64; id += 128) {
    const int r = id & 63, cc = id >> 6;
    L11[r * 64 + cc] = D[(size_t)cc * np + r];
    L21[r * 64 + cc] = D[(size_t)cc * np + 64 + r];
    L22[r * 64 + cc] = D[(size_t)(64 + cc) * np + 64 + r];
  }
  __syncthreads();
  const int row = (J + 1) * PB + blockIdx.x * 128 + tid;
  if (row >= np) return;
  double* prow = Gc + (size_t)J * PB * np + row;
  double x[64];
#pragma unroll
  for (int j = 0; j < 64; ++j) x[j] = prow[(size_t)j * np];
#pragma unroll
  for (int j = 0; j < 64; ++j) {
    double s = x[j];
#pragma unroll
    for (int p = 0; p < j; ++p) s -= x[p] * L11[j * 64 + p];
    x[j] = s / L11[j * 64 + j];
  }
#pragma unroll
  for (int j = 0; j < 64; ++j) prow[(size_t)j * np] = x[j];
  // second half: a2_j = A[row][64 + j] - sum_p x1[p] L21[j][p]
  double* prow2 = prow + (size_t)64 * np;
#pragma unroll 4
  for (int j = 0; j < 64; ++j) {
    double s = prow2[(size_t)j * np];
#pragma unroll
    for (int p = 0; p < 64; ++p) s -= x[p] * L21[j * 64 + p];
    prow2[(size_t)j * np] = s;
  }
#pragma unroll
  for (int j = 0; j < 64; ++j) x[j] = prow2[(size_t)j * np];
#pragma unroll
  for (int j = 0; j < 64; ++j) {
    double s = x[j];
#pragma unroll
    for (int p = 0; p < j; ++p) s -= x[p] * L22[j * 64 + p];
    x[j] = s / L22[j * 64 + j];
  }
#pragma unroll
  for (int j = 0; j < 64; ++j) prow2[(size_t)j * np] = x[j];
}

// ------------------------------------------------------------------------------------------------------------
// backward solve  L' x = w  (w in rhs, overwritten by x), left-looking over 128-blocks from the bottom.
// grid = C, block = 256 (8 warps): warp per column for the matvec with the rows below (coalesced, 4 columns in
// flight per warp), then a 128 x 128 transposed triangular solve in shared memory.
// ------------------------------------------------------------------------------------------------------------
constexpr size_t TRSVB_SMEM = sizeof(double) * ((size_t)PB * PB_LD + 1024 + PB);

__global__ void __launch_bounds__(256) k_trsv_bwd128(const double* __restrict__ G, size_t chain_stride, int np,
                                                     double* __restrict__ rhs) {
  extern __shared__ double sm[];
  double* Ls = sm;                     // [r][c] lower block
  double* x = sm + PB * PB_LD;         // [np] solution so far (entries >= (J+1)*128 valid)
  double* b = x + 1024;                // [128] current right-hand side
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
    // b_j = w_j - sum_{i >= r0} L[i][J*128 + j] x_i
    for (int jj = warp * 16; jj < warp * 16 + 16; jj += 4) {
      double acc[4] = {0.0, 0.0, 0.0, 0.0};
      const double* col = Gc + (size_t)(J * PB + jj) * np + r0;
      for (int i = lane; i < nrow; i += 32) {
        const double xi = x[r0 + i];
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] += col[(size_t)u * np + i] * xi;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
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
      for (int j = PB - 1; j >= 0; --j) {
        const int u = j >> 5, src = j & 31;
        double xj = 0.0;
#pragma unroll
        for (int uu = 0; uu < 4; ++uu) if (uu == u) xj = __shfl_sync(0xffffffffu, v[uu], src);
        xj /= Ls[j * PB_LD + j];
#pragma unroll
        for (int uu = 0; uu < 4; ++uu) {
          const int idx = lane + 32 * uu;
          if (idx == j) v[uu] = xj;
          else if (idx < j) v[uu] -= Ls[j * PB_LD + idx] * xj;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { x[J * PB + lane + 32 * u] = v[u]; rc[J * PB + lane + 32 * u] = v[u]; }
    }
    __syncthreads();
  }
}

