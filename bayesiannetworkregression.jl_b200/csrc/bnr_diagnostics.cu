// Posterior diagnostics computed on the device from the recorded traces (iteration-major rows):
//   * Summary (src/gibbs.jl:1214-1250): per-edge posterior mean and the two order statistics sort(gamma)[lw], [hi]
//     that bound the credible interval, per-node mean xi            -> k_summary_select (radix select, no sort)
//   * gamma / xi effective sample size (BASELINE metric "gamma ESS/sec"; not in the reference): multi-chain Geyer
//     initial-monotone-sequence estimator on direct-lag autocovariances -> k_chain_mean, k_acov_sum, k_ess_finish
// Everything is deterministic (fixed reduction orders, integer histograms).
#include "bnr_engine.cuh"
#include "bnr_kernels.h"

namespace bnr {

// ------------------------------------------------------------------------------------------------------------
// Summary: exact k-th order statistics by 8-bit MSB radix select on the order-preserving 64-bit key of a double.
// One block handles SUM_EPB adjacent parameters (64 contiguous bytes per trace row); thread t reads parameter
// t % SUM_EPB of rows t / SUM_EPB, t / SUM_EPB + 32, ...  Two ranks (lower / upper bound) are selected together.
// grid = ceil(nelem / SUM_EPB), block = 256.
// ------------------------------------------------------------------------------------------------------------
constexpr int SUM_EPB = 8;

__device__ __forceinline__ unsigned long long order_key(double x) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_to_double(unsigned long long k) {
  const unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

__global__ void __launch_bounds__(256) k_summary_select(const double* __restrict__ rows, size_t rowlen, int off,
                                                        int nelem, long long first, long long count,
                                                        long long rank_lo, long long rank_hi,
                                                        double* __restrict__ mean_out, double* __restrict__ lo_out,
                                                        double* __restrict__ hi_out) {
  __shared__ int hist[2][SUM_EPB][256];
  __shared__ unsigned long long prefix[2][SUM_EPB];
  __shared__ long long remain[2][SUM_EPB];
  __shared__ double psum[32][SUM_EPB];
  const int tid = threadIdx.x, e = tid % SUM_EPB, lane = tid / SUM_EPB;   // 32 row lanes
  const int el = blockIdx.x * SUM_EPB + e;
  const bool live = el < nelem;
  const double* base = rows + (size_t)first * rowlen + off + (live ? el : 0);
  if (tid < 2 * SUM_EPB) {
    prefix[tid / SUM_EPB][tid % SUM_EPB] = 0ull;
    remain[tid / SUM_EPB][tid % SUM_EPB] = (tid / SUM_EPB == 0) ? rank_lo : rank_hi;   // 1-based ranks
  }
  double s = 0.0;
  for (int pass = 0; pass < 8; ++pass) {
    const int shift = 56 - 8 * pass;
    for (int i = tid; i < 2 * SUM_EPB * 256; i += 256) (&hist[0][0][0])[i] = 0;
    __syncthreads();
    const unsigned long long p0 = prefix[0][e], p1 = prefix[1][e];
    if (live) {
      for (long long r = lane; r < count; r += 32) {
        const double x = base[(size_t)r * rowlen];
        if (pass == 0) s += x;
        const unsigned long long k = order_key(x);
        const unsigned long long hi = (pass == 0) ? 0ull : (k >> (shift + 8));
        const int dg = (int)((k >> shift) & 0xffull);
        if (hi == p0) atomicAdd(&hist[0][e][dg], 1);
        if (hi == p1) atomicAdd(&hist[1][e][dg], 1);
      }
    }
    __syncthreads();
    if (tid < 2 * SUM_EPB) {
      const int t = tid / SUM_EPB, ee = tid % SUM_EPB;
      long long rem = remain[t][ee];
      int b = 0;
      for (; b < 255; ++b) {
        const int cnt = hist[t][ee][b];
        if (rem <= cnt) break;
        rem -= cnt;
      }
      remain[t][ee] = rem;
      prefix[t][ee] = (prefix[t][ee] << 8) | (unsigned long long)b;
    }
    __syncthreads();
  }
  psum[lane][e] = s;
  __syncthreads();
  if (tid < SUM_EPB && blockIdx.x * SUM_EPB + tid < nelem) {
    double tot = 0.0;
    for (int l = 0; l < 32; ++l) tot += psum[l][tid];
    const int o = blockIdx.x * SUM_EPB + tid;
    mean_out[o] = tot / (double)count;
    if (lo_out) lo_out[o] = key_to_double(prefix[0][tid]);
    if (hi_out) hi_out[o] = key_to_double(prefix[1][tid]);
  }
}

void launch_summary_select(const double* rows, size_t rowlen, int off, int nelem, long long first, long long count,
                           long long rank_lo, long long rank_hi, double* mean_out, double* lo_out, double* hi_out,
                           cudaStream_t s) {
  ++g_launches;
  k_summary_select<<<(nelem + SUM_EPB - 1) / SUM_EPB, 256, 0, s>>>(rows, rowlen, off, nelem, first, count, rank_lo,
                                                                    rank_hi, mean_out, lo_out, hi_out);
}

// ------------------------------------------------------------------------------------------------------------
// ESS.  tr: [C][trace_rows][P] rows of (xi, gamma); window = rows [first, first + N).
//   k_chain_mean : cmean[c][p]                       grid = (ceil(P/128), C), block = 128
//   k_acov_sum   : acov[l][p] = sum_c 1/N sum_t (x_t - m_c)(x_{t+l} - m_c), l = 0 .. L (biased, per-chain centred)
//                  grid = (ceil(P/32), ceil((L+1)/64)), block = 256 = 32 parameters x 8 lag groups of 8 lags;
//                  row chunks of 128 are staged in shared memory, each thread slides an 8-lag register window.
//   k_ess_finish : Geyer initial monotone sequence over the pooled autocorrelations -> ess[p], thread per parameter
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_chain_mean(const double* __restrict__ tr, long long trace_rows, int P,
                                                    long long first, long long N, double* __restrict__ cmean) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
  if (p >= P) return;
  const double* b = tr + ((size_t)c * trace_rows + first) * P + p;
  double s = 0.0;
  for (long long r = 0; r < N; ++r) s += b[(size_t)r * P];
  cmean[(size_t)c * P + p] = s / (double)N;
}

constexpr int AC_TCH = 128;    // rows per staged chunk
constexpr int AC_LB = 64;      // lags per block
constexpr size_t ACOV_SMEM = sizeof(double) * ((size_t)AC_TCH * 32 + (size_t)(AC_TCH + AC_LB) * 32);

__global__ void __launch_bounds__(256) k_acov_sum(const double* __restrict__ tr, long long trace_rows, int P, int C,
                                                  long long first, long long N, int L,
                                                  const double* __restrict__ cmean, double* __restrict__ acov) {
  extern __shared__ double sm[];
  double* A = sm;                       // [AC_TCH][32]       x_t - m
  double* B = sm + AC_TCH * 32;         // [AC_TCH + 64][32]  x_{t + l0 + .} - m
  const int tid = threadIdx.x, pl = tid & 31, g = tid >> 5;
  const int p0 = blockIdx.x * 32, l0 = blockIdx.y * AC_LB;
  const int p = p0 + pl;
  double acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.0;
  for (int c = 0; c < C; ++c) {
    const double* base = tr + ((size_t)c * trace_rows + first) * P;
    const double m = (p < P) ? cmean[(size_t)c * P + p] : 0.0;
    // every thread stages column pl of the tiles: its own parameter, so m is the right centre
    for (long long t0 = 0; t0 < N; t0 += AC_TCH) {
      __syncthreads();
      for (int r = g; r < AC_TCH; r += 8) {
        const long long t = t0 + r;
        A[r * 32 + pl] = (p < P && t < N) ? base[(size_t)t * P + p] - m : 0.0;
      }
      for (int r = g; r < AC_TCH + AC_LB; r += 8) {
        const long long t = t0 + l0 + r;
        B[r * 32 + pl] = (p < P && t < N) ? base[(size_t)t * P + p] - m : 0.0;
      }
      __syncthreads();
      // lags l0 + 8g + k, k = 0..7: acc[k] += A[t] * B[t + 8g + k]
      double w[8];
#pragma unroll
      for (int k = 0; k < 7; ++k) w[k + 1] = B[(8 * g + k) * 32 + pl];
#pragma unroll 8
      for (int t = 0; t < AC_TCH; ++t) {
#pragma unroll
        for (int k = 0; k < 7; ++k) w[k] = w[k + 1];
        w[7] = B[(t + 8 * g + 7) * 32 + pl];
        const double a = A[t * 32 + pl];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += a * w[k];
      }
    }
  }
  if (p < P) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int l = l0 + 8 * g + k;
      if (l <= L) acov[(size_t)l * P + p] = acc[k] / (double)N;
    }
  }
}

// acov_parts: [nparts][L+1][P] (one part per rank), cmeans: [chains][P].  ess[p]; lag_used[p] = lags consumed
// (L + 1 when the Geyer sequence had not terminated inside the lag budget: the estimate is then an upper bound).
__global__ void k_ess_finish(const double* __restrict__ acov_parts, int nparts, const double* __restrict__ cmeans,
                             int chains, int P, long long N, int L, double* __restrict__ ess,
                             double* __restrict__ lag_used) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const size_t part = (size_t)(L + 1) * P;
  auto mean_acov = [&](int l) {
    double s = 0.0;
    for (int r = 0; r < nparts; ++r) s += acov_parts[r * part + (size_t)l * P + p];
    return s / (double)chains;
  };
  const double n = (double)N;
  const double W = mean_acov(0) * n / (n - 1.0);
  double B_over_n = 0.0;
  if (chains > 1) {
    double mm = 0.0;
    for (int c = 0; c < chains; ++c) mm += cmeans[(size_t)c * P + p];
    mm /= chains;
    for (int c = 0; c < chains; ++c) { const double dm = cmeans[(size_t)c * P + p] - mm; B_over_n += dm * dm; }
    B_over_n /= (chains - 1);
  }
  const double var_plus = W * (n - 1.0) / n + B_over_n;
  if (!(var_plus > 0.0)) { ess[p] = nan(""); if (lag_used) lag_used[p] = 0.0; return; }
  double tau = -1.0, prev = INFINITY;
  int t = 0;
  for (; t + 1 <= L && t + 1 < N; t += 2) {
    const double r0 = 1.0 - (W - mean_acov(t)) / var_plus;
    const double r1 = 1.0 - (W - mean_acov(t + 1)) / var_plus;
    double pair = r0 + r1;
    if (pair < 0.0) break;
    pair = pair < prev ? pair : prev;
    tau += 2.0 * pair;
    prev = pair;
  }
  const double nm = n * chains;
  const double floor_tau = 1.0 / log10(nm > 10.0 ? nm : 10.0);
  ess[p] = nm / (tau > floor_tau ? tau : floor_tau);
  if (lag_used) lag_used[p] = (double)t;
}

// ------------------------------------------------------------------------------------------------------------
// Streaming ESS statistics: the same (acov, chain means) as k_chain_mean + k_acov_sum, accumulated while the chains
// run, so that no chain needs a trace (64 chains x 20 000 draws x 5150 parameters would be 53 GB of traces).
// Per chain and parameter, with x_0 the first retained draw and y_t = x_t - x_0 (centring on a value of the chain
// itself keeps the raw lagged products well conditioned):
//     A_l = sum_{t >= l} y_t y_(t-l)  (l = 0..L),   S = sum_t y_t,   the first L draws (head) and the last L + ES_B (ring)
// and, at the end, with m = S / N, H_l = sum_{t < l} y_t, T_l = sum_{t >= N-l} y_t:
//     sum_{t >= l} (y_t - m)(y_(t-l) - m) = A_l - m (2 S - H_l - T_l) + (N - l) m^2 .
// k_ess_stream runs once per sweep (thread per (chain, parameter)): it appends the draw to the ring and, every ES_B
// draws (and at the last draw), adds the block's lagged products with a sliding ES_B-wide register window: one ring
// load and one read-modify-write of A_l per lag and per ES_B draws, i.e. 1/ES_B of the naive per-sweep traffic.
// grid = (ceil(P/128), C), block = 128.
// ------------------------------------------------------------------------------------------------------------
constexpr int ES_B = 8;

__global__ void __launch_bounds__(128) k_ess_stream(Engine e) {
  const Dims& d = e.d;
  const int P = d.V + d.q;
  const int p = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
  if (p >= P) return;
  const long long first = e.ess_win[0], N = e.ess_win[1];
  const int L = (int)e.ess_win[2], cap = (int)e.ess_win[3];
  const long long k = (*e.iter + 1) - first;        // index of this sweep's draw inside the window
  if (k < 0 || k >= N) return;
  const double x = (p < d.V) ? e.xi[c * d.V + p] : e.gamma[(size_t)c * d.qp + (p - d.V)];
  double* sum = e.ess_sum + (size_t)c * 2 * P;
  double* ring = e.ess_ring + (size_t)c * cap * P + p;
  double ref, S;
  if (k == 0) { ref = x; S = 0.0; sum[p] = ref; }
  else { ref = sum[p]; S = sum[P + p]; }
  const double y = x - ref;
  sum[P + p] = S + y;
  ring[(size_t)(k % cap) * P] = y;
  if (k < L) e.ess_head[((size_t)c * L + k) * P + p] = y;
  const int b = (int)(k % ES_B) + 1;
  if (b != ES_B && k != N - 1) return;
  // block of the b newest draws k0 .. k: A_l += sum_j y_(k0+j) y_(k0+j-l)
  const long long k0 = k - b + 1;
  double Y[ES_B], w[ES_B];
#pragma unroll
  for (int j = 0; j < ES_B; ++j) {
    Y[j] = (j < b) ? ring[(size_t)((k0 + j) % cap) * P] : 0.0;
    w[j] = Y[j];
  }
  double* A = e.ess_acc + (size_t)c * (L + 1) * P + p;
  const int lmax = (long long)L < k ? L : (int)k;
  const bool fresh = (k0 == 0);                     // first block of the window: A starts from zero
  for (int l = 0; l <= L; ++l) {
    if (l > 0) {
#pragma unroll
      for (int j = ES_B - 1; j > 0; --j) w[j] = w[j - 1];
      w[0] = (k0 - l >= 0) ? ring[(size_t)((k0 - l) % cap) * P] : 0.0;
    }
    if (l > lmax) { if (fresh) A[(size_t)l * P] = 0.0; continue; }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < ES_B; ++j) s += Y[j] * ((k0 + j - l >= 0) ? w[j] : 0.0);
    A[(size_t)l * P] = (fresh ? 0.0 : A[(size_t)l * P]) + s;
  }
}

// acov[l][p] = sum_c (1/N) sum_{t >= l} (x_t - mean_c)(x_(t-l) - mean_c), cmean[c][p]; thread per parameter, chains in
// a fixed order (deterministic).  grid = ceil(P/128), block = 128.
__global__ void __launch_bounds__(128) k_ess_stream_finalize(Engine e, long long N, double* __restrict__ acov,
                                                             double* __restrict__ cmean) {
  const Dims& d = e.d;
  const int P = d.V + d.q, L = e.ess_L, cap = e.ess_cap;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const double n = (double)N;
  for (int c = 0; c < d.C; ++c) {
    const double ref = e.ess_sum[(size_t)c * 2 * P + p], S = e.ess_sum[(size_t)c * 2 * P + P + p];
    const double m = S / n;
    cmean[(size_t)c * P + p] = ref + m;
    const double* A = e.ess_acc + (size_t)c * (L + 1) * P + p;
    const double* head = e.ess_head + (size_t)c * L * P + p;
    const double* ring = e.ess_ring + (size_t)c * cap * P + p;
    double H = 0.0, T = 0.0;
    for (int l = 0; l <= L; ++l) {
      if (l >= 1) {
        H += head[(size_t)(l - 1) * P];
        T += ring[(size_t)((N - l) % cap) * P];
      }
      const double v = (A[(size_t)l * P] - m * (2.0 * S - H - T) + (n - l) * m * m) / n;
      acov[(size_t)l * P + p] = (c == 0 ? 0.0 : acov[(size_t)l * P + p]) + v;
    }
  }
}

void launch_ess_stream(const Engine& e, cudaStream_t s) {
  dim3 grid((e.d.V + e.d.q + 127) / 128, e.d.C);
  ++g_launches; k_ess_stream<<<grid, 128, 0, s>>>(e);
}

void launch_ess_stream_finalize(const Engine& e, long long N, double* acov, double* cmean, cudaStream_t s) {
  ++g_launches; k_ess_stream_finalize<<<(e.d.V + e.d.q + 127) / 128, 128, 0, s>>>(e, N, acov, cmean);
}

void launch_chain_mean(const double* tr, long long trace_rows, int P, int C, long long first, long long N,
                       double* cmean, cudaStream_t s) {
  dim3 grid((P + 127) / 128, C);
  ++g_launches; k_chain_mean<<<grid, 128, 0, s>>>(tr, trace_rows, P, first, N, cmean);
}

void launch_acov_sum(const double* tr, long long trace_rows, int P, int C, long long first, long long N, int L,
                     const double* cmean, double* acov, cudaStream_t s) {
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(k_acov_sum, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ACOV_SMEM); attr = true; }
  dim3 grid((P + 31) / 32, (L + 1 + AC_LB - 1) / AC_LB);
  ++g_launches; k_acov_sum<<<grid, 256, ACOV_SMEM, s>>>(tr, trace_rows, P, C, first, N, L, cmean, acov);
}

void launch_ess_finish(const double* acov_parts, int nparts, const double* cmeans, int chains, int P, long long N,
                       int L, double* ess, double* lag_used, cudaStream_t s) {
  ++g_launches; k_ess_finish<<<(P + 127) / 128, 128, 0, s>>>(acov_parts, nparts, cmeans, chains, P, N, L, ess, lag_used);
}

}  // namespace bnr
