// lab harness (not part of the product): phase timings of k_potf2_inv from in-kernel globaltimer stamps, and
// event timings of the panel-chain kernels on a synthetic SPD matrix.   nvcc ... -DBNR_POTF2_STAMPS tools/lab/potf2_lab.cu
#include "../../bayesiannetworkregression.jl_b200/csrc/bnr_linalg.cu"
#include <cstdio>
#include <cmath>
#include <vector>
namespace bnr { thread_local long long g_launches = 0; }
using namespace bnr;

int main(int argc, char** argv) {
  const int C = argc > 1 ? atoi(argv[1]) : 8, N = argc > 2 ? atoi(argv[2]) : 1024;
  const int T = N / PB;
  const size_t cs = (size_t)N * N;
  double *G, *G0, *Linv, *xout; int* status;
  cudaMalloc(&xout, sizeof(double) * (size_t)C * N);
  cudaMalloc(&G, sizeof(double) * cs * C + 16384); cudaMalloc(&G0, sizeof(double) * cs * C + 16384);
  cudaMalloc(&Linv, sizeof(double) * (size_t)C * T * PB * PB); cudaMalloc(&status, sizeof(int) * C);
  cudaMemset(status, 0, sizeof(int) * C);
  cudaMemset(Linv, 0, sizeof(double) * (size_t)C * T * PB * PB);      // the zero blocks of Linv are never written
  std::vector<double> h(cs);
  for (int j = 0; j < N; ++j)
    for (int i = 0; i < N; ++i) h[(size_t)j * N + i] = (i == j) ? N + 1.0 : 0.5 + 0.3 * (((i * 31 + j * 17) % 13) / 13.0);
  for (int c = 0; c < C; ++c) cudaMemcpy(G0 + c * cs, h.data(), sizeof(double) * cs, cudaMemcpyHostToDevice);
  linalg_setup();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto timeit = [&](const char* name, auto fn) {
    float best = 1e9;
    for (int r = 0; r < 6; ++r) {
      cudaMemcpy(G, G0, sizeof(double) * cs * C, cudaMemcpyDeviceToDevice);
      cudaDeviceSynchronize();
      cudaEventRecord(e0); fn(); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("%-34s %8.1f us   (%s)\n", name, best * 1e3, cudaGetErrorString(cudaGetLastError()));
  };
  for (int J = 0; J < 2; ++J) {
    char nm[64]; snprintf(nm, 64, "k_potf2_inv J=%d late=%d", J, J > 0);
    timeit(nm, [&] { k_potf2_inv<<<C, 256, POTF2_SMEM>>>(G, cs, N, J, Linv, T, status, J > 0); });
    {
      // host check of the last chain: factor and inverse of the (late-updated) diagonal block
      const int cc = C - 1;
      std::vector<double> Dh((size_t)PB * PB), Lh((size_t)PB * PB, 0.0), Xh((size_t)PB * PB, 0.0), Ld((size_t)PB * PB), Xdv((size_t)PB * PB);
      for (int j = 0; j < PB; ++j)
        for (int i = 0; i < PB; ++i) {
          long double v = h[(size_t)(J * PB + j) * N + J * PB + i];
          if (J > 0) for (int k = 0; k < PB; ++k) v -= (long double)h[(size_t)((J - 1) * PB + k) * N + J * PB + i] * h[(size_t)((J - 1) * PB + k) * N + J * PB + j];
          Dh[(size_t)j * PB + i] = (double)v;
        }
      for (int j = 0; j < PB; ++j) {
        long double d = Dh[(size_t)j * PB + j];
        for (int k = 0; k < j; ++k) d -= (long double)Lh[(size_t)k * PB + j] * Lh[(size_t)k * PB + j];
        const long double ljj = sqrtl(d);
        Lh[(size_t)j * PB + j] = (double)ljj;
        for (int i = j + 1; i < PB; ++i) {
          long double v = Dh[(size_t)j * PB + i];
          for (int k = 0; k < j; ++k) v -= (long double)Lh[(size_t)k * PB + i] * Lh[(size_t)k * PB + j];
          Lh[(size_t)j * PB + i] = (double)(v / ljj);
        }
      }
      for (int j = 0; j < PB; ++j)                       // X = L^-1, column j by forward substitution
        for (int i = j; i < PB; ++i) {
          long double v = (i == j) ? 1.0L : 0.0L;
          for (int k = j; k < i; ++k) v -= (long double)Lh[(size_t)k * PB + i] * Xh[(size_t)j * PB + k];
          Xh[(size_t)j * PB + i] = (double)(v / Lh[(size_t)i * PB + i]);
        }
      cudaMemcpy2D(Ld.data(), PB * 8, G + cc * cs + (size_t)J * PB * N + J * PB, (size_t)N * 8, PB * 8, PB, cudaMemcpyDeviceToHost);
      cudaMemcpy(Xdv.data(), Linv + ((size_t)cc * T + J) * PB * PB, sizeof(double) * PB * PB, cudaMemcpyDeviceToHost);
      double eL = 0, eX = 0, eZ = 0, nL = 0, nX = 0;
      for (int j = 0; j < PB; ++j)
        for (int i = 0; i < PB; ++i) {
          if (i >= j) {
            eL = fmax(eL, fabs(Ld[(size_t)j * PB + i] - Lh[(size_t)j * PB + i])); nL = fmax(nL, fabs(Lh[(size_t)j * PB + i]));
            eX = fmax(eX, fabs(Xdv[(size_t)j * PB + i] - Xh[(size_t)j * PB + i])); nX = fmax(nX, fabs(Xh[(size_t)j * PB + i]));
          } else eZ = fmax(eZ, fabs(Xdv[(size_t)j * PB + i]));
        }
      printf("    check vs host (long double): factor err %.2e (max %.2e)  inverse err %.2e (max %.2e)  above-diagonal of Linv %.2e\n", eL, nL, eX, nX, eZ);
    }
#ifdef BNR_POTF2_STAMPS
    unsigned long long st[64];
    cudaMemcpyFromSymbol(st, g_potf2_stamps, sizeof(st));
    const char* names[18] = {"start", "load(+late update)", "s0 pivot block", "s0 wait for the others", "s0 next diag block",
                             "s1 pivot block", "s1 wait for the others", "s1 next diag block", "s2 pivot block",
                             "s2 wait for the others", "s2 next diag block", "s3 pivot block", "s3 wait for the others", "-",
                             "tail: last block row of the inverse", "Linv rows 64.. store", "-", "-"};
    for (int i = 1; i < 18; ++i) printf("    %-22s %7.2f us\n", names[i], (st[i] - st[i - 1]) * 1e-3);
    printf("    total                  %7.2f us\n", (st[17] - st[0]) * 1e-3);
#endif
  }
  const CUtensorMap mG1 = make_map(G, N, (uint64_t)N * C, N, SY_LDS), mG2 = make_map(G, N, (uint64_t)N * C, N, SY_BT / 2 + 4),
                    mG4 = make_map(G, N, (uint64_t)N * C, N, SY_BT / 4 + 4), mL = make_map(Linv, PB, (uint64_t)PB * T * C, PB, SY_LDS);
  for (int ns = 1; ns <= 4; ns *= 2) {
    const CUtensorMap& mI = ns == 4 ? mG4 : (ns == 2 ? mG2 : mG1);
    char nm[96];
    snprintf(nm, 96, "k_trsm_dmma 1 tile/chain, %d strips", ns);
    timeit(nm, [&] { dim3 g(C, ns); k_trsm_dmma<<<g, SY_THREADS, SYRK_SMEM>>>(mL, mI, G, cs, N, N, 0, 1, ns); });
    snprintf(nm, 96, "k_trsm_dmma T-2 tiles/chain, %d strips", ns);
    timeit(nm, [&] { dim3 g(C, (T - 2) * ns); k_trsm_dmma<<<g, SY_THREADS, SYRK_SMEM>>>(mL, mI, G, cs, N, N, 0, 2, ns); });
    snprintf(nm, 96, "k_chol_update depth 128, T-2 tiles, %d strips", ns);
    timeit(nm, [&] { dim3 g(C, (T - 2) * ns); k_chol_update<<<g, SY_THREADS, SYRK_SMEM>>>(mG1, mI, 0, G, cs, N, N, 8, 1, 2, ns); });
    const int jb = 4 < T ? 4 : T - 1;
    snprintf(nm, 96, "k_chol_update depth 512, col %d incl diag, %d strips", jb, ns);
    timeit(nm, [&] { dim3 g(C, (T - jb) * ns); k_chol_update<<<g, SY_THREADS, SYRK_SMEM>>>(mG1, mI, 0, G, cs, N, N, 32, jb, jb, ns); });
  }
  const CUtensorMap bG = make_map(G, N, (uint64_t)N * C, N, PB, BW_COLS), bL = make_map(Linv, PB, (uint64_t)PB * T * C, PB, PB, BW_COLS);
  timeit("k_bwd_stream", [&] { k_bwd_stream<<<C, BW_THREADS, bwd_smem(N)>>>(bG, bL, G, cs, N, N - 1, xout, N, nullptr, 0, bwd_stages(N)); });
  timeit("empty launch pair", [&] { k_augment<<<1, 256>>>(G, cs, N, 1, G0, N, nullptr, nullptr); });
  return 0;
}
