// lab harness: time k_gram_syrk variants on synthetic data (not part of the product)
#include "../bayesiannetworkregression.jl_b200/csrc/bnr_linalg.cu"
#include <cstdio>
#include <vector>
namespace bnr { thread_local long long g_launches = 0; }
int main() {
  using namespace bnr;
  Engine e{};
  Dims& d = e.d;
  d.n = 1000; d.V = 100; d.R = 7; d.q = 5050; d.C = 64; d.np = 1024; d.qp = 5056;
  double *X, *S, *G;
  cudaMalloc(&X, sizeof(double) * d.np * d.qp); cudaMalloc(&S, sizeof(double) * d.C * d.qp);
  cudaMalloc(&G, sizeof(double) * ((size_t)d.C * d.np * d.np + 2048));
  std::vector<double> h((size_t)d.np * d.qp, 0.25);
  cudaMemcpy(X, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice);
  std::vector<double> hs((size_t)d.C * d.qp, 1.5);
  cudaMemcpy(S, hs.data(), sizeof(double) * hs.size(), cudaMemcpyHostToDevice);
  e.X = X; e.S = S; e.G = G;
  linalg_setup();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch_syrk_G(e, 0); cudaDeviceSynchronize();
  float best = 1e9;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0); launch_syrk_G(e, 0); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
  }
  printf("%s: %.3f ms  (%.2f TFLOP/s algorithmic n^2 q)  err=%s\n", LABNAME, best, 64.0 * 1e6 * 5050 / (best * 1e-3) / 1e12,
         cudaGetErrorString(cudaGetLastError()));
  return 0;
}
