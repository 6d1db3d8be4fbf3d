// microbenchmark: FP64 MMA shapes on sm_100a, register-resident, 8 warps/SM x 148 SMs x occupancy
#include <cstdio>
#include <cuda_runtime.h>
template<int SHAPE> __device__ __forceinline__ void mma(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  if (SHAPE == 0) { // m8n8k4
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a[0]), "d"(b[0]));
  } else if (SHAPE == 1) { // m16n8k4: A 2 regs, B 1 reg, C 4 regs
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};" : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
  } else if (SHAPE == 2) { // m16n8k8: A 4, B 2, C 4
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
  } else { // m16n8k16: A 8, B 4, C 4
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};" : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  }
}
template<int SHAPE, int NACC> __global__ void k(double* out, int iters, double seed) {
  double a[8], b[4], c[NACC][4];
  for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 1e-3 + i;
  for (int i = 0; i < 4; ++i) b[i] = seed * 0.5 + i;
  for (int j = 0; j < NACC; ++j) for (int i = 0; i < 4; ++i) c[j][i] = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < NACC; ++j) mma<SHAPE>(c[j], a, b);
  }
  double s = 0; for (int j = 0; j < NACC; ++j) for (int i = 0; i < 4; ++i) s += c[j][i];
  if (s == 12345.678) out[0] = s;
}
// plain DFMA for comparison
template<int NACC> __global__ void kf(double* out, int iters, double seed) {
  double a = seed + threadIdx.x, c[NACC];
  for (int j = 0; j < NACC; ++j) c[j] = j;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < NACC; ++j) c[j] = fma(c[j], a, seed);
  }
  double s = 0; for (int j = 0; j < NACC; ++j) s += c[j];
  if (s == 12345.678) out[0] = s;
}
template<int SHAPE, int NACC> void run(const char* name, double flops_per_mma, int warps, int blocks_per_sm) {
  double* d; cudaMalloc(&d, 8);
  int iters = 20000;
  dim3 grid(148 * blocks_per_sm), block(32 * warps);
  k<SHAPE, NACC><<<grid, block>>>(d, 100, 1.0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<SHAPE, NACC><<<grid, block>>>(d, iters, 1.0); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fl = (double)grid.x * warps * iters * NACC * flops_per_mma;
  printf("%-10s warps/blk=%d blk/SM=%d nacc=%d : %.2f TFLOP/s\n", name, warps, blocks_per_sm, NACC, fl / (ms * 1e-3) / 1e12);
  cudaFree(d);
}
int main() {
  for (int w : {4, 8, 16}) {
    if (w == 4) { run<0, 8>("m8n8k4", 512, 4, 1); run<1, 8>("m16n8k4", 1024, 4, 1); run<2, 8>("m16n8k8", 2048, 4, 1); run<3, 8>("m16n8k16", 4096, 4, 1); }
    if (w == 8) { run<0, 8>("m8n8k4", 512, 8, 1); run<1, 8>("m16n8k4", 1024, 8, 1); run<2, 8>("m16n8k8", 2048, 8, 1); run<3, 8>("m16n8k16", 4096, 8, 1); }
    if (w == 16) { run<0, 8>("m8n8k4", 512, 16, 1); run<1, 8>("m16n8k4", 1024, 16, 1); run<2, 8>("m16n8k8", 2048, 16, 1); run<3, 8>("m16n8k16", 4096, 16, 1); }
  }
  run<0, 2>("m8n8k4", 512, 8, 1); run<0, 4>("m8n8k4", 512, 8, 1); run<0, 16>("m8n8k4", 512, 8, 1);
  run<3, 2>("m16n8k16", 4096, 8, 1); run<3, 4>("m16n8k16", 4096, 8, 1);
  { double* d; cudaMalloc(&d, 8); cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kf<16><<<148 * 2, 512>>>(d, 100, 1.0);
    cudaEventRecord(e0); kf<16><<<148 * 2, 512>>>(d, 20000, 1.0); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("DFMA: %.2f TFLOP/s\n", 2.0 * 148 * 2 * 512 * 20000.0 * 16 / (ms * 1e-3) / 1e12); }
  return 0;
}
