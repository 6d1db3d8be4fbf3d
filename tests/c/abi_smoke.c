/* Plain-C client of include/bnr.h: proves the drop-in boundary is a C ABI (no C++ / torch types), exactly what a
 * Julia `ccall`, cgo or JNI binding would see.  Usage: abi_smoke [n V R chains sweeps]
 * Exit code 0 = ran a small fit and printed its R-hat; 3 = no usable GPU (BNR_ENODEV, the loud no-fallback path). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "bnr.h"

static double lcg(unsigned long long* s) {
  *s = *s * 6364136223846793005ULL + 1442695040888963407ULL;
  return (double)((*s >> 11) & 0x1FFFFFFFFFFFFFULL) / 9007199254740992.0;
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 40, V = argc > 2 ? atoi(argv[2]) : 8, R = argc > 3 ? atoi(argv[3]) : 3;
  const int chains = argc > 4 ? atoi(argv[4]) : 4, sweeps = argc > 5 ? atoi(argv[5]) : 60;
  const int q = V * (V + 1) / 2;
  double* X = (double*)malloc(sizeof(double) * n * q);
  double* y = (double*)malloc(sizeof(double) * n);
  unsigned long long s = 42;
  for (int j = 0; j < q; ++j)
    for (int i = 0; i < n; ++i) X[i + (size_t)n * j] = lcg(&s) < 0.5 ? 0.0 : 0.1 + lcg(&s); /* column-major n x q */
  for (int i = 0; i < n; ++i) {
    y[i] = 5.0 + lcg(&s);
    for (int j = 0; j < 3 && j < q; ++j) y[i] += 2.0 * X[i + (size_t)n * j];
  }
  bnr_params p;
  bnr_default_params(&p);
  p.n = n; p.V = V; p.R = R; p.num_chains = chains; p.seed = 7; p.trace_rows = sweeps + 1;
  bnr_handle* h = NULL;
  int rc = bnr_create(&p, X, y, &h);
  if (rc == BNR_ENODEV) {
    printf("no GPU: %s\n", bnr_last_error());
    return 3;
  }
  if (rc != BNR_OK) { printf("bnr_create failed %d: %s\n", rc, bnr_last_error()); return 1; }
  int ok = bnr_init_state(h) == BNR_OK && bnr_set_moment_window(h, sweeps / 2 + 1, sweeps / 2) == BNR_OK &&
           bnr_run(h, sweeps) == BNR_OK && bnr_sync(h) == BNR_OK;
  double* rx = (double*)malloc(sizeof(double) * V);
  double* rg = (double*)malloc(sizeof(double) * q);
  double* tr = (double*)malloc(sizeof(double) * (sweeps + 1) * q);
  int32_t* st = (int32_t*)malloc(sizeof(int32_t) * chains);
  int64_t it = 0;
  int32_t mode = 0;
  ok = ok && bnr_rhat(h, rx, rg) == BNR_OK && bnr_iteration(h, &it) == BNR_OK && bnr_status(h, st) == BNR_OK &&
       bnr_get_trace(h, 0, BNR_VAR_GAMMA, 0, sweeps + 1, tr) == BNR_OK && bnr_gamma_mode(h, &mode) == BNR_OK;
  if (!ok) { printf("call failed: %s\n", bnr_last_error()); return 1; }
  double mx = 0.0;
  for (int j = 0; j < q; ++j) if (isfinite(rg[j]) && rg[j] > mx) mx = rg[j];
  printf("sweeps %lld  gamma_mode %d  max rhat(gamma) %.3f  gamma_1 after the last sweep %.4f  status[0] %d\n",
         (long long)it, (int)mode, mx, tr[sweeps], (int)st[0]);
  bnr_destroy(h);
  bnr_trim_cache();
  free(X); free(y); free(rx); free(rg); free(tr); free(st);
  return (it == sweeps && mx > 0.0) ? 0 : 2;
}
