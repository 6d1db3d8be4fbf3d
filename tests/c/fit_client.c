/* Plain-C client of bnr_fit (include/bnr.h): a whole Fit!(X, y, R; mingen, maxgen, ...) -- chain generation, the
 * doubling PSRF loop, R-hat over all chains, Summary statistics -- without any host language on top, and, when the box
 * has two GPUs, the same fit sharded over both (library-side exchange of the moments) compared with one GPU.
 * Usage: fit_client [n_devices]      exit 0 = ok, 3 = no usable GPU (BNR_ENODEV), 4 = fewer GPUs than asked for. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "bnr.h"

static double lcg(unsigned long long* s) {
  *s = *s * 6364136223846793005ULL + 1442695040888963407ULL;
  return (double)((*s >> 11) & 0x1FFFFFFFFFFFFFULL) / 9007199254740992.0;
}

static int run_fit(const double* X, const double* y, int n, int V, int R, int chains_per_dev, int n_dev, double* rhat_gamma,
                   double* last_gamma, bnr_fit_info* info) {
  const int q = V * (V + 1) / 2;
  bnr_fit_params p;
  bnr_fit_default_params(&p);
  p.base.n = n; p.base.V = V; p.base.R = R; p.base.num_chains = chains_per_dev; p.base.seed = 11;
  p.mingen = 40; p.maxgen = 120; p.psrf_cutoff = 0.0;          /* never "converged": two doubling rounds */
  p.return_state = BNR_STATE_GAMMA_XI; p.n_devices = n_dev;
  bnr_fit_result* r = NULL;
  int rc = bnr_fit(&p, X, y, &r);
  if (rc != BNR_OK) { printf("bnr_fit failed %d: %s / %s\n", rc, bnr_fit_last_error(), bnr_last_error()); return rc; }
  rc = bnr_fit_get_info(r, info);
  double* tr = (double*)malloc(sizeof(double) * (size_t)info->rows * q);
  double* mean = (double*)malloc(sizeof(double) * q);
  if (rc == BNR_OK) rc = bnr_fit_rhat(r, NULL, rhat_gamma);
  if (rc == BNR_OK) rc = bnr_fit_state(r, BNR_VAR_GAMMA, tr);
  if (rc == BNR_OK) rc = bnr_fit_summary(r, mean, NULL, NULL, NULL);
  if (rc == BNR_OK)
    for (int j = 0; j < q; ++j) last_gamma[j] = tr[(info->rows - 1) + (size_t)info->rows * j];
  free(tr); free(mean);
  bnr_fit_free(r);
  return rc;
}

int main(int argc, char** argv) {
  const int want_dev = argc > 1 ? atoi(argv[1]) : 1;
  const int n = 40, V = 8, R = 3, q = V * (V + 1) / 2;
  double* X = (double*)malloc(sizeof(double) * n * q);
  double* y = (double*)malloc(sizeof(double) * n);
  unsigned long long s = 42;
  for (int j = 0; j < q; ++j)
    for (int i = 0; i < n; ++i) X[i + (size_t)n * j] = lcg(&s) < 0.5 ? 0.0 : 0.1 + lcg(&s);
  for (int i = 0; i < n; ++i) {
    y[i] = 5.0 + lcg(&s);
    for (int j = 0; j < 3; ++j) y[i] += 2.0 * X[i + (size_t)n * j];
  }
  double *rg1 = (double*)malloc(sizeof(double) * q), *rg2 = (double*)malloc(sizeof(double) * q);
  double *g1 = (double*)malloc(sizeof(double) * q), *g2 = (double*)malloc(sizeof(double) * q);
  bnr_fit_info i1, i2;
  int rc = run_fit(X, y, n, V, R, 4, 1, rg1, g1, &i1);
  if (rc == BNR_ENODEV) { printf("no GPU: %s\n", bnr_last_error()); return 3; }
  if (rc != BNR_OK) return 1;
  /* mingen = 40, maxgen = 120: 40 + 2 x 40 rows generated, the last 60 retained in a table of 80 rows */
  printf("one device : tot_generated %lld burn_in %lld sampled %lld rows %lld psrf evaluations %lld chains %d\n",
         (long long)i1.tot_generated, (long long)i1.burn_in, (long long)i1.sampled, (long long)i1.rows,
         (long long)i1.n_psrf, (int)i1.total_chains);
  if (i1.tot_generated != 120 || i1.burn_in != 20 || i1.sampled != 60 || i1.rows != 80 || i1.n_psrf != 3) return 2;
  if (want_dev >= 2) {
    rc = run_fit(X, y, n, V, R, 2, 2, rg2, g2, &i2);
    if (rc != BNR_OK) return rc == BNR_EINVAL ? 4 : 1;
    double worst = 0.0;
    for (int j = 0; j < q; ++j) {
      if (isfinite(rg1[j]) && fabs(rg2[j] - rg1[j]) > worst * fabs(rg1[j])) worst = fabs(rg2[j] - rg1[j]) / fabs(rg1[j]);
      if (g1[j] != g2[j]) { printf("chain 1 differs between 1 and 2 devices at gamma_%d\n", j); return 2; }
    }
    printf("two devices: %d chains, exchange %s, max relative R-hat difference vs one device %.2e\n", (int)i2.total_chains,
           i2.exchange == 1 ? "NCCL all-gather" : (i2.exchange == 2 ? "peer copies" : "none"), worst);
    if (i2.total_chains != 4 || worst > 1e-12) return 2;
  }
  bnr_trim_cache();
  free(X); free(y); free(rg1); free(rg2); free(g1); free(g2);
  return 0;
}
