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
    ++g_launches; k_potf2_128<<<d.C, 256, POTF2_SMEM, s>>>(e.G, cs, d.np, J, e.rhs, e.status);
    if (J + 1 < T) {
      dim3 g1(T - J - 1, d.C);
      ++g_launches; k_trsm_128<<<g1, 128, TRSM_SMEM, s>>>(e.G, cs, d.np, J);
    }
  }
}

// rhs_c <- L_c^-T rhs_c  (the forward half already happened inside launch_cholesky)
void launch_chol_solve(const Engine& e, cudaStream_t s) {
  const Dims& d = e.d;
  ++g_launches; k_trsv_bwd128<<<d.C, 256, TRSVB_SMEM, s>>>(e.G, (size_t)d.np * d.np, d.np, e.rhs);
}

}  // namespace bnr
