#pragma once
#include "common.cuh"

struct SimtEpilogue {
  const float* bias = nullptr;
  const void* residual = nullptr;  // same logical shape as C
  int64_t ldr = 0;
  int res_is_bf16 = 0;
  const void* aux = nullptr;       // (M,N) ldaux, same dtype as residual: result *= (aux > 0 ? aux_scale : 0)
  int64_t ldaux = 0;
  float aux_scale = 1.f;
  int flags = 0;
  float p = 0.f;
  const uint64_t* rng = nullptr;
  uint32_t site = 0;
  int accumulate = 0;   // C += result
  int atomic = 0;       // set internally for split-K
  int allow_split = 0;  // caller allows split-K (fp32 output, linear epilogue only)
};

// C[M,N] = epilogue(A[M,K]·B[K,N]); A(m,k) at A[m*sam + k*sak], B(k,n) at B[k*sbk + n*sbn].
int gemm_simt(const void* A, int a_dtype, int64_t sam, int64_t sak, const void* B, int b_dtype, int64_t sbk,
              int64_t sbn, void* C, int c_dtype, int64_t ldc, int64_t M, int64_t N, int64_t K,
              const SimtEpilogue& epi, cudaStream_t st);
