#pragma once
#include "common.cuh"

// D[M,N] = epilogue(A·B) with reduction length Kr, bf16 operands, fp32 accumulation in TMEM.
struct TcGemmArgs {
  const void* A = nullptr;   // K-major: (M,Kr) row-major, lda ; MN-major: (Kr,M) row-major, lda
  int64_t lda = 0;
  const void* B = nullptr;   // K-major: (N,Kr) row-major, ldb ; MN-major: (Kr,N) row-major, ldb
  int64_t ldb = 0;
  int a_mn_major = 0, b_mn_major = 0;
  void* out = nullptr;       // (M,N) row-major, ldo ; bf16 or fp32
  int64_t ldo = 0;
  int out_fp32 = 0;
  int64_t M = 0, N = 0, Kr = 0;
  const float* bias = nullptr;
  const void* residual = nullptr;  // bf16 (M,N), ldr
  int64_t ldr = 0;
  const void* aux = nullptr;       // bf16 (M,N), ldaux: result *= (aux > 0 ? aux_scale : 0)  (activation backward)
  int64_t ldaux = 0;
  float aux_scale = 1.f;
  float* colsum = nullptr;         // fp32 (N): += column sums of the bf16 result (after the whole epilogue)
  int flags = 0;
  float p_drop = 0.f;
  const uint64_t* rng = nullptr;
  uint32_t site = 0;
  int accumulate = 0;   // fp32 out += result
  int allow_split = 0;  // split-K with red.add (fp32 out, linear epilogue only)
};

bool gemm_tcgen05_supported(const TcGemmArgs& a);
int gemm_tcgen05(const TcGemmArgs& a, cudaStream_t st);
