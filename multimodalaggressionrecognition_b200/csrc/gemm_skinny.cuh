#pragma once
#include "common.cuh"

// Tiny-N (N <= 16) linear layers: the class-logit heads.
bool skinny_supported(int64_t N);
int skinny_fwd(const void* x, int64_t ldx, const void* w, const float* bias, void* out, int64_t ldo, int64_t M, int64_t N,
               int64_t K, int in_dtype, int out_dtype, cudaStream_t st);
int skinny_dgrad(const void* dz, const void* w, const void* add, void* dx, int64_t lddx, int64_t M, int64_t N, int64_t K,
                 int dtype, cudaStream_t st);
int skinny_wgrad(const void* dz, const void* x, int64_t ldx, float* dw, int64_t M, int64_t N, int64_t K, int dtype,
                 int accumulate, cudaStream_t st);
