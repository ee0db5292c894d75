// Tiny-N linear layers (the C-class logit heads, N = class_num = 2): forward, dgrad and wgrad as
// latency-optimised SIMT kernels.  The tiled GEMM kernels are the wrong tool here: a 128x128 tile grid over
// N = 2 launches two CTAs that walk K serially (measured 119 us for M=256, K=512); these kernels spread the
// rows / columns over the whole chip and finish in a few microseconds.  fp32 accumulation, any dtype mix the
// SIMT GEMM takes.
#include "common.cuh"
#include "gemm_skinny.cuh"

namespace {

constexpr int MAXN = 16;

// out[m, n] = bias[n] + Σ_k x[m,k] w[n,k] : one warp per row m
template <typename T, typename TO>
__global__ void __launch_bounds__(256)
skinny_fwd_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ w, const float* __restrict__ bias,
                  TO* __restrict__ out, int64_t ldo, int M, int N, int K) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= M) return;
  float acc[MAXN];
#pragma unroll
  for (int n = 0; n < MAXN; n++) acc[n] = 0.f;
  const T* xr = x + (int64_t)m * ldx;
  for (int k = lane; k < K; k += 32) {
    const float xv = to_f32<T>(xr[k]);
#pragma unroll
    for (int n = 0; n < MAXN; n++)
      if (n < N) acc[n] = fmaf(xv, to_f32<T>(w[(int64_t)n * K + k]), acc[n]);
  }
#pragma unroll
  for (int n = 0; n < MAXN; n++) {
    if (n < N) {
      const float s = warp_sum(acc[n]);
      if (lane == 0) out[(int64_t)m * ldo + n] = from_f32<TO>(s + (bias ? bias[n] : 0.f));
    }
  }
}

// dx[m, k] = add[m,k] + Σ_n dz[m,n] w[n,k] : one thread per (m, k)
template <typename T>
__global__ void __launch_bounds__(256)
skinny_dgrad_kernel(const T* __restrict__ dz, const T* __restrict__ w, const T* __restrict__ add, T* __restrict__ dx,
                    int64_t lddx, int M, int N, int K) {
  pdl_entry();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)M * K) return;
  const int m = (int)(idx / K), k = (int)(idx % K);
  float s = add ? to_f32<T>(add[(int64_t)m * lddx + k]) : 0.f;
  for (int n = 0; n < N; n++) s = fmaf(to_f32<T>(dz[(int64_t)m * N + n]), to_f32<T>(w[(int64_t)n * K + k]), s);
  dx[(int64_t)m * lddx + k] = from_f32<T>(s);
}

// dw[n, k] (+)= Σ_m dz[m,n] x[m,k] : thread per k, blockIdx.y over row chunks, atomics across chunks
template <typename T>
__global__ void __launch_bounds__(256)
skinny_wgrad_kernel(const T* __restrict__ dz, const T* __restrict__ x, int64_t ldx, float* __restrict__ dw, int M, int N,
                    int K, int rows_per_block) {
  pdl_entry();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const int m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
  float acc[MAXN];
#pragma unroll
  for (int n = 0; n < MAXN; n++) acc[n] = 0.f;
  for (int m = m0; m < m1; m++) {
    const float xv = to_f32<T>(x[(int64_t)m * ldx + k]);
#pragma unroll
    for (int n = 0; n < MAXN; n++)
      if (n < N) acc[n] = fmaf(to_f32<T>(dz[(int64_t)m * N + n]), xv, acc[n]);
  }
#pragma unroll
  for (int n = 0; n < MAXN; n++)
    if (n < N) atomicAdd(dw + (int64_t)n * K + k, acc[n]);
}

}  // namespace

bool skinny_supported(int64_t N) { return N >= 1 && N <= MAXN; }

int skinny_fwd(const void* x, int64_t ldx, const void* w, const float* bias, void* out, int64_t ldo, int64_t M, int64_t N,
               int64_t K, int in_dtype, int out_dtype, cudaStream_t st) {
  mar_set_engine(MAR_ENGINE_SIMT);
  const unsigned blocks = (unsigned)ceil_div(M, 8);
  if (in_dtype == MAR_BF16 && out_dtype == MAR_F32)
    mar_launch(skinny_fwd_kernel<bf16, float>, blocks, 256, 0, st, (const bf16*)x, ldx, (const bf16*)w, bias, (float*)out, ldo, (int)M, (int)N, (int)K);
  else if (in_dtype == MAR_BF16 && out_dtype == MAR_BF16)
    mar_launch(skinny_fwd_kernel<bf16, bf16>, blocks, 256, 0, st, (const bf16*)x, ldx, (const bf16*)w, bias, (bf16*)out, ldo, (int)M, (int)N, (int)K);
  else if (in_dtype == MAR_F32 && out_dtype == MAR_F32)
    mar_launch(skinny_fwd_kernel<float, float>, blocks, 256, 0, st, (const float*)x, ldx, (const float*)w, bias, (float*)out, ldo, (int)M, (int)N, (int)K);
  else MAR_UNSUPPORTED("skinny_fwd: dtype %d -> %d", in_dtype, out_dtype);
  MAR_LAUNCH_CHECK("skinny_fwd");
  return MAR_OK;
}

int skinny_dgrad(const void* dz, const void* w, const void* add, void* dx, int64_t lddx, int64_t M, int64_t N, int64_t K,
                 int dtype, cudaStream_t st) {
  mar_set_engine(MAR_ENGINE_SIMT);
  const unsigned blocks = (unsigned)ceil_div(M * K, 256);
  if (dtype == MAR_BF16)
    mar_launch(skinny_dgrad_kernel<bf16>, blocks, 256, 0, st, (const bf16*)dz, (const bf16*)w, (const bf16*)add, (bf16*)dx, lddx, (int)M, (int)N, (int)K);
  else
    mar_launch(skinny_dgrad_kernel<float>, blocks, 256, 0, st, (const float*)dz, (const float*)w, (const float*)add, (float*)dx, lddx, (int)M, (int)N, (int)K);
  MAR_LAUNCH_CHECK("skinny_dgrad");
  return MAR_OK;
}

int skinny_wgrad(const void* dz, const void* x, int64_t ldx, float* dw, int64_t M, int64_t N, int64_t K, int dtype,
                 int accumulate, cudaStream_t st) {
  mar_set_engine(MAR_ENGINE_SIMT);
  if (!accumulate) MAR_CUDA(cudaMemsetAsync(dw, 0, (size_t)N * K * 4, st));
  const int rows_per_block = 32;
  dim3 grid((unsigned)ceil_div(K, 256), (unsigned)ceil_div(M, rows_per_block));
  if (dtype == MAR_BF16)
    mar_launch(skinny_wgrad_kernel<bf16>, grid, 256, 0, st, (const bf16*)dz, (const bf16*)x, ldx, dw, (int)M, (int)N, (int)K, rows_per_block);
  else
    mar_launch(skinny_wgrad_kernel<float>, grid, 256, 0, st, (const float*)dz, (const float*)x, ldx, dw, (int)M, (int)N, (int)K, rows_per_block);
  MAR_LAUNCH_CHECK("skinny_wgrad");
  return MAR_OK;
}
