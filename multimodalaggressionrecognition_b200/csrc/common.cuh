// Shared device/host helpers for libmar.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "../../include/mar.h"

// ---------------------------------------------------------------------------------------------
// host-side error plumbing
// ---------------------------------------------------------------------------------------------
void mar_set_error(const char* fmt, ...);
void mar_count_launch(int n = 1);
void mar_set_engine(int e);
bool mar_debug_sync();   // MAR_DEBUG_SYNC=1: synchronise after every launch and report the failing kernel by name

#define MAR_CHECK_ARG(cond, ...)                  \
  do {                                            \
    if (!(cond)) {                                \
      mar_set_error(__VA_ARGS__);                 \
      return MAR_ERR_INVALID;                     \
    }                                             \
  } while (0)

#define MAR_UNSUPPORTED(...)                      \
  do {                                            \
    mar_set_error(__VA_ARGS__);                   \
    return MAR_ERR_UNSUPPORTED;                   \
  } while (0)

// Checks the launch (not the execution: no sync) and counts it.
#define MAR_LAUNCH_CHECK(name)                                                        \
  do {                                                                                \
    cudaError_t e_ = cudaGetLastError();                                              \
    if (e_ != cudaSuccess) {                                                          \
      mar_set_error("%s: launch failed: %s", name, cudaGetErrorString(e_));           \
      return MAR_ERR_CUDA;                                                            \
    }                                                                                 \
    mar_count_launch();                                                               \
    if (mar_debug_sync()) {                                                           \
      e_ = cudaDeviceSynchronize();                                                   \
      if (e_ != cudaSuccess) {                                                        \
        mar_set_error("%s: execution failed: %s", name, cudaGetErrorString(e_));      \
        return MAR_ERR_CUDA;                                                          \
      }                                                                               \
    }                                                                                 \
  } while (0)

#define MAR_CUDA(call)                                                                \
  do {                                                                                \
    cudaError_t e_ = (call);                                                          \
    if (e_ != cudaSuccess) {                                                          \
      mar_set_error("%s failed: %s", #call, cudaGetErrorString(e_));                  \
      return MAR_ERR_CUDA;                                                            \
    }                                                                                 \
  } while (0)

int mar_sm_count();
// MAR_PDL = bit mask of the kernel classes launched with programmatic dependent launch: 1 tcgen05 GEMM, 2 LayerNorm,
// 4 elementwise / skinny GEMM / Adam, 8 attention (default: see api.cu; 0 switches it off)
bool mar_pdl_enabled(int kernel_class);
#ifndef MAR_PDL_CLASS
#define MAR_PDL_CLASS 4
#endif

// Programmatic dependent launch (sm_90+): a kernel launched with mar_launch may be STARTED while the previous kernel of the
// stream is still running — its CTAs take SMs as they free up and sit in pdl_entry() until that kernel has completed and
// flushed its memory — so the launch latency between two kernels (2-4 us of idle GPU per boundary, ~100 boundaries per
// train step) overlaps the previous kernel's tail.  pdl_entry() must come before the first global-memory access; kernels
// with a prologue that touches no global memory (barrier init, TMEM allocation) call it after that prologue.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_entry() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <typename... KArgs, typename... Args>
inline cudaError_t mar_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = mar_pdl_enabled(MAR_PDL_CLASS) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// dtype helpers: every kernel computes in fp32 and stores T ∈ {float, __nv_bfloat16}
// ---------------------------------------------------------------------------------------------
typedef __nv_bfloat16 bf16;

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// 8-element vector load/store (16 B for bf16, 32 B for fp32); pointer must be 16 B aligned.
template <typename T> struct Vec8;
template <> struct Vec8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p);
    float4 b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <> struct Vec8<bf16> {
  static __device__ __forceinline__ void load(const bf16* p, float (&v)[8]) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store(bf16* p, const float (&v)[8]) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ---------------------------------------------------------------------------------------------
// warp / block reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------------------------
// Counter-hash RNG for dropout.  mask(element) is a pure function of (seed, step, site, index):
//   key   = mix(seed, step, site)                     (once per thread)
//   r32   = lowbias32(pair_index ^ key)               (one hash per PAIR of elements)
//   keep  = u16 half of r32 >= thr16,  thr16 = round(p * 65536)
// The same function is used by every kernel (SIMT and tcgen05 GEMM epilogues, both attention
// engines, forward and backward), so a mask never has to be stored.
// lowbias32: C. Wellons' 32-bit integer hash (bias ≈ 0.1).
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
  x ^= x >> 16; x *= 0x21f0aaadu;
  x ^= x >> 15; x *= 0x735a2d97u;
  x ^= x >> 15;
  return x;
}

struct DropKey {
  uint32_t key;
  uint32_t thr16;   // element kept iff rand16 >= thr16
  float scale;      // 1 / (1 - p)
};

__device__ __forceinline__ DropKey make_drop_key(const uint64_t* rng_state, uint32_t site, float p) {
  DropKey k;
  uint64_t seed = rng_state[0], step = rng_state[1];
  uint32_t a = lowbias32((uint32_t)step + 0x9E3779B9u * (site + 1u));
  a = lowbias32((uint32_t)(seed >> 32) ^ (uint32_t)(step >> 32) ^ a);
  k.key = lowbias32((uint32_t)seed ^ a);
  float t = p * 65536.0f + 0.5f;
  k.thr16 = t >= 65535.0f ? 65535u : (uint32_t)t;
  k.scale = 1.0f / (1.0f - p);
  return k;
}

// 32 random bits for the pair of elements (2*pair, 2*pair+1); `hi` folds index bits above 2^33.
__device__ __forceinline__ uint32_t drop_rand_pair(const DropKey& k, uint64_t pair_index) {
  uint32_t lo = (uint32_t)pair_index, hi = (uint32_t)(pair_index >> 32);
  return lowbias32(lo ^ k.key ^ (hi * 0x85ebca6bu));
}
// the same 32 random bits when the pair index is known to fit 32 bits (hi == 0 folds away)
__device__ __forceinline__ uint32_t drop_rand_pair32(const DropKey& k, uint32_t pair_index) {
  return lowbias32(pair_index ^ k.key);
}
// keep flag of one element with linear index e
__device__ __forceinline__ bool drop_keep(const DropKey& k, uint64_t e) {
  uint32_t r = drop_rand_pair(k, e >> 1);
  uint32_t r16 = (e & 1) ? (r >> 16) : (r & 0xffffu);
  return r16 >= k.thr16;
}
// the same two keep flags when the pair index fits 32 bits: (r & 0xffff) >= t  <=>  (r << 16) >= (t << 16) and
// (r >> 16) >= t  <=>  r >= (t << 16), so the mask is bit-identical to drop_keep2 with 6 instructions less per pair
__device__ __forceinline__ void drop_keep2_32(const DropKey& k, uint32_t pair_index, bool& k0, bool& k1) {
  const uint32_t r = lowbias32(pair_index ^ k.key);
  const uint32_t thr_hi = k.thr16 << 16;
  k0 = (r << 16) >= thr_hi;
  k1 = r >= thr_hi;
}
// keep flags of the pair (2*pair, 2*pair+1)
__device__ __forceinline__ void drop_keep2(const DropKey& k, uint64_t pair_index, bool& k0, bool& k1) {
  uint32_t r = drop_rand_pair(k, pair_index);
  k0 = (r & 0xffffu) >= k.thr16;
  k1 = (r >> 16) >= k.thr16;
}

// ---------------------------------------------------------------------------------------------
// Attention dropout: ONE bit per (b, h, query, key), drawn once per attention call by attn_dropbits_kernel
// (attention_dropbits.cu) from the same (seed, step, site) counter hash and kept for the backward pass
// (1 bit per score: 16-25 MB per encoder layer at the reference's shapes).  Every engine — SIMT, mma.sync,
// tcgen05, forward and backward — reads the same words, so their masks agree by construction, and none of
// them spends ALU work on hashing inside the softmax loop (the hash was 8 of the forward's 13 and 11 of the
// backward's 21 instructions per score).
//   words[(bh·T + q)·W + (k >> 5)] bit (k & 31) = 1 ⇔ P[q,k] is kept;   W = 4·ceil(T / 128) words per query row.
// The keep probability is quantised to 8 bits, keep = m / 256 with m = round((1 - p)·256) — the granularity torch's
// own fused SDPA kernels use for their dropout threshold — and the kept scores are scaled by 256 / m, so the
// expectation is exact.
// ---------------------------------------------------------------------------------------------
struct DropBits {
  const uint32_t* words;
  int64_t W;        // words per (b,h,q) row
  float scale;      // 256 / m
};
__host__ __device__ __forceinline__ int drop_keep_m(float p) {
  int m = (int)((1.f - p) * 256.f + 0.5f);
  return m < 1 ? 1 : (m > 255 ? 255 : m);
}
__host__ __device__ __forceinline__ int64_t drop_words_per_row(int64_t T) { return 4 * ((T + 127) / 128); }
__device__ __forceinline__ DropBits make_drop_bits(const uint32_t* words, int64_t T, float p) {
  DropBits d;
  d.words = words;
  d.W = drop_words_per_row(T);
  d.scale = p > 0.f ? 256.f / (float)drop_keep_m(p) : 1.f;
  return d;
}
// keep flag of score (row = bh·T + q, key k).  Rows up to 256 past the end of the tensor are readable (padding).
__device__ __forceinline__ bool dropbit(const DropBits& d, int64_t row, int k) {
  return (d.words[row * d.W + (k >> 5)] >> (k & 31)) & 1u;
}
