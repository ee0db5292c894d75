#pragma once
#include "common.cuh"

// SIMT fp32-math engine (attention_simt.cu)
int attention_fwd_simt(const void* qkv, const uint8_t* key_mask, void* out, float* lse, int64_t B, int64_t T, int64_t H,
                       int64_t dh, int dtype, float p, const uint32_t* dbits, cudaStream_t st);
int attention_bwd_simt(const void* qkv, const uint8_t* key_mask, const void* out, const void* dout, const float* lse,
                       float* delta, void* dqkv, int64_t B, int64_t T, int64_t H, int64_t dh, int dtype, float p,
                       const uint32_t* dbits, cudaStream_t st);

// bf16 tensor-core engine (attention_mma.cu); returns MAR_ERR_UNSUPPORTED for shapes it does not take
bool attention_mma_supported(int64_t T, int64_t dh, int dtype);
int attention_fwd_mma(const void* qkv, const uint8_t* key_mask, void* out, float* lse, int64_t B, int64_t T, int64_t H,
                      int64_t dh, float p, const uint32_t* dbits, cudaStream_t st);
int attention_bwd_mma(const void* qkv, const uint8_t* key_mask, const void* out, const void* dout, const float* lse,
                      float* delta, void* dqkv, int64_t B, int64_t T, int64_t H, int64_t dh, float p,
                      const uint32_t* dbits, cudaStream_t st);

// tcgen05 / TMEM / TMA engine (attention_tc.cu), bf16, dh in {64, 96, 128}
bool attention_tc_supported(int64_t B, int64_t T, int64_t H, int64_t dh, int dtype);
int attention_fwd_tc(const void* qkv, const uint8_t* key_mask, void* out, float* lse, int64_t B, int64_t T, int64_t H,
                     int64_t dh, float p, const uint32_t* dbits, cudaStream_t st);
// the same for T > 128: two query tiles per CTA, two softmax warp-groups (attention_tc2.cu)
int attention_fwd_tc2(const void* qkv, const uint8_t* key_mask, void* out, float* lse, int64_t B, int64_t T, int64_t H,
                      int64_t dh, float p, const uint32_t* dbits, cudaStream_t st);
// backward: `work` holds attention_bwd_tc_work_floats() floats (delta, then the fp32 dQ accumulator when T > 128)
int64_t attention_bwd_tc_work_floats(int64_t B, int64_t T, int64_t H, int64_t dh);
// dbias (3d fp32, or nullptr): += column sums of dqkv over all tokens, taken inside the kernel
int attention_bwd_tc(const void* qkv, const uint8_t* key_mask, const void* out, const void* dout, const float* lse,
                     float* work, void* dqkv, float* dbias, int64_t B, int64_t T, int64_t H, int64_t dh, float p,
                     const uint32_t* dbits, cudaStream_t st);

// dropout keep bits of one attention call (attention_dropbits.cu; layout: common.cuh DropBits)
int64_t attention_dropbits_words(int64_t B, int64_t T, int64_t H);
int attention_dropbits(uint32_t* words, int64_t B, int64_t T, int64_t H, float p, const uint64_t* rng, uint32_t site,
                       cudaStream_t st);
