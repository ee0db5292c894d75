// bf16 tensor-core attention engine (placeholder until the kernel lands: reports "unsupported" so AUTO
// selects the SIMT engine; an explicit request for this engine fails loudly).
#include "common.cuh"
#include "attention.cuh"

bool attention_mma_supported(int64_t, int64_t, int) { return false; }
int attention_fwd_mma(const void*, const uint8_t*, void*, float*, int64_t, int64_t, int64_t, int64_t, float,
                      const uint64_t*, uint32_t, cudaStream_t) {
  MAR_UNSUPPORTED("attention tensor-core engine not built");
}
int attention_bwd_mma(const void*, const uint8_t*, const void*, const void*, const float*, float*, void*, int64_t,
                      int64_t, int64_t, int64_t, float, const uint64_t*, uint32_t, cudaStream_t) {
  MAR_UNSUPPORTED("attention tensor-core engine not built");
}
