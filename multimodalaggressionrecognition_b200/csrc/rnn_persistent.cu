// Persistent cluster GRU engine (placeholder until the kernel lands).
#include "common.cuh"
#include "rnn.cuh"

bool gru_persistent_supported(int64_t, int64_t, int64_t, int) { return false; }
int gru_fwd_persistent(const void*, const void*, const float*, void*, void*, float*, int64_t, int64_t, int64_t, cudaStream_t) {
  MAR_UNSUPPORTED("persistent GRU engine not built");
}
