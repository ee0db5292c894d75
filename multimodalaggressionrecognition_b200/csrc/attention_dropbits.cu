// Attention dropout mask: one keep bit per (b, h, query, key), 32 keys per word (layout in common.cuh, DropBits).
//
// Bit-sliced Bernoulli: 8 independent uniform 32-bit words u_0..u_7 (one counter hash each) are folded along the binary
// digits of the keep probability m / 256 = 0.d7 d6 .. d0:   x <- d_i ? (u_i | x) : (u_i & x),  i = 0 (LSB) .. 7,
// which leaves every bit of x set with probability exactly m / 256, independently of the others.  One thread per
// word; HBM-bound write of B·H·T·W words (T = 250: 16 MB per call), ALU ~80 instructions per 32 scores.
#include "common.cuh"
#include "attention.cuh"

namespace {

__global__ void __launch_bounds__(256)
attn_dropbits_kernel(uint32_t* __restrict__ words, int64_t n_words, int m, const uint64_t* __restrict__ rng, uint32_t site) {
  pdl_entry();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_words) return;
  const DropKey dk = make_drop_key(rng, site, 0.5f);           // only the key is used
  const uint64_t base = (uint64_t)i * 8u;
  uint32_t x = 0u;
#pragma unroll
  for (int d = 0; d < 8; d++) {
    const uint32_t u = drop_rand_pair(dk, base + (uint64_t)d);
    x = ((m >> d) & 1) ? (u | x) : (u & x);
  }
  words[i] = x;
}

}  // namespace

int64_t attention_dropbits_words(int64_t B, int64_t T, int64_t H) {
  return (B * H * T + 256) * drop_words_per_row(T);
}

int attention_dropbits(uint32_t* words, int64_t B, int64_t T, int64_t H, float p, const uint64_t* rng, uint32_t site,
                       cudaStream_t st) {
  const int64_t n = attention_dropbits_words(B, T, H);
  mar_launch(attn_dropbits_kernel, (unsigned)ceil_div(n, 256), 256, 0, st, words, n, drop_keep_m(p), rng, site);
  MAR_LAUNCH_CHECK("attn_dropbits");
  return MAR_OK;
}
