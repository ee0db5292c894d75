// bf16 tensor-core attention over the packed in-projection output (B,T,3d): forward, dQ and dK/dV.
//
// Flash-style (scores never reach HBM): a CTA of 4 warps owns 64 query rows (forward, dQ) or 64 key rows
// (dK/dV) and streams 64-row tiles of the other operand through a double-buffered cp.async ring; every
// contraction (QKᵀ, PV, dO·Vᵀ, dS·K, P̃ᵀ·dO, dSᵀ·Q) is mma.sync.m16n8k16 bf16 with fp32 accumulators, fragments
// via ldmatrix from 16 B-padded rows (conflict-free).  Key-padding mask, per-row running max/sum (exp2 domain),
// dropout on P from the shared counter-hash (regenerated in backward), LSE saved in natural-log units so
// the SIMT engine and this one interchange.
//
// Tensor-pipe note: attention is 6.5 % of the fusion step's FLOPs at T <= 314; the tcgen05
// GEMM carries the rest.  A tcgen05/TMEM version of these kernels is the next step for the T >= 1024 sweep.
// Algorithmic FLOPs: forward 4·T²·dh per (b,h); dQ kernel 6·T²·dh; dK/dV kernel 8·T²·dh.
#include "common.cuh"
#include "attention.cuh"

namespace {

constexpr int BM = 64;    // rows owned by a CTA (16 per warp)
constexpr int BN = 64;    // rows of the streamed tile
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  const int n = valid ? 16 : 0;   // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr(dst)), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_addr(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_addr(p)));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Loads rows [r0, r0+64) x DH of a (row pitch `pitch` elements) bf16 matrix into smem [64][DH+8]; rows >= nrows zero.
template <int DH>
__device__ __forceinline__ void load_tile(bf16* s, const bf16* g, int64_t pitch, int r0, int nrows) {
  constexpr int LD = DH + 8, CH = DH / 8;
  for (int i = threadIdx.x; i < 64 * CH; i += 128) {
    const int r = i / CH, c = i % CH;
    const bool ok = r0 + r < nrows;
    cp_async16(s + r * LD + c * 8, g + (int64_t)(ok ? r0 + r : 0) * pitch + c * 8, ok);
  }
}

// A-operand fragments (16 rows x DH) of warp-owned rows [w*16, +16) from a [64][DH+8] tile.
template <int DH>
__device__ __forceinline__ void load_a_frags(uint32_t (&f)[DH / 16][4], const bf16* s, int warp, int lane) {
  constexpr int LD = DH + 8;
#pragma unroll
  for (int kk = 0; kk < DH / 16; kk++)
    ldsm_x4(f[kk], s + (warp * 16 + (lane & 15)) * LD + kk * 16 + (lane >> 4) * 8);
}

// acc[j] (16 x 8 keys, j < 8) += A(16 x DH, frags) · Xᵀ where X is a [64][DH+8] tile (rows = n index).
template <int DH>
__device__ __forceinline__ void mma_a_xt(float (&acc)[8][4], const uint32_t (&a)[DH / 16][4], const bf16* x, int lane) {
  constexpr int LD = DH + 8;
#pragma unroll
  for (int kk = 0; kk < DH / 16; kk++) {
#pragma unroll
    for (int np = 0; np < 4; np++) {
      uint32_t b[4];
      ldsm_x4(b, x + (np * 16 + (lane & 7) + (lane >> 4) * 8) * LD + kk * 16 + ((lane >> 3) & 1) * 8);
      mma16816(acc[2 * np], a[kk], b[0], b[1]);
      mma16816(acc[2 * np + 1], a[kk], b[2], b[3]);
    }
  }
}

// out[n] (16 x 8 cols, n < DH/8) += P(16 x 64, from accumulator-layout regs) · X where X is [64][DH+8] (rows = k index).
template <int DH>
__device__ __forceinline__ void mma_p_x(float (&out)[DH / 8][4], const uint32_t (&pa)[4][4], const bf16* x, int lane) {
  constexpr int LD = DH + 8;
#pragma unroll
  for (int kk = 0; kk < 4; kk++) {
#pragma unroll
    for (int np = 0; np < DH / 16; np++) {
      uint32_t b[4];
      ldsm_x4_t(b, x + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + np * 16 + (lane >> 4) * 8);
      mma16816(out[2 * np], pa[kk], b[0], b[1]);
      mma16816(out[2 * np + 1], pa[kk], b[2], b[3]);
    }
  }
}

// accumulator-layout tile (8 n-tiles x 4) -> A fragments over the 64 columns (4 k-steps)
__device__ __forceinline__ void acc_to_afrag(uint32_t (&pa)[4][4], const float (&s)[8][4]) {
#pragma unroll
  for (int kk = 0; kk < 4; kk++) {
    pa[kk][0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
    pa[kk][1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
    pa[kk][2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
    pa[kk][3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
  }
}

struct Dims { int B, T, H; };

// ==========================================================================================
// forward
// ==========================================================================================
template <int DH>
__global__ void __launch_bounds__(128)
attn_fwd_mma_kernel(const bf16* __restrict__ qkv, const uint8_t* __restrict__ key_mask, bf16* __restrict__ out,
                    float* __restrict__ lse, Dims dm, float p_drop, const uint32_t* __restrict__ dbits) {
  constexpr int LD = DH + 8;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* sQ = reinterpret_cast<bf16*>(smem_raw);
  bf16* sK = sQ + BM * LD;            // [2][BN][LD]
  bf16* sV = sK + 2 * BN * LD;        // [2][BN][LD]
  uint8_t* sM = reinterpret_cast<uint8_t*>(sV + 2 * BN * LD);   // [2][BN]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  const int T = dm.T, bh = blockIdx.y, b = bh / dm.H, h = bh % dm.H;
  const int64_t d = (int64_t)dm.H * DH, pitch = 3 * d;
  const int q0 = blockIdx.x * BM;
  const bf16* base = qkv + (int64_t)b * T * pitch + h * DH;
  const int nkv = (T + BN - 1) / BN;
  const float scale2 = rsqrtf((float)DH) * LOG2E;
  const bool drop = p_drop > 0.f;
  const DropBits dk = make_drop_bits(dbits, T, p_drop);

  auto load_kv = [&](int j, int buf) {
    load_tile<DH>(sK + buf * BN * LD, base + d, pitch, j * BN, T);
    load_tile<DH>(sV + buf * BN * LD, base + 2 * d, pitch, j * BN, T);
    if (threadIdx.x < BN) {
      const int kk = j * BN + threadIdx.x;
      sM[buf * BN + threadIdx.x] = (kk >= T || (key_mask != nullptr && key_mask[(int64_t)b * T + kk])) ? 1 : 0;
    }
  };
  load_tile<DH>(sQ, base, pitch, q0, T);
  load_kv(0, 0);
  cp_async_commit();

  uint32_t qf[DH / 16][4];
  float o[DH / 8][4];
#pragma unroll
  for (int n = 0; n < DH / 8; n++) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

  for (int j = 0; j < nkv; j++) {
    const int buf = j & 1;
    if (j + 1 < nkv) { load_kv(j + 1, buf ^ 1); cp_async_commit(); cp_async_wait<1>(); }
    else cp_async_wait<0>();
    __syncthreads();
    if (j == 0) load_a_frags<DH>(qf, sQ, warp, lane);
    const bf16* k_s = sK + buf * BN * LD;
    const bf16* v_s = sV + buf * BN * LD;
    const uint8_t* m_s = sM + buf * BN;

    float s[8][4];
#pragma unroll
    for (int n = 0; n < 8; n++) { s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f; }
    mma_a_xt<DH>(s, qf, k_s, lane);

    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int n = 0; n < 8; n++) {
      const int c = n * 8 + t4 * 2;
      const bool v0 = !m_s[c], v1 = !m_s[c + 1];
      s[n][0] = v0 ? s[n][0] * scale2 : -INFINITY;
      s[n][1] = v1 ? s[n][1] * scale2 : -INFINITY;
      s[n][2] = v0 ? s[n][2] * scale2 : -INFINITY;
      s[n][3] = v1 ? s[n][3] * scale2 : -INFINITY;
      mx[0] = fmaxf(mx[0], fmaxf(s[n][0], s[n][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[n][2], s[n][3]));
    }
    float corr[2], m_use[2];
#pragma unroll
    for (int r = 0; r < 2; r++) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);
      m_use[r] = (m_new == -INFINITY) ? 0.f : m_new;
      corr[r] = exp2f(m_run[r] - m_use[r]);     // m_run = -inf -> 0
      m_run[r] = m_new;
    }
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int n = 0; n < 8; n++) {
      s[n][0] = exp2f(s[n][0] - m_use[0]); s[n][1] = exp2f(s[n][1] - m_use[0]);
      s[n][2] = exp2f(s[n][2] - m_use[1]); s[n][3] = exp2f(s[n][3] - m_use[1]);
      rs[0] += s[n][0] + s[n][1];
      rs[1] += s[n][2] + s[n][3];
    }
    l_run[0] = l_run[0] * corr[0] + rs[0];
    l_run[1] = l_run[1] * corr[1] + rs[1];
#pragma unroll
    for (int n = 0; n < DH / 8; n++) { o[n][0] *= corr[0]; o[n][1] *= corr[0]; o[n][2] *= corr[1]; o[n][3] *= corr[1]; }
    if (drop) {
      const int64_t row0 = (int64_t)bh * T + (q0 + warp * 16 + g), row1 = row0 + 8;
#pragma unroll
      for (int n = 0; n < 8; n++) {
        const int c = j * BN + n * 8 + t4 * 2;
        bool k0 = dropbit(dk, row0, c), k1 = dropbit(dk, row0, c + 1);
        s[n][0] = k0 ? s[n][0] * dk.scale : 0.f; s[n][1] = k1 ? s[n][1] * dk.scale : 0.f;
        k0 = dropbit(dk, row1, c); k1 = dropbit(dk, row1, c + 1);
        s[n][2] = k0 ? s[n][2] * dk.scale : 0.f; s[n][3] = k1 ? s[n][3] * dk.scale : 0.f;
      }
    }
    uint32_t pa[4][4];
    acc_to_afrag(pa, s);
    mma_p_x<DH>(o, pa, v_s, lane);
    __syncthreads();
  }

  // finalize: O /= l ; stage through sQ (this warp's 16 rows) for 16 B coalesced stores
#pragma unroll
  for (int r = 0; r < 2; r++) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = l_run[0] > 0.f ? 1.f / l_run[0] : 0.f, inv1 = l_run[1] > 0.f ? 1.f / l_run[1] : 0.f;
  bf16* so = sQ + warp * 16 * LD;
#pragma unroll
  for (int n = 0; n < DH / 8; n++) {
    *reinterpret_cast<uint32_t*>(so + g * LD + n * 8 + t4 * 2) = pack_bf16x2(o[n][0] * inv0, o[n][1] * inv0);
    *reinterpret_cast<uint32_t*>(so + (g + 8) * LD + n * 8 + t4 * 2) = pack_bf16x2(o[n][2] * inv1, o[n][3] * inv1);
  }
  if (t4 == 0) {
    const int qa = q0 + warp * 16 + g, qb = qa + 8;
    if (qa < T) lse[(int64_t)bh * T + qa] = l_run[0] > 0.f ? (m_run[0] + log2f(l_run[0])) * LN2 : -INFINITY;
    if (qb < T) lse[(int64_t)bh * T + qb] = l_run[1] > 0.f ? (m_run[1] + log2f(l_run[1])) * LN2 : -INFINITY;
  }
  __syncwarp();
  constexpr int CH = DH / 8;
  for (int i = lane; i < 16 * CH; i += 32) {
    const int r = i / CH, c = i % CH;
    const int q = q0 + warp * 16 + r;
    if (q < T)
      *reinterpret_cast<uint4*>(out + ((int64_t)b * T + q) * d + h * DH + c * 8) = *reinterpret_cast<const uint4*>(so + r * LD + c * 8);
  }
}

// ==========================================================================================
// delta = rowsum(dO ⊙ O)
// ==========================================================================================
__global__ void attn_delta_bf16_kernel(const bf16* __restrict__ out, const bf16* __restrict__ dout, float* __restrict__ delta,
                                       int64_t B, int64_t T, int64_t H, int64_t dh) {
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  if (w >= B * T * H) return;
  const int64_t h = w % H, bq = w / H, b = bq / T, q = bq % T;
  const int64_t base = bq * H * dh + h * dh;
  float s = 0.f;
  for (int c = lane * 8; c < dh; c += 256) {
    float a[8], g[8];
    Vec8<bf16>::load(out + base + c, a);
    Vec8<bf16>::load(dout + base + c, g);
#pragma unroll
    for (int j = 0; j < 8; j++) s += a[j] * g[j];
  }
  s = warp_sum(s);
  if (lane == 0) delta[(b * H + h) * T + q] = s;
}

// ==========================================================================================
// dQ: CTA owns 64 queries, streams K/V tiles.   dS = P ⊙ (dP~ - delta),  dQ = dS·K / sqrt(dh)
// ==========================================================================================
template <int DH>
__global__ void __launch_bounds__(128)
attn_bwd_dq_mma_kernel(const bf16* __restrict__ qkv, const uint8_t* __restrict__ key_mask, const bf16* __restrict__ dout,
                       const float* __restrict__ lse, const float* __restrict__ delta, bf16* __restrict__ dqkv, Dims dm,
                       float p_drop, const uint32_t* __restrict__ dbits) {
  constexpr int LD = DH + 8;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* sQ = reinterpret_cast<bf16*>(smem_raw);   // Q, later dQ staging
  bf16* sO = sQ + BM * LD;                        // dO
  bf16* sK = sO + BM * LD;                        // [2][BN][LD]
  bf16* sV = sK + 2 * BN * LD;
  uint8_t* sM = reinterpret_cast<uint8_t*>(sV + 2 * BN * LD);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  const int T = dm.T, bh = blockIdx.y, b = bh / dm.H, h = bh % dm.H;
  const int64_t d = (int64_t)dm.H * DH, pitch = 3 * d;
  const int q0 = blockIdx.x * BM;
  const bf16* base = qkv + (int64_t)b * T * pitch + h * DH;
  const int nkv = (T + BN - 1) / BN;
  const float scale = rsqrtf((float)DH), scale2 = scale * LOG2E;
  const bool drop = p_drop > 0.f;
  const DropBits dk = make_drop_bits(dbits, T, p_drop);

  auto load_kv = [&](int j, int buf) {
    load_tile<DH>(sK + buf * BN * LD, base + d, pitch, j * BN, T);
    load_tile<DH>(sV + buf * BN * LD, base + 2 * d, pitch, j * BN, T);
    if (threadIdx.x < BN) {
      const int kk = j * BN + threadIdx.x;
      sM[buf * BN + threadIdx.x] = (kk >= T || (key_mask != nullptr && key_mask[(int64_t)b * T + kk])) ? 1 : 0;
    }
  };
  load_tile<DH>(sQ, base, pitch, q0, T);
  load_tile<DH>(sO, dout + (int64_t)b * T * d + h * DH, d, q0, T);
  load_kv(0, 0);
  cp_async_commit();

  const int qa = q0 + warp * 16 + g, qb = qa + 8;
  float l2[2], dl[2];   // lse in log2 units, delta
  l2[0] = qa < T ? lse[(int64_t)bh * T + qa] * LOG2E : INFINITY;   // +inf => p = 0 for rows beyond T
  l2[1] = qb < T ? lse[(int64_t)bh * T + qb] * LOG2E : INFINITY;
  dl[0] = qa < T ? delta[(int64_t)bh * T + qa] : 0.f;
  dl[1] = qb < T ? delta[(int64_t)bh * T + qb] : 0.f;
  if (l2[0] == -INFINITY) l2[0] = INFINITY;    // fully masked row: p = 0
  if (l2[1] == -INFINITY) l2[1] = INFINITY;

  uint32_t qf[DH / 16][4], of[DH / 16][4];
  float acc[DH / 8][4];
#pragma unroll
  for (int n = 0; n < DH / 8; n++) { acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f; }

  for (int j = 0; j < nkv; j++) {
    const int buf = j & 1;
    if (j + 1 < nkv) { load_kv(j + 1, buf ^ 1); cp_async_commit(); cp_async_wait<1>(); }
    else cp_async_wait<0>();
    __syncthreads();
    if (j == 0) { load_a_frags<DH>(qf, sQ, warp, lane); load_a_frags<DH>(of, sO, warp, lane); }
    const bf16* k_s = sK + buf * BN * LD;
    const bf16* v_s = sV + buf * BN * LD;
    const uint8_t* m_s = sM + buf * BN;

    float s[8][4], dp[8][4];
#pragma unroll
    for (int n = 0; n < 8; n++) {
      s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
      dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f;
    }
    mma_a_xt<DH>(s, qf, k_s, lane);
    mma_a_xt<DH>(dp, of, v_s, lane);
    const int64_t row0 = (int64_t)bh * T + (q0 + warp * 16 + g), row1 = row0 + 8;
#pragma unroll
    for (int n = 0; n < 8; n++) {
      const int c = n * 8 + t4 * 2;
      const bool v0 = !m_s[c], v1 = !m_s[c + 1];
      float p0 = v0 ? exp2f(s[n][0] * scale2 - l2[0]) : 0.f, p1 = v1 ? exp2f(s[n][1] * scale2 - l2[0]) : 0.f;
      float p2 = v0 ? exp2f(s[n][2] * scale2 - l2[1]) : 0.f, p3 = v1 ? exp2f(s[n][3] * scale2 - l2[1]) : 0.f;
      float d0 = dp[n][0], d1 = dp[n][1], d2 = dp[n][2], d3 = dp[n][3];
      if (drop) {
        bool k0 = dropbit(dk, row0, j * BN + c), k1 = dropbit(dk, row0, j * BN + c + 1);
        d0 = k0 ? d0 * dk.scale : 0.f; d1 = k1 ? d1 * dk.scale : 0.f;
        k0 = dropbit(dk, row1, j * BN + c); k1 = dropbit(dk, row1, j * BN + c + 1);
        d2 = k0 ? d2 * dk.scale : 0.f; d3 = k1 ? d3 * dk.scale : 0.f;
      }
      s[n][0] = p0 * (d0 - dl[0]); s[n][1] = p1 * (d1 - dl[0]);
      s[n][2] = p2 * (d2 - dl[1]); s[n][3] = p3 * (d3 - dl[1]);
    }
    uint32_t pa[4][4];
    acc_to_afrag(pa, s);
    mma_p_x<DH>(acc, pa, k_s, lane);
    __syncthreads();
  }

  bf16* so = sQ + warp * 16 * LD;
#pragma unroll
  for (int n = 0; n < DH / 8; n++) {
    *reinterpret_cast<uint32_t*>(so + g * LD + n * 8 + t4 * 2) = pack_bf16x2(acc[n][0] * scale, acc[n][1] * scale);
    *reinterpret_cast<uint32_t*>(so + (g + 8) * LD + n * 8 + t4 * 2) = pack_bf16x2(acc[n][2] * scale, acc[n][3] * scale);
  }
  __syncwarp();
  constexpr int CH = DH / 8;
  for (int i = lane; i < 16 * CH; i += 32) {
    const int r = i / CH, c = i % CH;
    const int q = q0 + warp * 16 + r;
    if (q < T)
      *reinterpret_cast<uint4*>(dqkv + ((int64_t)b * T + q) * pitch + h * DH + c * 8) = *reinterpret_cast<const uint4*>(so + r * LD + c * 8);
  }
}

// ==========================================================================================
// dK, dV: CTA owns 64 keys, streams Q/dO tiles.  Works on the transposed score tile Sᵀ[k,q].
//   dV = P̃ᵀ·dO,   dK = dSᵀ·Q / sqrt(dh)
// ==========================================================================================
template <int DH>
__global__ void __launch_bounds__(128)
attn_bwd_dkv_mma_kernel(const bf16* __restrict__ qkv, const uint8_t* __restrict__ key_mask, const bf16* __restrict__ dout,
                        const float* __restrict__ lse, const float* __restrict__ delta, bf16* __restrict__ dqkv, Dims dm,
                        float p_drop, const uint32_t* __restrict__ dbits) {
  constexpr int LD = DH + 8;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* sK = reinterpret_cast<bf16*>(smem_raw);   // K (own keys), later dK staging
  bf16* sV = sK + BM * LD;                        // V (own keys), later dV staging
  bf16* sQ = sV + BM * LD;                        // [2][BN][LD]
  bf16* sO = sQ + 2 * BN * LD;                    // [2][BN][LD] dO
  float* sL = reinterpret_cast<float*>(sO + 2 * BN * LD);   // [2][BN] lse (log2 units; +inf = no contribution)
  float* sD = sL + 2 * BN;                                   // [2][BN] delta
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  const int T = dm.T, bh = blockIdx.y, b = bh / dm.H, h = bh % dm.H;
  const int64_t d = (int64_t)dm.H * DH, pitch = 3 * d;
  const int k0 = blockIdx.x * BM;
  const bf16* base = qkv + (int64_t)b * T * pitch + h * DH;
  const bf16* dobase = dout + (int64_t)b * T * d + h * DH;
  const int nq = (T + BN - 1) / BN;
  const float scale = rsqrtf((float)DH), scale2 = scale * LOG2E;
  const bool drop = p_drop > 0.f;
  const DropBits dk = make_drop_bits(dbits, T, p_drop);

  auto load_q = [&](int j, int buf) {
    load_tile<DH>(sQ + buf * BN * LD, base, pitch, j * BN, T);
    load_tile<DH>(sO + buf * BN * LD, dobase, d, j * BN, T);
    if (threadIdx.x < BN) {
      const int q = j * BN + threadIdx.x;
      float l = q < T ? lse[(int64_t)bh * T + q] : INFINITY;
      if (l == -INFINITY) l = INFINITY;
      sL[buf * BN + threadIdx.x] = l * LOG2E;
      sD[buf * BN + threadIdx.x] = q < T ? delta[(int64_t)bh * T + q] : 0.f;
    }
  };
  load_tile<DH>(sK, base + d, pitch, k0, T);
  load_tile<DH>(sV, base + 2 * d, pitch, k0, T);
  load_q(0, 0);
  cp_async_commit();

  // validity of this thread's two key rows (g, g+8 of the warp's 16)
  const int ka = k0 + warp * 16 + g, kb = ka + 8;
  const bool va = ka < T && !(key_mask != nullptr && key_mask[(int64_t)b * T + ka]);
  const bool vb = kb < T && !(key_mask != nullptr && key_mask[(int64_t)b * T + kb]);

  uint32_t kf[DH / 16][4], vf[DH / 16][4];
  float dK[DH / 8][4], dV[DH / 8][4];
#pragma unroll
  for (int n = 0; n < DH / 8; n++) {
    dK[n][0] = dK[n][1] = dK[n][2] = dK[n][3] = 0.f;
    dV[n][0] = dV[n][1] = dV[n][2] = dV[n][3] = 0.f;
  }

  for (int j = 0; j < nq; j++) {
    const int buf = j & 1;
    if (j + 1 < nq) { load_q(j + 1, buf ^ 1); cp_async_commit(); cp_async_wait<1>(); }
    else cp_async_wait<0>();
    __syncthreads();
    if (j == 0) { load_a_frags<DH>(kf, sK, warp, lane); load_a_frags<DH>(vf, sV, warp, lane); }
    const bf16* q_s = sQ + buf * BN * LD;
    const bf16* o_s = sO + buf * BN * LD;
    const float* l_s = sL + buf * BN;
    const float* d_s = sD + buf * BN;

    float st[8][4], dpt[8][4];     // Sᵀ and dPᵀ tiles: rows = keys (g, g+8), cols = queries
#pragma unroll
    for (int n = 0; n < 8; n++) {
      st[n][0] = st[n][1] = st[n][2] = st[n][3] = 0.f;
      dpt[n][0] = dpt[n][1] = dpt[n][2] = dpt[n][3] = 0.f;
    }
    mma_a_xt<DH>(st, kf, q_s, lane);
    mma_a_xt<DH>(dpt, vf, o_s, lane);
    uint32_t pa[4][4], da[4][4];
    {
      float pt[8][4];
#pragma unroll
      for (int n = 0; n < 8; n++) {
        const int c = n * 8 + t4 * 2;               // query column within the tile
        const float l0 = l_s[c], l1 = l_s[c + 1], e0 = d_s[c], e1 = d_s[c + 1];
        float p0 = va ? exp2f(st[n][0] * scale2 - l0) : 0.f, p1 = va ? exp2f(st[n][1] * scale2 - l1) : 0.f;
        float p2 = vb ? exp2f(st[n][2] * scale2 - l0) : 0.f, p3 = vb ? exp2f(st[n][3] * scale2 - l1) : 0.f;
        float d0 = dpt[n][0], d1 = dpt[n][1], d2 = dpt[n][2], d3 = dpt[n][3];
        float q0v = p0, q1v = p1, q2v = p2, q3v = p3;      // P~ (dropped, scaled)
        if (drop) {
          // element (query q, key k): index (bh*T + q)*Tp + k ; the two queries of this thread are adjacent COLUMNS
          const int64_t qrow0 = (int64_t)bh * T + (j * BN + c), qrow1 = qrow0 + 1;
          const bool a0 = dropbit(dk, qrow0, (int)ka), a1 = dropbit(dk, qrow1, (int)ka);
          const bool b0 = dropbit(dk, qrow0, (int)kb), b1 = dropbit(dk, qrow1, (int)kb);
          q0v = a0 ? p0 * dk.scale : 0.f; q1v = a1 ? p1 * dk.scale : 0.f;
          q2v = b0 ? p2 * dk.scale : 0.f; q3v = b1 ? p3 * dk.scale : 0.f;
          d0 = a0 ? d0 * dk.scale : 0.f; d1 = a1 ? d1 * dk.scale : 0.f;
          d2 = b0 ? d2 * dk.scale : 0.f; d3 = b1 ? d3 * dk.scale : 0.f;
        }
        pt[n][0] = q0v; pt[n][1] = q1v; pt[n][2] = q2v; pt[n][3] = q3v;
        st[n][0] = p0 * (d0 - e0); st[n][1] = p1 * (d1 - e1);
        st[n][2] = p2 * (d2 - e0); st[n][3] = p3 * (d3 - e1);
      }
      acc_to_afrag(pa, pt);
    }
    acc_to_afrag(da, st);
    mma_p_x<DH>(dV, pa, o_s, lane);
    mma_p_x<DH>(dK, da, q_s, lane);
    __syncthreads();
  }

  bf16* sk = sK + warp * 16 * LD;
  bf16* sv = sV + warp * 16 * LD;
#pragma unroll
  for (int n = 0; n < DH / 8; n++) {
    *reinterpret_cast<uint32_t*>(sk + g * LD + n * 8 + t4 * 2) = pack_bf16x2(dK[n][0] * scale, dK[n][1] * scale);
    *reinterpret_cast<uint32_t*>(sk + (g + 8) * LD + n * 8 + t4 * 2) = pack_bf16x2(dK[n][2] * scale, dK[n][3] * scale);
    *reinterpret_cast<uint32_t*>(sv + g * LD + n * 8 + t4 * 2) = pack_bf16x2(dV[n][0], dV[n][1]);
    *reinterpret_cast<uint32_t*>(sv + (g + 8) * LD + n * 8 + t4 * 2) = pack_bf16x2(dV[n][2], dV[n][3]);
  }
  __syncwarp();
  constexpr int CH = DH / 8;
  for (int i = lane; i < 16 * CH; i += 32) {
    const int r = i / CH, c = i % CH;
    const int k = k0 + warp * 16 + r;
    if (k < T) {
      bf16* dst = dqkv + ((int64_t)b * T + k) * pitch + h * DH + c * 8;
      *reinterpret_cast<uint4*>(dst + d) = *reinterpret_cast<const uint4*>(sk + r * LD + c * 8);
      *reinterpret_cast<uint4*>(dst + 2 * d) = *reinterpret_cast<const uint4*>(sv + r * LD + c * 8);
    }
  }
}

template <int DH>
int fwd_launch(const void* qkv, const uint8_t* key_mask, void* out, float* lse, int64_t B, int64_t T, int64_t H, float p,
               const uint32_t* dbits, cudaStream_t st) {
  constexpr int LD = DH + 8;
  constexpr int smem = (BM + 4 * BN) * LD * 2 + 2 * BN;
  static bool cfg = false;
  if (!cfg) { MAR_CUDA(cudaFuncSetAttribute(attn_fwd_mma_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); cfg = true; }
  Dims dm{(int)B, (int)T, (int)H};
  dim3 grid((unsigned)ceil_div(T, BM), (unsigned)(B * H));
  attn_fwd_mma_kernel<DH><<<grid, 128, smem, st>>>((const bf16*)qkv, key_mask, (bf16*)out, lse, dm, p, dbits);
  MAR_LAUNCH_CHECK("attn_fwd_mma");
  return MAR_OK;
}

template <int DH>
int bwd_launch(const void* qkv, const uint8_t* key_mask, const void* out, const void* dout, const float* lse, float* delta,
               void* dqkv, int64_t B, int64_t T, int64_t H, float p, const uint32_t* dbits, cudaStream_t st) {
  constexpr int LD = DH + 8;
  {
    int64_t warps = B * T * H;
    attn_delta_bf16_kernel<<<(unsigned)ceil_div(warps, 8), 256, 0, st>>>((const bf16*)out, (const bf16*)dout, delta, B, T, H, DH);
    MAR_LAUNCH_CHECK("attn_delta");
  }
  Dims dm{(int)B, (int)T, (int)H};
  dim3 grid((unsigned)ceil_div(T, BM), (unsigned)(B * H));
  {
    constexpr int smem = (2 * BM + 4 * BN) * LD * 2 + 2 * BN;
    static bool cfg = false;
    if (!cfg) { MAR_CUDA(cudaFuncSetAttribute(attn_bwd_dq_mma_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); cfg = true; }
    attn_bwd_dq_mma_kernel<DH><<<grid, 128, smem, st>>>((const bf16*)qkv, key_mask, (const bf16*)dout, lse, delta, (bf16*)dqkv, dm, p, dbits);
    MAR_LAUNCH_CHECK("attn_bwd_dq_mma");
  }
  {
    constexpr int smem = (2 * BM + 4 * BN) * LD * 2 + 4 * BN * 4;
    static bool cfg = false;
    if (!cfg) { MAR_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_mma_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); cfg = true; }
    attn_bwd_dkv_mma_kernel<DH><<<grid, 128, smem, st>>>((const bf16*)qkv, key_mask, (const bf16*)dout, lse, delta, (bf16*)dqkv, dm, p, dbits);
    MAR_LAUNCH_CHECK("attn_bwd_dkv_mma");
  }
  return MAR_OK;
}

}  // namespace

bool attention_mma_supported(int64_t T, int64_t dh, int dtype) {
  return dtype == MAR_BF16 && (dh == 32 || dh == 64 || dh == 96 || dh == 128) && T >= 1 && T < (1ll << 30);
}

int attention_fwd_mma(const void* qkv, const uint8_t* key_mask, void* out, float* lse, int64_t B, int64_t T, int64_t H,
                      int64_t dh, float p, const uint32_t* dbits, cudaStream_t st) {
  MAR_CHECK_ARG(((uintptr_t)qkv % 16 == 0) && ((uintptr_t)out % 16 == 0), "attention: pointers must be 16 B aligned");
  MAR_CHECK_ARG(B * H <= 65535, "attention: B*H = %lld exceeds one launch (65535)", (long long)(B * H));
  switch (dh) {
    case 32: return fwd_launch<32>(qkv, key_mask, out, lse, B, T, H, p, dbits, st);
    case 64: return fwd_launch<64>(qkv, key_mask, out, lse, B, T, H, p, dbits, st);
    case 96: return fwd_launch<96>(qkv, key_mask, out, lse, B, T, H, p, dbits, st);
    case 128: return fwd_launch<128>(qkv, key_mask, out, lse, B, T, H, p, dbits, st);
  }
  MAR_UNSUPPORTED("attention (tensor-core engine): head dim %lld", (long long)dh);
}

int attention_bwd_mma(const void* qkv, const uint8_t* key_mask, const void* out, const void* dout, const float* lse,
                      float* delta, void* dqkv, int64_t B, int64_t T, int64_t H, int64_t dh, float p, const uint32_t* dbits,
                      cudaStream_t st) {
  MAR_CHECK_ARG(((uintptr_t)qkv % 16 == 0) && ((uintptr_t)out % 16 == 0) && ((uintptr_t)dout % 16 == 0) &&
                    ((uintptr_t)dqkv % 16 == 0), "attention: pointers must be 16 B aligned");
  switch (dh) {
    case 32: return bwd_launch<32>(qkv, key_mask, out, dout, lse, delta, dqkv, B, T, H, p, dbits, st);
    case 64: return bwd_launch<64>(qkv, key_mask, out, dout, lse, delta, dqkv, B, T, H, p, dbits, st);
    case 96: return bwd_launch<96>(qkv, key_mask, out, dout, lse, delta, dqkv, B, T, H, p, dbits, st);
    case 128: return bwd_launch<128>(qkv, key_mask, out, dout, lse, delta, dqkv, B, T, H, p, dbits, st);
  }
  MAR_UNSUPPORTED("attention (tensor-core engine): head dim %lld", (long long)dh);
}
