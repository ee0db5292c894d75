// tcgen05 / TMEM / TMA attention over the packed in-projection output (B,T,3d): backward, ONE kernel.
//
//   P = exp(S - LSE),  P̃ = dropout(P),  dV = P̃ᵀ·dO,  dP = dropoutᵀ(dO·Vᵀ),  dS = P ⊙ (dP - delta),
//   dQ = dS·K / √dh,   dK = dSᵀ·Q / √dh                          delta = rowsum(dO ⊙ O)
//
// One CTA owns 128 KEYS of one (b,h) and walks over the query tiles (128 rows each); everything is computed on
// the TRANSPOSED score tile so that the probability operands sit in TMEM with keys as lanes:
//   Sᵀ  = K·Qᵀ       tcgen05.mma  M=128 keys, N=queries, K=dh    A = K tile, B = Q tile (both K-major smem)
//   dPᵀ = V·dOᵀ      tcgen05.mma  same shape                     A = V tile, B = dO tile
//   dV += P̃ᵀ·dO      tcgen05.mma  M=128 keys, N=dh, K=queries    A = P̃ᵀ bf16 in TMEM,  B = dO tile (MN-major)
//   dK += dSᵀ·Q      tcgen05.mma  same shape                     A = dSᵀ bf16 in TMEM, B = Q tile  (MN-major)
//   dQ  = dS·K       tcgen05.mma  M=128 queries, N=dh, K=keys    A = dSᵀ bf16 in smem (MN-major), B = K tile (MN-major)
// dK / dV accumulate in TMEM over the whole walk; dQ tiles are reduced across the key tiles of a (b,h) with
// red.global.add.v4.f32 into an fp32 (B,T,d) workspace (converted to bf16 by attn_dq_convert_kernel), or
// written straight to dqkv when T <= 128 (one key tile).  5 contractions per tile pair, nothing recomputed.
//
// 18 warps: warp 0 lane 0 = TMA producer + MMA issuer; warp 1 = TMEM allocator + LSE/delta stager;
// warps 2-17 = 16 compute warps, thread <-> key row (TMEM lane), each warp owns one 32-query column chunk of one
// TMEM lane quarter.  Q / dO tiles are double-buffered by TMA (3-D maps, rows >= T zero-filled).
//
// TMEM columns (512): Sᵀ [0,128) (P̃ᵀ packed bf16 in [0,64)) | dPᵀ [128,256) (dSᵀ packed bf16 in [192,256)) |
//                     dQ [64,64+dh) (aliases the dead halves of Sᵀ/dPᵀ) | dK [256,256+dh) | dV [384,384+dh).
// Dropout: the shared counter hash with element index ((b·H+h)·T + q)·Tp + k (common.cuh), i.e. the mask the
// forward of ANY engine drew.  Fully masked rows (LSE = -inf) contribute nothing.
#include "common.cuh"
#include "ptx_sm100.cuh"
#include "attention.cuh"
#include "attention_tc.cuh"

using namespace sm100;
using namespace attn_tc;

namespace {

constexpr int BT = 128;                 // keys per CTA and queries per step
constexpr int NCOMPUTE = 16;            // compute warps
constexpr int NTHREADS = (2 + NCOMPUTE) * 32;
constexpr uint32_t COL_S = 0, COL_P = 0, COL_DP = 128, COL_DS = 192, COL_DQ = 64, COL_DK = 256, COL_DV = 384;

struct BwdParams {
  const uint8_t* key_mask;
  const float* lse;
  const float* delta;
  float* dq_acc;      // (B,T,d) fp32, zero-initialised; nullptr when T <= 128 (dQ goes straight to dqkv)
  bf16* dqkv;         // (B,T,3d)
  int B, T, H;
  float p_drop;
  const uint64_t* rng;
  uint32_t site;
  int smem_bytes;
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int DH>
__global__ void __launch_bounds__(NTHREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do, const BwdParams p) {
  constexpr int NBOX = (DH + 63) / 64;
  constexpr int OP_BYTES = NBOX * BOX_BYTES;
  constexpr int KSTEPS = DH / 16;
  constexpr int DS_BYTES = 2 * BOX_BYTES;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + OP_BYTES;
  uint8_t* sQ = sV + OP_BYTES;            // [2][OP_BYTES]
  uint8_t* sDO = sQ + 2 * OP_BYTES;       // [2][OP_BYTES]
  uint8_t* sDS = sDO + 2 * OP_BYTES;      // dSᵀ bf16 [2 atoms of 64 queries][128 keys][64], 128 B swizzle
  float* sL = reinterpret_cast<float*>(sDS + DS_BYTES);   // [2][128] LSE in log2 units (+inf = no contribution)
  float* sD = sL + 2 * BT;                                // [2][128] delta
  uint64_t* bars = reinterpret_cast<uint64_t*>(sD + 2 * BT);
  uint64_t* bar_kv = bars + 0;
  uint64_t* bar_qdo = bars + 1;       // [2] Q_i / dO_i tiles landed
  uint64_t* bar_ld = bars + 3;        // [2] LSE / delta of tile i staged
  uint64_t* bar_ldfree = bars + 5;    // [2] compute warps are done with that LSE / delta stage
  uint64_t* bar_s = bars + 7;         // Sᵀ, dPᵀ complete in TMEM
  uint64_t* bar_pds = bars + 8;       // P̃ᵀ, dSᵀ written (TMEM + smem) by the compute warps
  uint64_t* bar_mma2 = bars + 9;      // dV, dK, dQ MMAs complete (Q_i / dO_i smem stage free)
  uint64_t* bar_dqr = bars + 10;      // dQ tile read out of TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);
  if (reinterpret_cast<uint8_t*>(tmem_slot + 2) > smem_raw + p.smem_bytes) __trap();   // dynamic smem base less aligned than assumed

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = p.T, H = p.H;
  const int n_t = (T + BT - 1) / BT;      // key tiles == query tiles
  const int bh = blockIdx.x / n_t, kt = blockIdx.x % n_t;
  const int b = bh / H, h = bh % H;
  const int d = H * DH;
  const int nk = min(BT, T - kt * BT);    // keys of this tile inside the sequence

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tm_qkv);
    prefetch_tensormap(&tm_do);
    mbar_init(bar_kv, 1);
    mbar_init(bar_qdo + 0, 1); mbar_init(bar_qdo + 1, 1);
    mbar_init(bar_ld + 0, 1); mbar_init(bar_ld + 1, 1);
    mbar_init(bar_ldfree + 0, NCOMPUTE); mbar_init(bar_ldfree + 1, NCOMPUTE);
    mbar_init(bar_s, 1);
    mbar_init(bar_pds, NCOMPUTE);
    mbar_init(bar_mma2, 1);
    mbar_init(bar_dqr, NCOMPUTE);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------------------------------------------------------- TMA producer + MMA issuer
      auto load_tile = [&](uint8_t* dst, const CUtensorMap* map, uint64_t* bar, int col0, int row0) {
#pragma unroll
        for (int bx = 0; bx < NBOX; bx++) tma_load_3d(dst + bx * BOX_BYTES, map, bar, col0 + bx * 64, row0, b);
      };
      auto load_qdo = [&](int i) {
        const int s = i & 1;
        mbar_expect_tx(bar_qdo + s, 2 * OP_BYTES);
        load_tile(sQ + s * OP_BYTES, &tm_qkv, bar_qdo + s, h * DH, i * BT);
        load_tile(sDO + s * OP_BYTES, &tm_do, bar_qdo + s, h * DH, i * BT);
      };
      mbar_expect_tx(bar_kv, 2 * OP_BYTES);
      load_tile(sK, &tm_qkv, bar_kv, d + h * DH, kt * BT);
      load_tile(sV, &tm_qkv, bar_kv, 2 * d + h * DH, kt * BT);
      load_qdo(0);
      if (n_t > 1) load_qdo(1);
      mbar_wait(bar_kv, 0);

      constexpr uint32_t idesc_acc = make_idesc_bf16(BT, DH, 0, 1);   // dV, dK: A in TMEM, B MN-major
      constexpr uint32_t idesc_dq = make_idesc_bf16(BT, DH, 1, 1);    // dQ: A and B MN-major smem
      const int ksteps_keys = (nk + 15) / 16;
      for (int i = 0; i < n_t; i++) {
        const int s = i & 1;
        const int nq = min(BT, T - i * BT);
        mbar_wait(bar_qdo + s, (i >> 1) & 1);
        if (i > 0) mbar_wait(bar_dqr, (i - 1) & 1);     // dQ_{i-1} has left the columns Sᵀ_i will overwrite
        tc_fence_after();
        const uint32_t idesc_s = make_idesc_bf16(BT, (nq + 15) & ~15, 0, 0);
        const uint8_t* q_s = sQ + s * OP_BYTES;
        const uint8_t* do_s = sDO + s * OP_BYTES;
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ks++)
          umma_f16(tmem_base + COL_S, make_desc_kmajor(smem_u32(sK + (ks / 4) * BOX_BYTES), ks % 4),
                   make_desc_kmajor(smem_u32(q_s + (ks / 4) * BOX_BYTES), ks % 4), idesc_s, ks > 0 ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ks++)
          umma_f16(tmem_base + COL_DP, make_desc_kmajor(smem_u32(sV + (ks / 4) * BOX_BYTES), ks % 4),
                   make_desc_kmajor(smem_u32(do_s + (ks / 4) * BOX_BYTES), ks % 4), idesc_s, ks > 0 ? 1u : 0u);
        umma_commit(bar_s);

        mbar_wait(bar_pds, i & 1);
        tc_fence_after();
        const int qsteps = (nq + 15) / 16;
        for (int ks = 0; ks < qsteps; ks++)
          umma_f16_ts(tmem_base + COL_DV, tmem_base + COL_P + ks * 8, make_desc_mnmajor(smem_u32(do_s), ks, BOX_BYTES),
                      idesc_acc, (i > 0 || ks > 0) ? 1u : 0u);
        for (int ks = 0; ks < qsteps; ks++)
          umma_f16_ts(tmem_base + COL_DK, tmem_base + COL_DS + ks * 8, make_desc_mnmajor(smem_u32(q_s), ks, BOX_BYTES),
                      idesc_acc, (i > 0 || ks > 0) ? 1u : 0u);
        for (int ks = 0; ks < ksteps_keys; ks++)
          umma_f16(tmem_base + COL_DQ, make_desc_mnmajor(smem_u32(sDS), ks, BOX_BYTES),
                   make_desc_mnmajor(smem_u32(sK), ks, BOX_BYTES), idesc_dq, ks > 0 ? 1u : 0u);
        umma_commit(bar_mma2);
        if (i + 2 < n_t) {
          mbar_wait(bar_mma2, i & 1);       // stage s is free again
          load_qdo(i + 2);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- LSE / delta stager
    for (int i = 0; i < n_t; i++) {
      const int s = i & 1;
      if (i >= 2) mbar_wait(bar_ldfree + s, ((i >> 1) - 1) & 1);
      for (int k = lane; k < BT; k += 32) {
        const int q = i * BT + k;
        float l = INFINITY, dl = 0.f;
        if (q < T) {
          l = p.lse[(int64_t)bh * T + q] * LOG2E;
          if (l == -INFINITY) l = INFINITY;        // fully masked row: P = 0
          dl = p.delta[(int64_t)bh * T + q];
        }
        sL[s * BT + k] = l;
        sD[s * BT + k] = dl;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_ld + s);
    }
  } else {
    // ---------------------------------------------------------------- compute warps: thread <-> key row
    const int quarter = warp & 3;                       // TMEM lanes this warp may access
    const int chunk = (warp - 2) >> 2;                  // 32-query column chunk
    const int row = quarter * 32 + lane;
    const int key = kt * BT + row;
    const bool kvalid = key < T && !(p.key_mask != nullptr && p.key_mask[(int64_t)b * T + key] != 0);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const float scale = rsqrtf((float)DH), scale2 = scale * LOG2E;
    const bool drop = p.p_drop > 0.f;
    DropKey dk;
    dk.key = 0; dk.thr16 = 0; dk.scale = 1.f;
    if (drop) dk = make_drop_key(p.rng, p.site, p.p_drop);
    const uint64_t Tp = (uint64_t)((T + 1) & ~1);
    const uint64_t half_tp = Tp >> 1;
    const bool hi_half = (key & 1) != 0;
    // dSᵀ smem row of this key: [atom = chunk/2][row][64 queries], 16 B chunks XOR-swizzled by (row % 8)
    uint8_t* ds_row = sDS + (chunk >> 1) * BOX_BYTES + row * 128;
    const int ds_c0 = (chunk & 1) * 4;
    const int d3 = 3 * d;

    for (int i = 0; i < n_t; i++) {
      const int s = i & 1;
      const int nq = min(BT, T - i * BT);
      const bool active = chunk * 32 < nq;
      const float* l_s = sL + s * BT + chunk * 32;
      const float* d_s = sD + s * BT + chunk * 32;
      uint32_t pk[16], dsk[16];
      mbar_wait(bar_ld + s, (i >> 1) & 1);
      mbar_wait(bar_s, i & 1);
      tc_fence_after();
      if (active) {
        // pair index of element (q, key): (((bh*T + q) * Tp) >> 1) + (key >> 1)
        uint64_t pair = (((uint64_t)bh * (uint64_t)T + (uint64_t)(i * BT + chunk * 32)) * Tp >> 1) + (uint64_t)(key >> 1);
#pragma unroll
        for (int hf = 0; hf < 2; hf++) {
          uint32_t rs[16], rp[16];
          tmem_ld_32x32b_x16(lane_addr + COL_S + chunk * 32 + hf * 16, rs);
          tmem_ld_32x32b_x16(lane_addr + COL_DP + chunk * 32 + hf * 16, rp);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            float pv[2], pd[2], ds[2];
#pragma unroll
            for (int u = 0; u < 2; u++) {
              const int ql = hf * 16 + e + u;
              const bool ok = kvalid && (chunk * 32 + ql < nq);
              const float pr = ok ? ex2f(fmaf(__uint_as_float(rs[e + u]), scale2, -l_s[ql])) : 0.f;
              float dp = __uint_as_float(rp[e + u]);
              float pdv = pr;
              if (drop) {
                const uint32_t r = drop_rand_pair(dk, pair);
                const bool keep = (hi_half ? (r >> 16) : (r & 0xffffu)) >= dk.thr16;
                pdv = keep ? pr * dk.scale : 0.f;
                dp = keep ? dp * dk.scale : 0.f;
                pair += half_tp;
              }
              pv[u] = pr; pd[u] = pdv;
              ds[u] = ok ? pv[u] * (dp - d_s[ql]) : 0.f;
            }
            pk[hf * 8 + e / 2] = pack_bf16x2(pd[0], pd[1]);
            dsk[hf * 8 + e / 2] = pack_bf16x2(ds[0], ds[1]);
          }
        }
      }
      // every warp of this lane quarter has read its Sᵀ / dPᵀ columns: the packed results may overwrite them
      named_bar_sync(1 + quarter, 128);
      if (active) {
        tmem_st_32x32b_x16(lane_addr + COL_P + chunk * 16, pk);
        tmem_st_32x32b_x16(lane_addr + COL_DS + chunk * 16, dsk);
#pragma unroll
        for (int c4 = 0; c4 < 4; c4++) {
          const int phys = (ds_c0 + c4) ^ (row & 7);
          *reinterpret_cast<uint4*>(ds_row + phys * 16) = make_uint4(dsk[c4 * 4], dsk[c4 * 4 + 1], dsk[c4 * 4 + 2], dsk[c4 * 4 + 3]);
        }
        tmem_st_wait();
        fence_proxy_async();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(bar_pds); mbar_arrive(bar_ldfree + s); }

      // dQ tile: thread <-> query row
      mbar_wait(bar_mma2, i & 1);
      tc_fence_after();
      if (chunk * 32 < DH) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(lane_addr + COL_DQ + chunk * 32, r);
        tmem_ld_wait();
        const int q = i * BT + row;
        if (q < T) {
          if (p.dq_acc != nullptr) {
            float* dst = p.dq_acc + ((int64_t)b * T + q) * d + h * DH + chunk * 32;
#pragma unroll
            for (int g = 0; g < 8; g++)
              red_add_v4(dst + g * 4, __uint_as_float(r[g * 4]) * scale, __uint_as_float(r[g * 4 + 1]) * scale,
                         __uint_as_float(r[g * 4 + 2]) * scale, __uint_as_float(r[g * 4 + 3]) * scale);
          } else {
            bf16* dst = p.dqkv + ((int64_t)b * T + q) * d3 + h * DH + chunk * 32;
#pragma unroll
            for (int g = 0; g < 4; g++) {
              uint4 u;
              u.x = pack_bf16x2(__uint_as_float(r[g * 8 + 0]) * scale, __uint_as_float(r[g * 8 + 1]) * scale);
              u.y = pack_bf16x2(__uint_as_float(r[g * 8 + 2]) * scale, __uint_as_float(r[g * 8 + 3]) * scale);
              u.z = pack_bf16x2(__uint_as_float(r[g * 8 + 4]) * scale, __uint_as_float(r[g * 8 + 5]) * scale);
              u.w = pack_bf16x2(__uint_as_float(r[g * 8 + 6]) * scale, __uint_as_float(r[g * 8 + 7]) * scale);
              *reinterpret_cast<uint4*>(dst + g * 8) = u;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_dqr);
    }

    // epilogue: dK (scaled) and dV of this key row -> dqkv   (the last bar_mma2 wait above covers them)
    if (chunk * 32 < DH) {
      bf16* dst = p.dqkv + ((int64_t)b * T + key) * d3 + h * DH + chunk * 32;
#pragma unroll
      for (int which = 0; which < 2; which++) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(lane_addr + (which == 0 ? COL_DK : COL_DV) + chunk * 32, r);   // warp-collective
        tmem_ld_wait();
        const float sc = which == 0 ? scale : 1.f;
        bf16* o = dst + (which == 0 ? d : 2 * d);
        if (key < T) {
#pragma unroll
          for (int g = 0; g < 4; g++) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(r[g * 8 + 0]) * sc, __uint_as_float(r[g * 8 + 1]) * sc);
            u.y = pack_bf16x2(__uint_as_float(r[g * 8 + 2]) * sc, __uint_as_float(r[g * 8 + 3]) * sc);
            u.z = pack_bf16x2(__uint_as_float(r[g * 8 + 4]) * sc, __uint_as_float(r[g * 8 + 5]) * sc);
            u.w = pack_bf16x2(__uint_as_float(r[g * 8 + 6]) * sc, __uint_as_float(r[g * 8 + 7]) * sc);
            *reinterpret_cast<uint4*>(o + g * 8) = u;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// delta (B,H,T) = rowsum over dh of dO ⊙ O.  One block per 4 token rows: thread <-> 8 columns, then H threads
// per row sum the dh/8 partials of their head.
constexpr int DELTA_ROWS = 8;
__global__ void __launch_bounds__(256)
attn_delta_tc_kernel(const bf16* __restrict__ out, const bf16* __restrict__ dout, float* __restrict__ delta, int64_t BT_rows,
                     int T, int H, int dh) {
  extern __shared__ float part[];                 // [DELTA_ROWS][d/8]
  const int d = H * dh, c8n = d / 8;
  const int64_t row0 = (int64_t)blockIdx.x * DELTA_ROWS;
  for (int i = threadIdx.x; i < DELTA_ROWS * c8n; i += blockDim.x) {
    const int r = i / c8n, c = i % c8n;
    float sacc = 0.f;
    if (row0 + r < BT_rows) {
      float a[8], g[8];
      Vec8<bf16>::load(out + (row0 + r) * d + c * 8, a);
      Vec8<bf16>::load(dout + (row0 + r) * d + c * 8, g);
#pragma unroll
      for (int j = 0; j < 8; j++) sacc += a[j] * g[j];
    }
    part[i] = sacc;
  }
  __syncthreads();
  const int per = dh / 8;
  for (int i = threadIdx.x; i < DELTA_ROWS * H; i += blockDim.x) {
    const int r = i / H, hh = i % H;
    if (row0 + r >= BT_rows) continue;
    float sacc = 0.f;
    for (int j = 0; j < per; j++) sacc += part[r * c8n + hh * per + j];
    const int64_t bq = row0 + r, bb = bq / T, q = bq % T;
    delta[(bb * H + hh) * T + q] = sacc;
  }
}

// dqkv[:, :, 0:d] (bf16) = dq_acc (fp32)
__global__ void __launch_bounds__(256)
attn_dq_convert_kernel(const float* __restrict__ acc, bf16* __restrict__ dqkv, int64_t rows, int d) {
  const int c8n = d / 8;
  const int64_t n = rows * c8n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / c8n;
    const int c = (int)(i % c8n);
    float v[8];
    Vec8<float>::load(acc + r * d + c * 8, v);
    Vec8<bf16>::store(dqkv + r * 3 * d + c * 8, v);
  }
}

template <int DH>
int bwd_launch(const void* qkv, const uint8_t* key_mask, const void* out, const void* dout, const float* lse, float* work,
               void* dqkv, int64_t B, int64_t T, int64_t H, float p, const uint64_t* rng, uint32_t site, cudaStream_t st) {
  constexpr int NBOX = (DH + 63) / 64;
  constexpr int USED = 6 * NBOX * BOX_BYTES + 2 * BOX_BYTES + 4 * BT * 4 + 96;
  constexpr int SMEM = USED + 1024 <= 232448 ? USED + 1024 : 232448;   // dh >= 96: 928 B of alignment slack (kernel traps if short)
  static_assert(USED <= 232448, "attention backward: shared memory budget");
  static bool cfg = false;
  if (!cfg) {
    MAR_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    cfg = true;
  }
  const int64_t d = H * DH;
  float* delta = work;
  float* dq_acc = T > BT ? work + ((B * H * T + 3) & ~(int64_t)3) : nullptr;
  {
    const int64_t rows = B * T;
    attn_delta_tc_kernel<<<(unsigned)ceil_div(rows, DELTA_ROWS), 256, DELTA_ROWS * (d / 8) * sizeof(float), st>>>(
        (const bf16*)out, (const bf16*)dout, delta, rows, (int)T, (int)H, DH);
    MAR_LAUNCH_CHECK("attn_delta_tc");
  }
  if (dq_acc != nullptr) MAR_CUDA(cudaMemsetAsync(dq_acc, 0, (size_t)(B * T * d) * sizeof(float), st));
  CUtensorMap tm_qkv, tm_do;
  int rc = make_map_btc(&tm_qkv, qkv, B, T, 3 * d, BT);
  if (rc) return rc;
  rc = make_map_btc(&tm_do, dout, B, T, d, BT);
  if (rc) return rc;
  BwdParams prm;
  prm.key_mask = key_mask; prm.lse = lse; prm.delta = delta; prm.dq_acc = dq_acc; prm.dqkv = (bf16*)dqkv;
  prm.B = (int)B; prm.T = (int)T; prm.H = (int)H; prm.p_drop = p; prm.rng = rng; prm.site = site; prm.smem_bytes = SMEM;
  const int64_t n_t = ceil_div(T, BT);
  attn_bwd_tc_kernel<DH><<<(unsigned)(B * H * n_t), NTHREADS, SMEM, st>>>(tm_qkv, tm_do, prm);
  MAR_LAUNCH_CHECK("attn_bwd_tc");
  if (dq_acc != nullptr) {
    const int64_t n = B * T * (d / 8);
    const int64_t blocks = std::min<int64_t>(ceil_div(n, 256), (int64_t)mar_sm_count() * 16);
    attn_dq_convert_kernel<<<(unsigned)blocks, 256, 0, st>>>(dq_acc, (bf16*)dqkv, B * T, (int)d);
    MAR_LAUNCH_CHECK("attn_dq_convert");
  }
  return MAR_OK;
}

}  // namespace

int64_t attention_bwd_tc_work_floats(int64_t B, int64_t T, int64_t H, int64_t dh) {
  return ((B * H * T + 3) & ~(int64_t)3) + (T > BT ? B * T * H * dh : 0);
}

int attention_bwd_tc(const void* qkv, const uint8_t* key_mask, const void* out, const void* dout, const float* lse,
                     float* work, void* dqkv, int64_t B, int64_t T, int64_t H, int64_t dh, float p, const uint64_t* rng,
                     uint32_t site, cudaStream_t st) {
  MAR_CHECK_ARG(((uintptr_t)qkv % 16 == 0) && ((uintptr_t)out % 16 == 0) && ((uintptr_t)dout % 16 == 0) &&
                    ((uintptr_t)dqkv % 16 == 0) && ((uintptr_t)work % 16 == 0), "attention: pointers must be 16 B aligned");
  switch (dh) {
    case 64: return bwd_launch<64>(qkv, key_mask, out, dout, lse, work, dqkv, B, T, H, p, rng, site, st);
    case 96: return bwd_launch<96>(qkv, key_mask, out, dout, lse, work, dqkv, B, T, H, p, rng, site, st);
    case 128: return bwd_launch<128>(qkv, key_mask, out, dout, lse, work, dqkv, B, T, H, p, rng, site, st);
  }
  MAR_UNSUPPORTED("attention backward (tcgen05 engine): head dim %lld", (long long)dh);
}
