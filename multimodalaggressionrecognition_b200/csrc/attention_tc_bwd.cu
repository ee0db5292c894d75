// tcgen05 / TMEM / TMA attention over the packed in-projection output (B,T,3d): backward, ONE kernel.
//
//   P = exp(S - LSE),  P̃ = dropout(P),  dV = P̃ᵀ·dO,  dP = dropoutᵀ(dO·Vᵀ),  dS = P ⊙ (dP - delta),
//   dQ = dS·K / √dh,   dK = dSᵀ·Q / √dh                          delta = rowsum(dO ⊙ O)
//
// One CTA owns 128 KEYS of one (b,h) and walks over the query tiles (64 rows each); everything is computed on
// the TRANSPOSED score tile so that the probability operands sit in TMEM with keys as lanes:
//   Sᵀ  = K·Qᵀ       tcgen05.mma  M=128 keys, N=64 queries, K=dh   A = K tile, B = Q tile (both K-major smem)
//   dPᵀ = V·dOᵀ      tcgen05.mma  same shape                       A = V tile, B = dO tile
//   dV += P̃ᵀ·dO      tcgen05.mma  M=128 keys, N=dh, K=64 queries   A = P̃ᵀ bf16 in TMEM,  B = dO tile (MN-major)
//   dK += dSᵀ·Q      tcgen05.mma  same shape                       A = dSᵀ bf16 in TMEM, B = Q tile  (MN-major)
//   dQᵀ = Kᵀ·dSᵀ     tcgen05.mma  M=128 (dh, padded), N=64 queries, K=keys
//                                                                  A = K tile (MN-major), B = dSᵀ bf16 in smem (MN-major)
// dK / dV accumulate in TMEM over the whole walk.  dQ is produced TRANSPOSED so that it costs 64 TMEM columns
// instead of dh, which leaves room for a DOUBLE-BUFFERED Sᵀ/dPᵀ (dh <= 96): the MMAs of tile i+1 run under
// the softmax-gradient arithmetic of tile i, and the accumulating MMAs + dQ read-out of tile i under tile i+1.
// dQᵀ tiles are reduced across the key tiles of a (b,h) with coalesced red.global.add.f32 (lane <-> dh
// column) into an fp32 (B,T,d) workspace (converted to bf16 by attn_dq_convert_kernel), or written straight
// to dqkv when T <= 128 (one key tile).  5 contractions per tile pair, nothing recomputed.
//
// 22 warps: warp 0 (one elected lane) = TMA producer + MMA issuer; warp 1 = TMEM allocator + LSE/delta stager;
// warps 2-17 = 16 compute warps, thread <-> key row (TMEM lane), each warp owns one 16-query column chunk of one
// TMEM lane quarter; warps 18-21 = dQᵀ read-out (one per lane quarter).  Q / dO tiles ride a 3-stage TMA ring
// (3-D maps, rows >= T zero-filled).
//
// TMEM columns: dK [0,dh) | dV [dh,2dh) | dQᵀ [2dh,2dh+64) | stage s: Sᵀ 64 (P̃ᵀ packed bf16 in its first 32)
//               + dPᵀ 64 (dSᵀ packed in its first 32).   dh=96: 192+64+2·128 = 512.   dh=128: one stage.
// Dropout: the keep bits the forward call drew (common.cuh DropBits), staged per query tile by the stager warp —
// a thread's key is one bit position of its lane quarter's word, so the test is one AND per score.
// Fully masked rows (LSE = -inf) contribute nothing.
#include <algorithm>
#include <type_traits>
#define MAR_PDL_CLASS 8
#include "common.cuh"
#include "ptx_sm100.cuh"
#include "attention.cuh"
#include "attention_tc.cuh"

using namespace sm100;
using namespace attn_tc;

namespace {

constexpr int BT = 128;                 // keys per CTA
constexpr int BQ = 64;                  // queries per step
constexpr int NCOMPUTE = 16;            // compute warps
constexpr int NREAD = 4;                // dQᵀ read-out warps
constexpr int NTHREADS = (2 + NCOMPUTE + NREAD) * 32;
constexpr int QSTAGES = 3;              // Q / dO smem ring (a 4th stage bought nothing: the TMA round trip is not what the walk waits for)
constexpr int LSTAGES = 4;              // LSE / delta smem ring
constexpr int QBOX_BYTES = BQ * 128;    // one 64-row x 64-column box

struct BwdParams {
  const uint8_t* key_mask;
  const float* lse;
  const float* delta;
  float* dq_acc;      // (B,T,d) fp32, zero-initialised; nullptr when T <= 128 (dQ goes straight to dqkv)
  bf16* dqkv;         // (B,T,3d)
  float* dbias;       // (3d) fp32 or nullptr: += column sums of dqkv over all tokens (the in-projection's bias gradient)
  int B, T, H;
  float p_drop;
  const uint32_t* dbits;     // dropout keep bits of the forward call (common.cuh DropBits), nullptr when p_drop == 0
  int smem_bytes;
};

// smem tile -> global tensor with element-wise fp32 add, performed by the TMA unit (no LSU atomics)
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* map, const void* smem_src, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// DROP: compile-time; the keep-bit logic is not compiled into the p = 0 kernel.  BIAS: the column sums of dqkv (the
// in-projection's bias gradient) are taken here — dK / dV in the epilogue, dQ in the read-out warps when it goes straight
// to dqkv (T <= 128), else in the convert kernel; without BIAS none of that code exists (its registers cost the walk
// 1-13 % even when the pointer was null)
// BIAS: 0 none, 1 dK / dV sums (dQ's are taken by the convert kernel: T > 128), 2 dK / dV / dQ sums (T <= 128)
template <int DH, bool DROP, int BIAS>
__global__ void __launch_bounds__(NTHREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_kv, const __grid_constant__ CUtensorMap tm_q,
                   const __grid_constant__ CUtensorMap tm_do, const __grid_constant__ CUtensorMap tm_dq, const BwdParams p) {
  pdl_entry();
  constexpr int NBOX = (DH + 63) / 64;
  constexpr int KV_BYTES = NBOX * BOX_BYTES;        // 128-row operand tile
  constexpr int Q_BYTES = NBOX * QBOX_BYTES;        // 64-row operand tile
  constexpr int KSTEPS = DH / 16;
  constexpr int DS_BYTES = BT * 128;                // dSᵀ bf16 [128 keys][64 queries], 128 B swizzle
  constexpr int NST = (2 * DH + 64 + 2 * 128 <= 512) ? 2 : 1;     // Sᵀ/dPᵀ TMEM stages
  constexpr int LOOK = NST - 1;
  constexpr uint32_t COL_DK = 0, COL_DV = DH, COL_DQ = 2 * DH, COL_ST = 2 * DH + 64;   // stage s at COL_ST + 128 s
  constexpr uint32_t TMEM_COLS = 512;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + KV_BYTES;
  uint8_t* sQ = sV + KV_BYTES;                  // [QSTAGES][Q_BYTES]
  uint8_t* sDO = sQ + QSTAGES * Q_BYTES;        // [QSTAGES][Q_BYTES]
  uint8_t* sDS = sDO + QSTAGES * Q_BYTES;       // [2][DS_BYTES]: tile i+1 is written while the dQᵀ MMA of tile i still reads
  constexpr int DQ_ROWS = DH <= 96 ? 32 : 16;        // query rows per TMA reduce (staging size is what is left of 227 KB)
  float* sDQ = reinterpret_cast<float*>(sDS + 2 * DS_BYTES);   // [DQ_ROWS][DH] fp32 staging of a dQ block
  float* sL = sDQ + DQ_ROWS * DH;   // [LSTAGES][64] LSE in log2 units (+inf = no contribution)
  float* sD = sL + LSTAGES * BQ;                              // [LSTAGES][64] delta
  uint32_t* sB = reinterpret_cast<uint32_t*>(sD + LSTAGES * BQ);   // [LSTAGES][4 lane quarters][64 queries] dropout keep words
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + LSTAGES * 4 * BQ);
  uint64_t* bar_kv = bars + 0;
  uint64_t* bar_qdo = bars + 1;                 // [QSTAGES] Q_i / dO_i tiles landed
  uint64_t* bar_ld = bars + 5;                  // [LSTAGES] LSE / delta of tile i staged
  uint64_t* bar_ldfree = bars + 9;              // [LSTAGES] compute warps are done with that stage
  uint64_t* bar_s = bars + 13;                  // [2] Sᵀ, dPᵀ of a TMEM stage complete
  uint64_t* bar_pds = bars + 15;                // [2] P̃ᵀ, dSᵀ written (TMEM + smem) by the compute warps
  uint64_t* bar_mma2 = bars + 17;               // dV, dK, dQᵀ MMAs of tile i complete
  uint64_t* bar_dqr = bars + 18;                // dQᵀ tile read out of TMEM
  uint64_t* bar_done = bars + 19;               // every MMA of this CTA complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  if (reinterpret_cast<uint8_t*>(tmem_slot + 2) > smem_raw + p.smem_bytes) __trap();   // dynamic smem base less aligned than assumed

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = p.T, H = p.H;
  const int n_kt = (T + BT - 1) / BT;     // key tiles
  const int n_q = (T + BQ - 1) / BQ;      // query tiles
  const int bh = blockIdx.x / n_kt, kt = blockIdx.x % n_kt;
  const int b = bh / H, h = bh % H;
  const int d = H * DH;
  const int nk = min(BT, T - kt * BT);    // keys of this tile inside the sequence
  // The key tiles of one (b,h) run on neighbouring SMs at the same time and all reduce into the same dQ rows:
  // each starts its walk over the query tiles at a different place, so concurrent red.adds hit different rows.
  const int q_rot = (kt * n_q) / n_kt;
  auto q_tile = [&](int i) { const int t = i + q_rot; return t >= n_q ? t - n_q : t; };

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tm_kv);
    prefetch_tensormap(&tm_q);
    prefetch_tensormap(&tm_do);
    mbar_init(bar_kv, 1);
    for (int i = 0; i < QSTAGES; i++) mbar_init(bar_qdo + i, 1);
    for (int i = 0; i < LSTAGES; i++) { mbar_init(bar_ld + i, 1); mbar_init(bar_ldfree + i, NCOMPUTE); }
    for (int i = 0; i < 2; i++) { mbar_init(bar_s + i, 1); mbar_init(bar_pds + i, NCOMPUTE); }
    mbar_init(bar_mma2, 1);
    mbar_init(bar_dqr, NREAD);
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer + MMA issuer
    // The whole warp walks this code converged and waits on the mbarriers together; one elected lane issues the
    // TMA / tcgen05 instructions, so every operand is warp-uniform (no per-thread operand marshalling loops).
    auto load_qdo = [&](int i) {
      const int s = i % QSTAGES;
      mbar_expect_tx(bar_qdo + s, 2 * Q_BYTES);
#pragma unroll
      for (int bx = 0; bx < NBOX; bx++) {
        tma_load_3d(sQ + s * Q_BYTES + bx * QBOX_BYTES, &tm_q, bar_qdo + s, h * DH + bx * 64, q_tile(i) * BQ, b);
        tma_load_3d(sDO + s * Q_BYTES + bx * QBOX_BYTES, &tm_do, bar_qdo + s, h * DH + bx * 64, q_tile(i) * BQ, b);
      }
    };
    if (elect_one()) {
      mbar_expect_tx(bar_kv, 2 * KV_BYTES);
#pragma unroll
      for (int bx = 0; bx < NBOX; bx++) {
        tma_load_3d(sK + bx * BOX_BYTES, &tm_kv, bar_kv, d + h * DH + bx * 64, kt * BT, b);
        tma_load_3d(sV + bx * BOX_BYTES, &tm_kv, bar_kv, 2 * d + h * DH + bx * 64, kt * BT, b);
      }
      for (int i = 0; i < QSTAGES && i < n_q; i++) load_qdo(i);
    }
    __syncwarp();

    constexpr uint32_t idesc_acc = make_idesc_bf16(BT, DH, 0, 1);   // dV, dK: A in TMEM, B MN-major
    constexpr uint32_t idesc_dq = make_idesc_bf16(128, BQ, 1, 1);   // dQᵀ: A = K tile MN-major, B = dSᵀ MN-major
    const uint32_t sK_a = smem_u32(sK), sV_a = smem_u32(sV), sQ_a = smem_u32(sQ), sDO_a = smem_u32(sDO), sDS_a = smem_u32(sDS);
    const int ksteps_keys = (nk + 15) / 16;
    mbar_wait(bar_kv, 0);

    int issued = 0;
    auto issue_s = [&](int i) {           // Sᵀ_i and dPᵀ_i into TMEM stage i % NST
      const int s = i % QSTAGES;
      const int nq = min(BQ, T - q_tile(i) * BQ);
      const uint32_t col = tmem_base + COL_ST + (uint32_t)(i % NST) * 128;
      const uint32_t idesc_s = make_idesc_bf16(BT, (nq + 15) & ~15, 0, 0);
      mbar_wait(bar_qdo + s, (i / QSTAGES) & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ks++)
          umma_f16(col, make_desc_kmajor(sK_a + (ks / 4) * BOX_BYTES, ks % 4),
                   make_desc_kmajor(sQ_a + s * Q_BYTES + (ks / 4) * QBOX_BYTES, ks % 4), idesc_s, ks > 0 ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ks++)
          umma_f16(col + 64, make_desc_kmajor(sV_a + (ks / 4) * BOX_BYTES, ks % 4),
                   make_desc_kmajor(sDO_a + s * Q_BYTES + (ks / 4) * QBOX_BYTES, ks % 4), idesc_s, ks > 0 ? 1u : 0u);
        umma_commit(bar_s + (i % NST));
      }
      __syncwarp();
    };
    auto ensure_s = [&](int upto) {
      while (issued <= upto && issued < n_q) { issue_s(issued); issued++; }
    };
    ensure_s(LOOK);
    for (int i = 0; i < n_q; i++) {
      const int s = i % QSTAGES;
      const int nq = min(BQ, T - q_tile(i) * BQ);
      const int qsteps = (nq + 15) / 16;
      const uint32_t col = tmem_base + COL_ST + (uint32_t)(i % NST) * 128;
      mbar_wait(bar_pds + (i % NST), (i / NST) & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll 1
        for (int ks = 0; ks < qsteps; ks++)
          umma_f16_ts(tmem_base + COL_DV, col + ks * 8, make_desc_mnmajor(sDO_a + s * Q_BYTES, ks, QBOX_BYTES), idesc_acc,
                      (i > 0 || ks > 0) ? 1u : 0u);
#pragma unroll 1
        for (int ks = 0; ks < qsteps; ks++)
          umma_f16_ts(tmem_base + COL_DK, col + 64 + ks * 8, make_desc_mnmajor(sQ_a + s * Q_BYTES, ks, QBOX_BYTES), idesc_acc,
                      (i > 0 || ks > 0) ? 1u : 0u);
      }
      __syncwarp();
      if (i > 0) {                            // dQᵀ_{i-1} has left its TMEM columns
        mbar_wait(bar_dqr, (i - 1) & 1);
        tc_fence_after();
      }
      if (elect_one()) {
#pragma unroll 1
        for (int ks = 0; ks < ksteps_keys; ks++)
          umma_f16(tmem_base + COL_DQ, make_desc_mnmajor(sK_a, ks, BOX_BYTES),
                   make_desc_mnmajor(sDS_a + (i & 1) * DS_BYTES, ks, DS_BYTES), idesc_dq, ks > 0 ? 1u : 0u);
        umma_commit(bar_mma2);
        if (i == n_q - 1) umma_commit(bar_done);
      }
      __syncwarp();
      ensure_s(i + 1 + LOOK);                 // its TMEM stage was freed by the (in-order) dV/dK MMAs just issued
      if (i + QSTAGES < n_q) {
        mbar_wait(bar_mma2, i & 1);           // Q_i / dO_i smem stage is free again
        if (elect_one()) load_qdo(i + QSTAGES);
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- LSE / delta / keep-bit stager
    const int64_t bits_w = drop_words_per_row(T);
    for (int i = 0; i < n_q; i++) {
      const int s = i % LSTAGES;
      if (i >= LSTAGES) mbar_wait(bar_ldfree + s, ((i / LSTAGES) - 1) & 1);
      for (int k = lane; k < BQ; k += 32) {
        const int q = q_tile(i) * BQ + k;
        float l = INFINITY, dl = 0.f;
        if (q < T) {
          l = p.lse[(int64_t)bh * T + q] * LOG2E;
          if (l == -INFINITY) l = INFINITY;        // fully masked row: P = 0
          dl = p.delta[(int64_t)bh * T + q];
        }
        sL[s * BQ + k] = l;
        sD[s * BQ + k] = dl;
        if (DROP) {     // the 4 keep words of (query q, this CTA's 128 keys), one per TMEM lane quarter
          const uint4 w = q < T ? __ldg(reinterpret_cast<const uint4*>(p.dbits + ((int64_t)bh * T + q) * bits_w) + kt)
                                : make_uint4(0u, 0u, 0u, 0u);
          uint32_t* dst = sB + s * 4 * BQ + k;
          dst[0] = w.x; dst[BQ] = w.y; dst[2 * BQ] = w.z; dst[3 * BQ] = w.w;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_ld + s);
    }
  } else if (warp < 2 + NCOMPUTE) {
    // ---------------------------------------------------------------- compute warps: thread <-> key row
    const int quarter = warp & 3;                       // TMEM lanes this warp may access
    const int chunk = (warp - 2) >> 2;                  // 16-query column chunk
    const int row = quarter * 32 + lane;
    const int key = kt * BT + row;
    const bool kvalid = key < T && !(p.key_mask != nullptr && p.key_mask[(int64_t)b * T + key] != 0);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const float scale = rsqrtf((float)DH), scale2 = scale * LOG2E;
    constexpr bool drop = DROP;
    const float dscale = drop ? 256.f / (float)drop_keep_m(p.p_drop) : 1.f;     // 1 / keep probability (common.cuh DropBits)
    const uint32_t lanebit = 1u << lane;                 // this key's bit in the keep word of its lane quarter
    // dSᵀ smem row of this key: [row][64 queries], 16 B chunks XOR-swizzled by (row % 8)
    const int ds_off0 = row * 128 + (((chunk * 2) ^ (row & 7)) << 4);
    const int ds_off1 = row * 128 + (((chunk * 2 + 1) ^ (row & 7)) << 4);
    const int d3 = 3 * d;

    for (int i = 0; i < n_q; i++) {
      const int ls = i % LSTAGES, st = i % NST;
      const int nq = min(BQ, T - q_tile(i) * BQ);
      const bool active = chunk * 16 < nq;
      const float* l_s = sL + ls * BQ + chunk * 16;
      const float* d_s = sD + ls * BQ + chunk * 16;
      const uint32_t col = lane_addr + COL_ST + (uint32_t)st * 128;
      uint32_t pk[8], dsk[8];
      mbar_wait(bar_ld + ls, (i / LSTAGES) & 1);
      mbar_wait(bar_s + st, (i / NST) & 1);
      tc_fence_after();
      if (active) {
        uint32_t rs[16], rp[16];
        tmem_ld_32x32b_x16(col + chunk * 16, rs);
        tmem_ld_32x32b_x16(col + 64 + chunk * 16, rp);
        const uint32_t b_addr = smem_u32(sB + (ls * 4 + quarter) * BQ + chunk * 16);
        const uint32_t l_addr = smem_u32(l_s), d_addr = smem_u32(d_s);
        // the whole 16-column chunk inside the sequence and every key of the warp valid: no per-element selects
        const bool fast = (chunk * 16 + 16 <= nq) && __all_sync(0xffffffffu, kvalid);
        tmem_ld_wait();
        auto body = [&](auto fast_tag) {
          constexpr bool FAST = decltype(fast_tag)::value;
#pragma unroll
          for (int e4 = 0; e4 < 16; e4 += 4) {
            const float4 l4 = lds128(l_addr + e4 * 4), d4 = lds128(d_addr + e4 * 4);
            const float lv[4] = {l4.x, l4.y, l4.z, l4.w}, dv[4] = {d4.x, d4.y, d4.z, d4.w};
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (drop) b4 = lds128(b_addr + e4 * 4);
            const uint32_t bv[4] = {__float_as_uint(b4.x), __float_as_uint(b4.y), __float_as_uint(b4.z), __float_as_uint(b4.w)};
            float pd[4], ds[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
              const int ql = e4 + u;
              const bool ok = FAST || (kvalid && (chunk * 16 + ql < nq));
              float pr = ex2f(fmaf(__uint_as_float(rs[ql]), scale2, -lv[u]));
              if (!FAST) pr = ok ? pr : 0.f;
              float dp = __uint_as_float(rp[ql]);
              float pdv = pr;
              if (drop) {               // P̃ goes unscaled into dV (scaled in the epilogue); dP = keep·dP̃ / keep-probability
                const bool keep = (bv[u] & lanebit) != 0u;
                pdv = keep ? pr : 0.f;
                dp = keep ? dp : 0.f;
              }
              pd[u] = pdv;
              ds[u] = pr * fmaf(dp, dscale, -dv[u]);
              if (!FAST) ds[u] = ok ? ds[u] : 0.f;
            }
            pk[e4 / 2] = pack_bf16x2(pd[0], pd[1]);
            pk[e4 / 2 + 1] = pack_bf16x2(pd[2], pd[3]);
            dsk[e4 / 2] = pack_bf16x2(ds[0], ds[1]);
            dsk[e4 / 2 + 1] = pack_bf16x2(ds[2], ds[3]);
          }
        };
        if (fast) body(std::true_type{}); else body(std::false_type{});
      }
      // every warp of this lane quarter has read its Sᵀ / dPᵀ columns: the packed results may overwrite them
      named_bar_sync(1 + quarter, 128);
      if (active) {
        tmem_st_32x32b_x8(col + chunk * 8, pk);
        tmem_st_32x32b_x8(col + 64 + chunk * 8, dsk);
        const uint32_t dsb = smem_u32(sDS + (i & 1) * DS_BYTES);
        sts128(dsb + ds_off0, dsk[0], dsk[1], dsk[2], dsk[3]);
        sts128(dsb + ds_off1, dsk[4], dsk[5], dsk[6], dsk[7]);
        tmem_st_wait();
        fence_proxy_async();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(bar_pds + st); mbar_arrive(bar_ldfree + ls); }
    }

    // epilogue: dK (scaled) and dV of this key row -> dqkv
    mbar_wait(bar_done, 0);
    tc_fence_after();
    if (chunk * 32 < DH) {
      bf16* dst = p.dqkv + ((int64_t)b * T + key) * d3 + h * DH + chunk * 32;
#pragma unroll
      for (int which = 0; which < 2; which++) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(lane_addr + (which == 0 ? COL_DK : COL_DV) + chunk * 32, r);   // warp-collective
        tmem_ld_wait();
        const float sc = which == 0 ? scale : dscale;
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
        if (BIAS) {
          // column sums over this warp's 32 key rows, in place in r: butterfly transpose-reduce, lane l ends with column l
#pragma unroll
          for (int j = 0; j < 32; j++) r[j] = key < T ? __float_as_uint(__uint_as_float(r[j]) * sc) : 0u;
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) {
            const bool upper = (lane & off) != 0;
#pragma unroll
            for (int j = 0; j < off; j++) {
              const uint32_t send = upper ? r[j] : r[j + off];
              const uint32_t keep = upper ? r[j + off] : r[j];
              r[j] = __float_as_uint(__uint_as_float(keep) + __uint_as_float(__shfl_xor_sync(0xffffffffu, send, off)));
            }
          }
          if (chunk * 32 + lane < DH) atomicAdd(p.dbias + (which == 0 ? d : 2 * d) + h * DH + chunk * 32 + lane, __uint_as_float(r[0]));
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- dQᵀ read-out: lane <-> dh column
    const int quarter = warp & 3;
    const int c = quarter * 32 + lane;                  // column of this head's dQ
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const float scale = rsqrtf((float)DH);
    const bool live = quarter * 32 < DH;                // warp-uniform
    float qsum = 0.f;                                   // this key tile's share of Σ_q dQ[q, c] (the in-projection's bias gradient)
    for (int i = 0; i < n_q; i++) {
      const int nq = min(BQ, T - q_tile(i) * BQ);
      uint32_t r0[32], r1[32];
      mbar_wait(bar_mma2, i & 1);
      tc_fence_after();
      if (live) {
        tmem_ld_32x32b_x32(lane_addr + COL_DQ, r0);
        tmem_ld_32x32b_x32(lane_addr + COL_DQ + 32, r1);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_dqr);      // the MMA issuer may overwrite dQᵀ while we drain the registers
      if (BIAS == 2 && live) {          // T <= 128 only: otherwise the convert kernel sums dQ's columns
        if (nq == BQ) {                           // warp-uniform: only the last query tile of a sequence is ragged
#pragma unroll
          for (int j = 0; j < 32; j++) qsum += __uint_as_float(r0[j]) + __uint_as_float(r1[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; j++) {
            if (j < nq) qsum += __uint_as_float(r0[j]);
            if (32 + j < nq) qsum += __uint_as_float(r1[j]);
          }
        }
      }
      if (p.dq_acc != nullptr) {
        // dQ block -> fp32 staging [q][dh] in shared memory -> ONE TMA reduce-add per DQ_ROWS query rows (the L2 does the
        // adds; scalar red.global.add ran at about one element per clock per SM and paced the whole kernel)
#pragma unroll
        for (int g0 = 0; g0 < BQ; g0 += DQ_ROWS) {
          if (g0 < nq) {                                  // warp-uniform
            if (warp == 2 + NCOMPUTE && lane == 0) tma_store_wait_read();      // the previous reduce has read the staging
            named_bar_sync(5, NREAD * 32);
            if (live && c < DH) {
#pragma unroll
              for (int j = 0; j < DQ_ROWS; j++) {
                const int jj = g0 + j;
                sDQ[j * DH + c] = __uint_as_float(jj < 32 ? r0[jj & 31] : r1[jj & 31]) * scale;
              }
            }
            fence_proxy_async();
            named_bar_sync(5, NREAD * 32);
            if (warp == 2 + NCOMPUTE && lane == 0) {
              tma_reduce_add_3d(&tm_dq, sDQ, h * DH, q_tile(i) * BQ + g0, b);
              tma_store_commit();
            }
          }
        }
      } else if (live && c < DH) {
        bf16* dst = p.dqkv + ((int64_t)b * T + (int64_t)q_tile(i) * BQ) * (3 * d) + h * DH + c;
#pragma unroll
        for (int j = 0; j < 32; j++)
          if (j < nq) dst[(int64_t)j * 3 * d] = __float2bfloat16_rn(__uint_as_float(r0[j]) * scale);
#pragma unroll
        for (int j = 0; j < 32; j++)
          if (32 + j < nq) dst[(int64_t)(32 + j) * 3 * d] = __float2bfloat16_rn(__uint_as_float(r1[j]) * scale);
      }
    }
    if (BIAS == 2 && live && c < DH) atomicAdd(p.dbias + h * DH + c, qsum * scale);
    if (warp == 2 + NCOMPUTE && lane == 0) tma_store_wait_all();      // every reduce has landed before the CTA retires
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// delta (B,H,T) = rowsum over dh of dO ⊙ O.  One block per 4 token rows: thread <-> 8 columns, then H threads
// per row sum the dh/8 partials of their head.
constexpr int DELTA_ROWS = 8;
__global__ void __launch_bounds__(256)
attn_delta_tc_kernel(const bf16* __restrict__ out, const bf16* __restrict__ dout, float* __restrict__ delta, int64_t BT_rows,
                     int T, int H, int dh) {
  pdl_entry();
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

// dqkv[:, :, 0:d] (bf16) = dq_acc (fp32); dbias != nullptr: dbias[0:d] += column sums (the Q part of the in-projection's
// bias gradient).  Block = (d/8 column groups) x (RL row lanes): a thread keeps ITS 8 columns over all rows of the block.
__global__ void __launch_bounds__(1024)
attn_dq_convert_kernel(const float* __restrict__ acc, bf16* __restrict__ dqkv, int64_t rows, int d, int rows_per_block,
                       float* __restrict__ dbias) {
  pdl_entry();
  extern __shared__ float csum_s[];                 // [RL][d], only with dbias
  const int c = threadIdx.x, rl = threadIdx.y, RL = blockDim.y;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = min(rows, r0 + rows_per_block);
  float cs[8];
#pragma unroll
  for (int j = 0; j < 8; j++) cs[j] = 0.f;
  for (int64_t r = r0 + rl; r < r1; r += RL) {
    float v[8];
    Vec8<float>::load(acc + r * d + c * 8, v);
    Vec8<bf16>::store(dqkv + r * 3 * d + c * 8, v);
#pragma unroll
    for (int j = 0; j < 8; j++) cs[j] += v[j];
  }
  if (dbias == nullptr) return;
#pragma unroll
  for (int j = 0; j < 8; j++) csum_s[rl * d + c * 8 + j] = cs[j];
  __syncthreads();
  for (int col = rl * blockDim.x + c; col < d; col += RL * blockDim.x) {
    float t = 0.f;
    for (int k = 0; k < RL; k++) t += csum_s[k * d + col];
    atomicAdd(dbias + col, t);
  }
}

template <int DH, bool DROP, int BIAS>
int bwd_launch_t(const void* qkv, const uint8_t* key_mask, const void* out, const void* dout, const float* lse, float* work,
               void* dqkv, float* dbias, int64_t B, int64_t T, int64_t H, float p, const uint32_t* dbits, cudaStream_t st) {
  constexpr int NBOX = (DH + 63) / 64;
  constexpr int USED = 2 * NBOX * BOX_BYTES + 2 * QSTAGES * NBOX * QBOX_BYTES + 2 * BT * 128 + (DH <= 96 ? 32 : 16) * DH * 4 +
                       2 * LSTAGES * BQ * 4 + LSTAGES * 4 * BQ * 4 + 20 * 8 + 16;
  static_assert(USED + 1024 <= 232448, "attention backward: shared memory budget");
  constexpr int SMEM = USED + 1024;
  static_assert(SMEM <= 232448, "attention backward: shared memory budget");
  static bool cfg = false;
  if (!cfg) {
    MAR_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<DH, DROP, BIAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    cfg = true;
  }
  const int64_t d = H * DH;
  float* delta = work;
  float* dq_acc = T > BT ? work + ((B * H * T + 3) & ~(int64_t)3) : nullptr;
  {
    const int64_t rows = B * T;
    mar_launch(attn_delta_tc_kernel, (unsigned)ceil_div(rows, DELTA_ROWS), 256, DELTA_ROWS * (d / 8) * sizeof(float), st, 
        (const bf16*)out, (const bf16*)dout, delta, rows, (int)T, (int)H, DH);
    MAR_LAUNCH_CHECK("attn_delta_tc");
  }
  if (dq_acc != nullptr) MAR_CUDA(cudaMemsetAsync(dq_acc, 0, (size_t)(B * T * d) * sizeof(float), st));
  CUtensorMap tm_kv, tm_q, tm_do, tm_dq;
  int rc = make_map_btc(&tm_kv, qkv, B, T, 3 * d, BT);
  if (rc) return rc;
  rc = make_map_btc(&tm_q, qkv, B, T, 3 * d, BQ);
  if (rc) return rc;
  rc = make_map_btc(&tm_do, dout, B, T, d, BQ);
  if (rc) return rc;
  if (dq_acc != nullptr) {
    rc = make_map_f32_btc(&tm_dq, dq_acc, B, T, d, DH, DH <= 96 ? 32 : 16);
    if (rc) return rc;
  } else {
    tm_dq = tm_do;            // unused by the kernel when T <= 128
  }
  BwdParams prm;
  prm.key_mask = key_mask; prm.lse = lse; prm.delta = delta; prm.dq_acc = dq_acc; prm.dqkv = (bf16*)dqkv; prm.dbias = dbias;
  prm.B = (int)B; prm.T = (int)T; prm.H = (int)H; prm.p_drop = p; prm.dbits = dbits; prm.smem_bytes = SMEM;
  const int64_t n_t = ceil_div(T, BT);
  mar_launch(attn_bwd_tc_kernel<DH, DROP, BIAS>, (unsigned)(B * H * n_t), NTHREADS, SMEM, st, tm_kv, tm_q, tm_do, tm_dq, prm);
  MAR_LAUNCH_CHECK("attn_bwd_tc");
  if (dq_acc != nullptr) {
    const int c8n = (int)(d / 8);
    const int RL = std::max(1, std::min(8, 512 / c8n));
    const int64_t rows = B * T;
    // ~4 blocks per SM; every block a whole number of row-lane rounds
    int64_t rpb = ceil_div(rows, (int64_t)mar_sm_count() * 4);
    rpb = ceil_div(rpb, RL) * RL;
    const size_t smem = BIAS ? (size_t)RL * d * sizeof(float) : 0;
    mar_launch(attn_dq_convert_kernel, dim3((unsigned)ceil_div(rows, rpb)), dim3((unsigned)c8n, (unsigned)RL), smem, st, dq_acc,
               (bf16*)dqkv, rows, (int)d, (int)rpb, BIAS ? dbias : (float*)nullptr);
    MAR_LAUNCH_CHECK("attn_dq_convert");
  }
  return MAR_OK;
}

template <int DH>
int bwd_launch(const void* qkv, const uint8_t* key_mask, const void* out, const void* dout, const float* lse, float* work,
               void* dqkv, float* dbias, int64_t B, int64_t T, int64_t H, float p, const uint32_t* dbits, cudaStream_t st) {
  if (dbias != nullptr && T > BT) {
    if (p <= 0.f) return bwd_launch_t<DH, false, 1>(qkv, key_mask, out, dout, lse, work, dqkv, dbias, B, T, H, p, dbits, st);
    return bwd_launch_t<DH, true, 1>(qkv, key_mask, out, dout, lse, work, dqkv, dbias, B, T, H, p, dbits, st);
  }
  if (dbias != nullptr) {
    if (p <= 0.f) return bwd_launch_t<DH, false, 2>(qkv, key_mask, out, dout, lse, work, dqkv, dbias, B, T, H, p, dbits, st);
    return bwd_launch_t<DH, true, 2>(qkv, key_mask, out, dout, lse, work, dqkv, dbias, B, T, H, p, dbits, st);
  }
  if (p <= 0.f) return bwd_launch_t<DH, false, 0>(qkv, key_mask, out, dout, lse, work, dqkv, dbias, B, T, H, p, dbits, st);
  return bwd_launch_t<DH, true, 0>(qkv, key_mask, out, dout, lse, work, dqkv, dbias, B, T, H, p, dbits, st);
}

}  // namespace

int64_t attention_bwd_tc_work_floats(int64_t B, int64_t T, int64_t H, int64_t dh) {
  return ((B * H * T + 3) & ~(int64_t)3) + (T > BT ? B * T * H * dh : 0);
}

int attention_bwd_tc(const void* qkv, const uint8_t* key_mask, const void* out, const void* dout, const float* lse,
                     float* work, void* dqkv, float* dbias, int64_t B, int64_t T, int64_t H, int64_t dh, float p,
                     const uint32_t* dbits, cudaStream_t st) {
  MAR_CHECK_ARG(((uintptr_t)qkv % 16 == 0) && ((uintptr_t)out % 16 == 0) && ((uintptr_t)dout % 16 == 0) &&
                    ((uintptr_t)dqkv % 16 == 0) && ((uintptr_t)work % 16 == 0) && ((uintptr_t)dbits % 16 == 0),
                "attention: pointers must be 16 B aligned");
  MAR_CHECK_ARG(p == 0.f || dbits, "attention: dropout needs the keep-bit buffer of the forward call");
  switch (dh) {
    case 64: return bwd_launch<64>(qkv, key_mask, out, dout, lse, work, dqkv, dbias, B, T, H, p, dbits, st);
    case 96: return bwd_launch<96>(qkv, key_mask, out, dout, lse, work, dqkv, dbias, B, T, H, p, dbits, st);
    case 128: return bwd_launch<128>(qkv, key_mask, out, dout, lse, work, dqkv, dbias, B, T, H, p, dbits, st);
  }
  MAR_UNSUPPORTED("attention backward (tcgen05 engine): head dim %lld", (long long)dh);
}
