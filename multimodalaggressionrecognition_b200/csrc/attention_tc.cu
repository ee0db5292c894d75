// tcgen05 / TMEM / TMA attention over the packed in-projection output (B,T,3d): forward.
//
//   O = softmax(Q·Kᵀ/√dh + key mask)·V   per (batch b, head h), flash-style (scores never reach HBM).
//
// One CTA owns 128 query rows of one (b,h); two CTAs are resident per SM (96 KB smem, 256 TMEM columns, 32 K registers
// each) so one CTA's softmax overlaps the other CTA's MMAs, loads, prologue and epilogue.  256 threads:
//   warp 0   TMA producer + MMA issuer (converged warp, one elected lane issues).  Q / K / V tiles are [128 rows x dh]
//            bf16, staged by TMA (3-D tensor map (col, t, b): rows t >= T are ZERO-filled by the hardware, so tiles
//            never read the next batch element) as 128 B-swizzled boxes of 64 columns.
//            S = Q·Kᵀ      : tcgen05.mma  M=128, N=keys (16..128), K=dh     A,B from smem (K-major)
//            O += P̃·V      : tcgen05.mma  M=128, N=dh,             K=keys   A = P̃ from TMEM, B = V from smem (MN-major)
//   warp 1   TMEM allocator; packs the key-padding mask into 128-bit validity words per key tile.  (warps 2-3 idle:
//            setmaxnreg works on aligned groups of 4 warps — the helper group gives its registers to the softmax group)
//   warps 4-7 softmax: thread r owns query row r (TMEM lane r) and reads its WHOLE score row once (4 tcgen05.ld in
//            flight, one wait, 128 registers), takes the row max with 3-input FMNMX and goes straight to exp2 / row
//            sum / keep-bit select / bf16 pack from the same registers, then tcgen05.st of P̃ into the columns S
//            occupied.  (Round 1 read S twice in 32-column pieces with a wait after each: a chain of dependent TMEM
//            round trips, 2 warps per scheduler, 46 % issue utilisation.)  The running max is only raised when the tile
//            max exceeds it by 2^8 (lazy rescaling), in which case the O accumulator row is rescaled in TMEM.
// Sequences of more than 512 keys go to the pair kernel (attention_tc2.cu: two query tiles per CTA share every K / V
// tile and their softmax warp-groups ping-pong against the tensor pipe).
//
// Dropout on P reads the call's keep bits (common.cuh DropBits: drawn once per call, shared by every engine and by
// the backward pass): one 16 B load per query row and key tile, a bit test per score, no hashing in the loop.
// Fully masked rows give O = 0 and LSE = -inf (torch 2.11 safe softmax).
//
// Roofline: tensor pipe (4·T²·dh FLOP per (b,h)); the softmax's exp2 (16/clk/SM) and its ALU work bound it below
// the MMA rate — see DESIGN.md §4.2.
#include <stdlib.h>
#define MAR_PDL_CLASS 8
#include "common.cuh"
#include "ptx_sm100.cuh"
#include "attention.cuh"
#include "attention_tc.cuh"

using namespace sm100;
using namespace attn_tc;

namespace {

constexpr int BQ = 128;                 // query rows per CTA
constexpr int BKV = 128;                // keys per tile
constexpr int MAX_KV_TILES = 128;       // T <= 16384
constexpr float RESCALE_THRESHOLD = 8.f;   // log2 units: P̃ <= 2^8 before the reference max is raised

struct FwdParams {
  const uint8_t* key_mask;
  bf16* out;
  float* lse;
  int B, T, H;
  float p_drop;
  const uint32_t* dbits;     // dropout keep bits (common.cuh DropBits), nullptr when p_drop == 0
};

// DROP: compile-time, the mask logic is not even compiled into the p = 0 kernel
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float m;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(m) : "f"(a), "f"(b), "f"(c));
  return m;
}

template <int DH, bool DROP>
__global__ void __launch_bounds__(256, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const FwdParams p) {
  pdl_entry();
  constexpr int NBOX = (DH + 63) / 64;
  constexpr int OP_BYTES = NBOX * BOX_BYTES;
  constexpr int KSTEPS = DH / 16;
  constexpr uint32_t TMEM_COLS = 256;
  constexpr uint32_t COL_S = 0, COL_P = 0, COL_O = 128;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + OP_BYTES;
  uint8_t* sV = smem + 2 * OP_BYTES;
  uint32_t* sValid = reinterpret_cast<uint32_t*>(smem + 3 * OP_BYTES);        // [MAX_KV_TILES][4]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sValid + MAX_KV_TILES * 4);
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_k = bars + 1;
  uint64_t* bar_v = bars + 2;
  uint64_t* bar_s = bars + 3;     // S tile complete in TMEM (and the K smem tile is free)
  uint64_t* bar_p = bars + 4;     // P̃ written to TMEM by the 4 softmax warps
  uint64_t* bar_pv = bars + 5;    // P̃·V complete (O updated, V smem tile free)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = p.T, H = p.H;
  const int q_tiles = (T + BQ - 1) / BQ;
  const int n_kv = (T + BKV - 1) / BKV;
  const int bh = blockIdx.x / q_tiles, qt = blockIdx.x % q_tiles;
  const int b = bh / H, h = bh % H;
  const int d = H * DH;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tm_qkv);
    mbar_init(bar_q, 1); mbar_init(bar_k, 1); mbar_init(bar_v, 1);
    mbar_init(bar_s, 1); mbar_init(bar_p, 4); mbar_init(bar_pv, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
    // validity words: bit i of word w of tile j <=> key j*128 + w*32 + i is attended to
    for (int w = 0; w < n_kv * 4; w++) {
      const int k = w * 32 + lane;
      const bool v = k < T && !(p.key_mask != nullptr && p.key_mask[(int64_t)b * T + k] != 0);
      const uint32_t word = __ballot_sync(0xffffffffu, v);
      if (lane == 0) sValid[w] = word;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  else asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer + MMA issuer
    // The whole warp walks this code converged and waits on the mbarriers together; one elected lane issues the TMA
    // and tcgen05 instructions, so every operand is warp-uniform (a one-lane branch makes ptxas wrap each MMA in a
    // per-lane operand-marshalling loop: ~20 instructions per issue, as long as a 128x128x16 MMA runs).
    auto load_tile = [&](uint8_t* dst, uint64_t* bar, int col0, int row0) {
      if (elect_one()) {
        mbar_expect_tx(bar, OP_BYTES);
#pragma unroll
        for (int bx = 0; bx < NBOX; bx++) tma_load_3d(dst + bx * BOX_BYTES, &tm_qkv, bar, col0 + bx * 64, row0, b);
      }
      __syncwarp();
    };
    const uint32_t sQ_a = smem_u32(sQ), sK_a = smem_u32(sK), sV_a = smem_u32(sV);
    auto issue_s = [&](int j) {
      const int nk = min(BKV, T - j * BKV);
      const uint32_t idesc = make_idesc_bf16(BQ, (nk + 15) & ~15, 0, 0);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ks++)
          umma_f16(tmem_base + COL_S, make_desc_kmajor(sQ_a + (ks / 4) * BOX_BYTES, ks % 4),
                   make_desc_kmajor(sK_a + (ks / 4) * BOX_BYTES, ks % 4), idesc, ks > 0 ? 1u : 0u);
        umma_commit(bar_s);
      }
      __syncwarp();
    };
    load_tile(sQ, bar_q, h * DH, qt * BQ);
    load_tile(sK, bar_k, d + h * DH, 0);
    load_tile(sV, bar_v, 2 * d + h * DH, 0);
    mbar_wait(bar_q, 0);
    mbar_wait(bar_k, 0);
    tc_fence_after();
    issue_s(0);
    constexpr uint32_t idesc_pv = make_idesc_bf16(BQ, DH, 0, 1);
    for (int j = 0; j < n_kv; j++) {
      const uint32_t ph = j & 1;
      if (j + 1 < n_kv) {                 // K tile is free once S_j is complete
        mbar_wait(bar_s, ph);
        load_tile(sK, bar_k, d + h * DH, (j + 1) * BKV);
      }
      mbar_wait(bar_p, ph);
      mbar_wait(bar_v, ph);
      tc_fence_after();
      const int nk = min(BKV, T - j * BKV);
      const int pv_steps = (nk + 15) / 16;
      if (elect_one()) {
#pragma unroll 1
        for (int ks = 0; ks < pv_steps; ks++)
          umma_f16_ts(tmem_base + COL_O, tmem_base + COL_P + ks * 8, make_desc_mnmajor(sV_a, ks, BOX_BYTES), idesc_pv,
                      (j > 0 || ks > 0) ? 1u : 0u);
        umma_commit(bar_pv);
      }
      __syncwarp();
      if (j + 1 < n_kv) {
        mbar_wait(bar_k, ph ^ 1);
        tc_fence_after();
        issue_s(j + 1);                   // executes after P̃·V_j (in-order tensor pipe): S may overwrite P̃_j
        mbar_wait(bar_pv, ph);            // V tile free
        load_tile(sV, bar_v, 2 * d + h * DH, (j + 1) * BKV);
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- softmax: thread <-> query row
    const int quarter = warp & 3;                        // TMEM lanes this warp may access
    const int row = quarter * 32 + lane;
    const int q = qt * BQ + row;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const float scale2 = rsqrtf((float)DH) * LOG2E;
    constexpr bool drop = DROP;
    // keep bits of this query row: 4 words per 128-key tile, one 16 B load per tile (rows >= T read the buffer's padding)
    const DropBits db = make_drop_bits(p.dbits, T, drop ? p.p_drop : 0.f);
    const uint4* brow = reinterpret_cast<const uint4*>(p.dbits + ((int64_t)bh * T + q) * db.W);
    float m_ref = -INFINITY, l_run = 0.f;

    for (int j = 0; j < n_kv; j++) {
      const uint32_t ph = j & 1;
      const int nk = min(BKV, T - j * BKV);
      const int nchunk = (nk + 31) / 32;
      const uint32_t* vw = sValid + j * 4;
      const bool full = (vw[0] & vw[1] & vw[2] & vw[3]) == 0xffffffffu;
      mbar_wait(bar_s, ph);
      tc_fence_after();

      // the whole score row, once
      uint32_t r0[32], r1[32], r2[32], r3[32];
      tmem_ld_32x32b_x32(lane_addr + COL_S, r0);
      if (nchunk > 1) tmem_ld_32x32b_x32(lane_addr + COL_S + 32, r1);
      if (nchunk > 2) tmem_ld_32x32b_x32(lane_addr + COL_S + 64, r2);
      if (nchunk > 3) tmem_ld_32x32b_x32(lane_addr + COL_S + 96, r3);
      tmem_ld_wait();

      float mx = -INFINITY;
      auto chunk_max = [&](uint32_t (&r)[32], int c) {
        if (c >= nchunk) return;
        if (!full) {
          const uint32_t word = vw[c];
#pragma unroll
          for (int i = 0; i < 32; i++) r[i] = ((word >> i) & 1u) ? r[i] : 0xff800000u;      // -inf
        }
        float a = -INFINITY, bq = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          a = fmax3(a, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
          bq = fmax3(bq, __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
        }
        mx = fmax3(mx, a, bq);
      };
      chunk_max(r0, 0); chunk_max(r1, 1); chunk_max(r2, 2); chunk_max(r3, 3);

      const float mt = mx * scale2;                      // -inf if every key of the tile is masked
      const bool raise = mt > m_ref + RESCALE_THRESHOLD; // first finite tile: m_ref = -inf -> true
      if (__any_sync(0xffffffffu, raise)) {
        float factor = 1.f;
        if (raise) {
          factor = (m_ref == -INFINITY) ? 0.f : ex2f(m_ref - mt);
          m_ref = mt;
          l_run *= factor;
        }
        if (j > 0) {                                     // O holds the previous tiles' sum (P̃·V_{j-1} is complete)
#pragma unroll
          for (int c = 0; c < DH / 32; c++) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(lane_addr + COL_O + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i++) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
            tmem_st_32x32b_x32(lane_addr + COL_O + c * 32, o);
          }
          tmem_st_wait();
        }
      }
      const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;

      // P = exp2(s*scale2 - m) (exp2(-inf) = 0 for masked keys), row sum, keep-bit select (the 1/(1-p) scale is applied
      // to O at the end), bf16 pack -> TMEM
      uint4 bw = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
      if (drop) bw = __ldg(brow + j);
      auto chunk_p = [&](uint32_t (&r)[32], int c, uint32_t dw) {
        if (c >= nchunk) return;
        float l0 = 0.f, l1 = 0.f;
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; i++) {
          const float e0 = ex2f(fmaf(__uint_as_float(r[2 * i]), scale2, -m_use));
          const float e1 = ex2f(fmaf(__uint_as_float(r[2 * i + 1]), scale2, -m_use));
          l0 += e0; l1 += e1;
          if (drop) pk[i] = pack_bf16x2(((dw >> (2 * i)) & 1u) ? e0 : 0.f, ((dw >> (2 * i + 1)) & 1u) ? e1 : 0.f);
          else pk[i] = pack_bf16x2(e0, e1);
        }
        l_run += l0 + l1;
        tmem_st_32x32b_x16(lane_addr + COL_P + c * 16, pk);
      };
      chunk_p(r0, 0, bw.x); chunk_p(r1, 1, bw.y); chunk_p(r2, 2, bw.z); chunk_p(r3, 3, bw.w);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
    }

    // epilogue: O / l -> bf16 -> global ; LSE
    mbar_wait(bar_pv, (n_kv - 1) & 1);
    tc_fence_after();
    const float inv = l_run > 0.f ? db.scale / l_run : 0.f;
    bf16* orow = p.out + ((int64_t)b * T + q) * d + h * DH;
#pragma unroll
    for (int c = 0; c < DH / 32; c++) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(lane_addr + COL_O + c * 32, r);
      tmem_ld_wait();
      if (q < T) {
#pragma unroll
        for (int g = 0; g < 4; g++) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(r[g * 8 + 0]) * inv, __uint_as_float(r[g * 8 + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(r[g * 8 + 2]) * inv, __uint_as_float(r[g * 8 + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(r[g * 8 + 4]) * inv, __uint_as_float(r[g * 8 + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(r[g * 8 + 6]) * inv, __uint_as_float(r[g * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + c * 32 + g * 8) = u;
        }
      }
    }
    if (q < T) p.lse[(int64_t)bh * T + q] = l_run > 0.f ? (m_ref + log2f(l_run)) * LN2 : -INFINITY;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int DH>
int fwd_launch(const void* qkv, const uint8_t* key_mask, void* out, float* lse, int64_t B, int64_t T, int64_t H, float p,
               const uint32_t* dbits, cudaStream_t st) {
  constexpr int NBOX = (DH + 63) / 64;
  constexpr int SMEM = 3 * NBOX * BOX_BYTES + MAX_KV_TILES * 16 + 64 + 1024;
  static bool cfg = false;
  if (!cfg) {
    MAR_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<DH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    MAR_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<DH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    cfg = true;
  }
  CUtensorMap tm;
  int rc = make_map_btc(&tm, qkv, B, T, 3 * H * DH, BQ);
  if (rc) return rc;
  FwdParams prm;
  prm.key_mask = key_mask; prm.out = (bf16*)out; prm.lse = lse; prm.B = (int)B; prm.T = (int)T; prm.H = (int)H;
  prm.p_drop = p; prm.dbits = dbits;
  const int64_t q_tiles = ceil_div(T, BQ);
  const unsigned grid = (unsigned)(B * H * q_tiles);
  if (p > 0.f) mar_launch(attn_fwd_tc_kernel<DH, true>, grid, 256, SMEM, st, tm, prm);
  else mar_launch(attn_fwd_tc_kernel<DH, false>, grid, 256, SMEM, st, tm, prm);
  MAR_LAUNCH_CHECK("attn_fwd_tc");
  return MAR_OK;
}

}  // namespace

bool attention_tc_supported(int64_t B, int64_t T, int64_t H, int64_t dh, int dtype) {
  if (dtype != MAR_BF16) return false;
  if (!(dh == 64 || dh == 96 || dh == 128)) return false;
  if (T < 1 || T > (int64_t)MAX_KV_TILES * BKV) return false;
  if (B * H * ceil_div(T, BQ) >= (1ll << 31)) return false;
  if ((3 * H * dh * 2) % 16 != 0) return false;
  return true;
}

int attention_fwd_tc(const void* qkv, const uint8_t* key_mask, void* out, float* lse, int64_t B, int64_t T, int64_t H,
                     int64_t dh, float p, const uint32_t* dbits, cudaStream_t st) {
  MAR_CHECK_ARG(((uintptr_t)qkv % 16 == 0) && ((uintptr_t)out % 16 == 0) && ((uintptr_t)dbits % 16 == 0),
                "attention: pointers must be 16 B aligned");
  MAR_CHECK_ARG(p == 0.f || dbits, "attention: dropout needs the keep-bit buffer");
  // more than one query tile: the pair kernel (two query tiles per CTA, ping-ponging softmax warp-groups);
  // MAR_ATTN_FWD1=1 keeps this one-tile kernel for A/B measurements
  static int fwd1 = -1;
  if (fwd1 < 0) { const char* e = getenv("MAR_ATTN_FWD1"); fwd1 = (e && e[0] == '1') ? 1 : 0; }
  if (T > 4 * BKV && !fwd1) return attention_fwd_tc2(qkv, key_mask, out, lse, B, T, H, dh, p, dbits, st);
  switch (dh) {
    case 64: return fwd_launch<64>(qkv, key_mask, out, lse, B, T, H, p, dbits, st);
    case 96: return fwd_launch<96>(qkv, key_mask, out, lse, B, T, H, p, dbits, st);
    case 128: return fwd_launch<128>(qkv, key_mask, out, lse, B, T, H, p, dbits, st);
  }
  MAR_UNSUPPORTED("attention (tcgen05 engine): head dim %lld", (long long)dh);
}
