// tcgen05 / TMEM / TMA attention forward for sequences longer than one query tile (T > 128): two query tiles per CTA,
// two softmax warp-groups ping-ponging against the tensor pipe.
//
//   O = softmax(Q·Kᵀ/√dh + key mask)·V   per (batch b, head h), flash-style (scores never reach HBM).
//
// Why a second forward kernel (round 2, measured on B200): the one-tile kernel (attention_tc.cu: one S tile per CTA,
// 4 softmax warps, two passes over S in TMEM) spends most of its time in dependent chains — S(j+1) cannot be
// issued before P̃(j)·V has read the columns it overwrites, the softmax cannot start before S(j+1) lands, every
// softmax thread reads its row from TMEM twice in 32-column pieces with a wait after each — and ran at 4 100 clocks
// per 128x128 tile against ~1 700 of ALU work and 770 of MMA work; removing the dropout hash from its loop did not
// make it any faster.  cuDNN's fused attention was 1.45-1.9x faster on the same shapes.  Here:
//   * one CTA owns a PAIR of 128-row query tiles of one (b,h) (one CTA per SM, 384 threads).  K / V tiles are loaded
//     ONCE for both (2-stage TMA ring), S0 / S1 / O0 / O1 live side by side in TMEM (128+128+dh+dh columns);
//   * warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator + key-validity words;
//     warps 4-7 = softmax of query tile 0, warps 8-11 = softmax of query tile 1 (thread <-> query row = TMEM lane).
//     The issue order  S0(j) S1(j) | P̃0(j)·V  S0(j+1) | P̃1(j)·V  S1(j+1) | ...  keeps one group's softmax running
//     while the tensor pipe works for the other;
//   * a softmax thread reads its WHOLE 128-score row once (4 tcgen05.ld in flight, one wait, 128 registers;
//     setmaxnreg moves the registers of the three helper warps to the softmax warps), takes the row max with
//     3-input FMNMX, and goes straight to exp2 / row sum / keep-bit select / bf16 pack from the same registers;
//   * lazy rescaling (the reference max is only raised when the tile max exceeds it by 2^8) as before.
// Dropout keep bits, masks, LSE and the fully-masked-row rule are those of attention_tc.cu.
//
// Roofline: tensor pipe (4·T²·dh FLOP per (b,h)); bounded in practice by the softmax's exp2 (MUFU, 16/clk/SM)
// and ALU work — DESIGN.md §4.2.
#define MAR_PDL_CLASS 8
#include "common.cuh"
#include "ptx_sm100.cuh"
#include "attention.cuh"
#include "attention_tc.cuh"

using namespace sm100;
using namespace attn_tc;

namespace {

constexpr int BQ = 128;                 // query rows per tile (two tiles per CTA)
constexpr int BKV = 128;                // keys per tile
constexpr int MAX_KV_TILES = 128;       // T <= 16384
constexpr int KV_STAGES = 2;
constexpr int NTHREADS2 = 384;
constexpr float RESCALE_THRESHOLD = 8.f;   // log2 units

struct Fwd2Params {
  const uint8_t* key_mask;
  bf16* out;
  float* lse;
  int B, T, H;
  float p_drop;
  const uint32_t* dbits;
};

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float m;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(m) : "f"(a), "f"(b), "f"(c));
  return m;
}
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

template <int DH, bool DROP>
__global__ void __launch_bounds__(NTHREADS2, 1)
attn_fwd_tc2_kernel(const __grid_constant__ CUtensorMap tm_qkv, const Fwd2Params p) {
  pdl_entry();
  constexpr int NBOX = (DH + 63) / 64;
  constexpr int OP_BYTES = NBOX * BOX_BYTES;
  constexpr int KSTEPS = DH / 16;
  constexpr uint32_t TMEM_COLS = 512;
  constexpr uint32_t COL_S0 = 0, COL_O0 = 256;           // S_w at COL_S0 + 128 w (P̃_w packed into its first 64), O_w at COL_O0 + DH w

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // [2][OP_BYTES]
  uint8_t* sK = sQ + 2 * OP_BYTES;                      // [KV_STAGES][OP_BYTES]
  uint8_t* sV = sK + KV_STAGES * OP_BYTES;              // [KV_STAGES][OP_BYTES]
  uint32_t* sValid = reinterpret_cast<uint32_t*>(sV + KV_STAGES * OP_BYTES);    // [MAX_KV_TILES][4]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sValid + MAX_KV_TILES * 4);
  uint64_t* bar_q = bars + 0;        // [2] Q_w landed
  uint64_t* bar_k = bars + 2;        // [2] K stage full
  uint64_t* bar_v = bars + 4;        // [2] V stage full
  uint64_t* bar_kfree = bars + 6;    // [2] every S MMA that reads the K stage is complete
  uint64_t* bar_vfree = bars + 8;    // [2] every P̃·V MMA that reads the V stage is complete
  uint64_t* bar_s = bars + 10;       // [2] S_w(j) complete in TMEM
  uint64_t* bar_p = bars + 12;       // [2] P̃_w(j) written by the 4 softmax warps of group w
  uint64_t* bar_pv = bars + 14;      // [2] O_w += P̃_w(j)·V(j) complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = p.T, H = p.H;
  const int q_tiles = (T + BQ - 1) / BQ;
  const int q_pairs = (q_tiles + 1) / 2;
  const int n_kv = (T + BKV - 1) / BKV;
  const int bh = blockIdx.x / q_pairs, qp = blockIdx.x % q_pairs;
  const int b = bh / H, h = bh % H;
  const int d = H * DH;
  const bool has1 = 2 * qp + 1 < q_tiles;               // the pair's second query tile exists (CTA-uniform)

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tm_qkv);
    for (int i = 0; i < 2; i++) {
      mbar_init(bar_q + i, 1); mbar_init(bar_k + i, 1); mbar_init(bar_v + i, 1);
      mbar_init(bar_kfree + i, 1); mbar_init(bar_vfree + i, 1);
      mbar_init(bar_s + i, 1); mbar_init(bar_p + i, 4); mbar_init(bar_pv + i, 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
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

  if (warp < 4) {
    reg_dec<40>();
    if (warp == 0) {
      // ---------------------------------------------------------------- TMA producer (converged warp, one elected lane issues)
      auto load_tile = [&](uint8_t* dst, uint64_t* bar, int col0, int row0) {
        if (elect_one()) {
          mbar_expect_tx(bar, OP_BYTES);
#pragma unroll
          for (int bx = 0; bx < NBOX; bx++) tma_load_3d(dst + bx * BOX_BYTES, &tm_qkv, bar, col0 + bx * 64, row0, b);
        }
        __syncwarp();
      };
      load_tile(sQ, bar_q + 0, h * DH, (2 * qp) * BQ);
      load_tile(sK, bar_k + 0, d + h * DH, 0);
      if (has1) load_tile(sQ + OP_BYTES, bar_q + 1, h * DH, (2 * qp + 1) * BQ);
      load_tile(sV, bar_v + 0, 2 * d + h * DH, 0);
      if (n_kv > 1) {
        load_tile(sK + OP_BYTES, bar_k + 1, d + h * DH, BKV);
        load_tile(sV + OP_BYTES, bar_v + 1, 2 * d + h * DH, BKV);
      }
      for (int j = 2; j < n_kv; j++) {
        const int s = j & 1;
        const uint32_t par = ((j - 2) >> 1) & 1;
        mbar_wait(bar_kfree + s, par);
        load_tile(sK + s * OP_BYTES, bar_k + s, d + h * DH, j * BKV);
        mbar_wait(bar_vfree + s, par);
        load_tile(sV + s * OP_BYTES, bar_v + s, 2 * d + h * DH, j * BKV);
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------------- MMA issuer (converged warp, one elected lane issues)
      const uint32_t sQ_a = smem_u32(sQ), sK_a = smem_u32(sK), sV_a = smem_u32(sV);
      constexpr uint32_t idesc_pv = make_idesc_bf16(BQ, DH, 0, 1);
      auto issue_s = [&](int w, int j) {                  // S_w(j) = Q_w · K(j)ᵀ
        const int nk = min(BKV, T - j * BKV);
        const uint32_t idesc = make_idesc_bf16(BQ, (nk + 15) & ~15, 0, 0);
        const uint32_t qa = sQ_a + w * OP_BYTES, ka = sK_a + (j & 1) * OP_BYTES;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < KSTEPS; ks++)
            umma_f16(tmem_base + COL_S0 + w * 128, make_desc_kmajor(qa + (ks / 4) * BOX_BYTES, ks % 4),
                     make_desc_kmajor(ka + (ks / 4) * BOX_BYTES, ks % 4), idesc, ks > 0 ? 1u : 0u);
          umma_commit(bar_s + w);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int w, int j) {                 // O_w += P̃_w(j) · V(j)
        const int nk = min(BKV, T - j * BKV);
        const int pv_steps = (nk + 15) / 16;
        const uint32_t va = sV_a + (j & 1) * OP_BYTES;
        if (elect_one()) {
#pragma unroll 1
          for (int ks = 0; ks < pv_steps; ks++)
            umma_f16_ts(tmem_base + COL_O0 + w * DH, tmem_base + COL_S0 + w * 128 + ks * 8, make_desc_mnmajor(va, ks, BOX_BYTES),
                        idesc_pv, (j > 0 || ks > 0) ? 1u : 0u);
          umma_commit(bar_pv + w);
        }
        __syncwarp();
      };
      auto commit = [&](uint64_t* bar) {
        if (elect_one()) umma_commit(bar);
        __syncwarp();
      };
      mbar_wait(bar_q + 0, 0);
      mbar_wait(bar_k + 0, 0);
      tc_fence_after();
      issue_s(0, 0);
      if (has1) {
        mbar_wait(bar_q + 1, 0);
        tc_fence_after();
        issue_s(1, 0);
      }
      commit(bar_kfree + 0);
      for (int j = 0; j < n_kv; j++) {
        const int s = j & 1;
        const bool more = j + 1 < n_kv;
        mbar_wait(bar_v + s, (j >> 1) & 1);
        mbar_wait(bar_p + 0, j & 1);
        tc_fence_after();
        issue_pv(0, j);
        if (more) {
          mbar_wait(bar_k + (s ^ 1), ((j + 1) >> 1) & 1);
          tc_fence_after();
          issue_s(0, j + 1);                // in-order tensor pipe: runs after P̃0(j)·V, whose operand columns it overwrites
        }
        if (has1) {
          mbar_wait(bar_p + 1, j & 1);
          tc_fence_after();
          issue_pv(1, j);
        }
        commit(bar_vfree + s);
        if (more) {
          if (has1) issue_s(1, j + 1);
          commit(bar_kfree + (s ^ 1));
        }
      }
    }
  } else {
    reg_inc<232>();
    // ---------------------------------------------------------------- softmax: warp-group w <-> query tile w, thread <-> query row
    const int w = (warp - 4) >> 2;
    if (w == 0 || has1) {
      const int quarter = warp & 3;                        // TMEM lanes this warp may access
      const int row = quarter * 32 + lane;
      const int q = (2 * qp + w) * BQ + row;
      const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
      const uint32_t col_s = lane_addr + COL_S0 + w * 128, col_o = lane_addr + COL_O0 + w * DH;
      const float scale2 = rsqrtf((float)DH) * LOG2E;
      const DropBits db = make_drop_bits(p.dbits, T, DROP ? p.p_drop : 0.f);
      const uint4* brow = reinterpret_cast<const uint4*>(p.dbits + ((int64_t)bh * T + q) * db.W);
      float m_ref = -INFINITY, l_run = 0.f;

      for (int j = 0; j < n_kv; j++) {
        const uint32_t ph = j & 1;
        const int nk = min(BKV, T - j * BKV);
        const int nchunk = (nk + 31) / 32;
        const uint32_t* vw = sValid + j * 4;
        const bool full = (vw[0] & vw[1] & vw[2] & vw[3]) == 0xffffffffu;
        uint4 bw = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
        if (DROP) bw = __ldg(brow + j);                    // keep bits of this row's 128 keys (rows >= T read the padding)
        mbar_wait(bar_s + w, ph);
        tc_fence_after();

        // the whole score row, once
        uint32_t r0[32], r1[32], r2[32], r3[32];
        tmem_ld_32x32b_x32(col_s, r0);
        if (nchunk > 1) tmem_ld_32x32b_x32(col_s + 32, r1);
        if (nchunk > 2) tmem_ld_32x32b_x32(col_s + 64, r2);
        if (nchunk > 3) tmem_ld_32x32b_x32(col_s + 96, r3);
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
          if (j > 0) {                                     // O_w holds the previous tiles' sum once P̃_w(j-1)·V is complete
            mbar_wait(bar_pv + w, (j - 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < DH / 32; c++) {
              uint32_t o[32];
              tmem_ld_32x32b_x32(col_o + c * 32, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; i++) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
              tmem_st_32x32b_x32(col_o + c * 32, o);
            }
            tmem_st_wait();
          }
        }
        const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;

        // P = exp2(s*scale2 - m) (exp2(-inf) = 0 for masked keys), row sum, keep-bit select, bf16 pack -> TMEM
        auto chunk_p = [&](uint32_t (&r)[32], int c, uint32_t dw) {
          if (c >= nchunk) return;
          float l0 = 0.f, l1 = 0.f;
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; i++) {
            const float e0 = ex2f(fmaf(__uint_as_float(r[2 * i]), scale2, -m_use));
            const float e1 = ex2f(fmaf(__uint_as_float(r[2 * i + 1]), scale2, -m_use));
            l0 += e0; l1 += e1;
            if (DROP) pk[i] = pack_bf16x2(((dw >> (2 * i)) & 1u) ? e0 : 0.f, ((dw >> (2 * i + 1)) & 1u) ? e1 : 0.f);
            else pk[i] = pack_bf16x2(e0, e1);
          }
          l_run += l0 + l1;
          tmem_st_32x32b_x16(col_s + c * 16, pk);
        };
        chunk_p(r0, 0, bw.x); chunk_p(r1, 1, bw.y); chunk_p(r2, 2, bw.z); chunk_p(r3, 3, bw.w);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p + w);
      }

      // epilogue: O / l (times the dropout scale) -> bf16 -> global ; LSE
      mbar_wait(bar_pv + w, (n_kv - 1) & 1);
      tc_fence_after();
      const float inv = l_run > 0.f ? db.scale / l_run : 0.f;
      bf16* orow = p.out + ((int64_t)b * T + q) * d + h * DH;
#pragma unroll
      for (int c = 0; c < DH / 32; c++) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(col_o + c * 32, r);
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
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int DH>
int fwd2_launch(const void* qkv, const uint8_t* key_mask, void* out, float* lse, int64_t B, int64_t T, int64_t H, float p,
                const uint32_t* dbits, cudaStream_t st) {
  constexpr int NBOX = (DH + 63) / 64;
  constexpr int SMEM = (2 + 2 * KV_STAGES) * NBOX * BOX_BYTES + MAX_KV_TILES * 16 + 16 * 8 + 16 + 1024;
  static_assert(SMEM <= 232448, "attention forward (pair kernel): shared memory budget");
  static bool cfg = false;
  if (!cfg) {
    MAR_CUDA(cudaFuncSetAttribute(attn_fwd_tc2_kernel<DH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    MAR_CUDA(cudaFuncSetAttribute(attn_fwd_tc2_kernel<DH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    cfg = true;
  }
  CUtensorMap tm;
  int rc = make_map_btc(&tm, qkv, B, T, 3 * H * DH, BQ);
  if (rc) return rc;
  Fwd2Params prm;
  prm.key_mask = key_mask; prm.out = (bf16*)out; prm.lse = lse; prm.B = (int)B; prm.T = (int)T; prm.H = (int)H;
  prm.p_drop = p; prm.dbits = dbits;
  const int64_t q_pairs = (ceil_div(T, BQ) + 1) / 2;
  const unsigned grid = (unsigned)(B * H * q_pairs);
  if (p > 0.f) mar_launch(attn_fwd_tc2_kernel<DH, true>, grid, NTHREADS2, SMEM, st, tm, prm);
  else mar_launch(attn_fwd_tc2_kernel<DH, false>, grid, NTHREADS2, SMEM, st, tm, prm);
  MAR_LAUNCH_CHECK("attn_fwd_tc2");
  return MAR_OK;
}

}  // namespace

int attention_fwd_tc2(const void* qkv, const uint8_t* key_mask, void* out, float* lse, int64_t B, int64_t T, int64_t H,
                      int64_t dh, float p, const uint32_t* dbits, cudaStream_t st) {
  MAR_CHECK_ARG(((uintptr_t)qkv % 16 == 0) && ((uintptr_t)out % 16 == 0) && ((uintptr_t)dbits % 16 == 0),
                "attention: pointers must be 16 B aligned");
  MAR_CHECK_ARG(p == 0.f || dbits, "attention: dropout needs the keep-bit buffer");
  switch (dh) {
    case 64: return fwd2_launch<64>(qkv, key_mask, out, lse, B, T, H, p, dbits, st);
    case 96: return fwd2_launch<96>(qkv, key_mask, out, lse, B, T, H, p, dbits, st);
    case 128: return fwd2_launch<128>(qkv, key_mask, out, lse, B, T, H, p, dbits, st);
  }
  MAR_UNSUPPORTED("attention (tcgen05 pair kernel): head dim %lld", (long long)dh);
}
