// Persistent, warp-specialised bf16 GEMM for sm_100a: TMA (128 B swizzle, multicast across a 2-CTA
// cluster) -> smem ring -> tcgen05.mma (cta_group::1, 128 x BN x 16) with fp32 accumulators in TMEM
// (double buffered) -> tcgen05.ld epilogue with the fused linear epilogue (bias / ReLU / dropout /
// aux mask / residual) -> swizzled smem staging -> TMA store;  or fp32 split-K reduction (wgrad).
//
//   D[M,N] = epilogue( A · B ),  reduction length Kr
//   A: K-major  = row-major (M, Kr)         or MN-major = row-major (Kr, M)
//   B: K-major  = row-major (N, Kr)         or MN-major = row-major (Kr, N)
//
//   forward  y  = x·Wᵀ   : A = x (M,K) K-major,        B = W (N,K) K-major
//   dgrad    dx = dz·W   : A = dz (M,N) K-major,       B = Wᵀ (K,N) K-major (bf16 transposed copy)
//   wgrad    dW = dzᵀ·x  : A = dz (rows,N) MN-major,   B = x (rows,K) MN-major, fp32 out, split-K
//
// Warp roles (384 threads): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-11
// epilogue (warp w reads TMEM lanes 32*(w%4)..+31, one accumulator row per thread; warps 4-7 take
// the left half of the tile's columns, warps 8-11 the right half).
//
// Epilogue (bf16 out): measured with ncu, per-thread 16 B global stores of an accumulator row touch 32
// different lines per instruction (2x the ideal sector count) and held the K=768 GEMMs at 0.95 PFLOP/s
// while the bare main loop ran at 1.2-1.4.  So every epilogue warp owns a 4 KB [32 rows x 64 cols]
// 128 B-swizzled staging box: the residual / mask tile is TMA-LOADED into it, combined in place with the
// accumulator, and the result TMA-STORED (full 128 B lines, clipped at the matrix edge by the hardware).
//
// Roofline: tensor pipe.  Algorithmic FLOPs per launch = 2·M·N·Kr.
#include "common.cuh"
#include "ptx_sm100.cuh"
#include "gemm_tcgen05.cuh"

using namespace sm100;

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;            // 64 bf16 = 128 B = one swizzle atom
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;   // 16 KB
constexpr int ATOM_BYTES = BLOCK_K * 128;              // one MN-major box: 64 k-rows x 128 B = 8 KB
// Epilogue warps: EW = 8 (warp <-> 32 rows x BN/2 columns, 4 KB staging box of 64 columns, 128 B swizzle) or EW = 16
// (warp <-> 32 rows x BN/4 columns, 2 KB staging box of 32 columns, 64 B swizzle; the producer warp-group hands its
// registers to the four epilogue warp-groups with setmaxnreg).  ncu on the K = 768 shapes with the dropout / residual
// epilogue (profiles/r02b_gemm_out_full_ncu.txt): the epilogue, not the tensor pipe, paces the kernel — 790 instructions
// per 32 x 32 chunk, 2 epilogue warps per scheduler, 45 % of the issue slots used, the rest latency (tcgen05.ld, bias
// loads, staged-box arrival, store-read waits) that two warps cannot cover.  Four warps per scheduler can.
constexpr int EPI_BYTES = 32 * 1024;                   // staging boxes, all epilogue warps together
template <int EW> struct Epi {
  static constexpr int THREADS = 128 + EW * 32;
  static constexpr int BOX_BYTES = EPI_BYTES / EW;     // 4 KB [32 rows][64 bf16] or 2 KB [32 rows][32 bf16]
  static constexpr int BOX_COLS = BOX_BYTES / 64;
};
constexpr int MAX_EPI_WARPS = 16;

template <int BN> struct Cfg {
  static constexpr int B_STAGE_BYTES = BN * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int TMEM_COLS = 2 * BN;   // 512 or 256: power of two
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align slack*/ + 512 /*barriers*/;
};

struct TcParams {
  int M, N, Kr;
  int splits, kb_per_split;
  void* out;              // fp32 path only (bf16 goes through the TMA store map)
  int64_t ldo;
  const float* bias;
  int staged_mode;        // 0 none, 1 residual add, 2 aux mask (result *= staged > 0 ? aux_scale : 0)
  float aux_scale;
  int flags;
  float p_drop;
  const uint64_t* rng;
  uint32_t site;
  float* colsum;          // bf16 out: colsum[n] += Σ_m result[m,n] (fp32; the bias gradient of the layer that produced A's gradient)
  int atomic_out;         // fp32 out via red.add (split-K)
  int accumulate;         // fp32 out += (no split)
};

// 32 lanes x 16 consecutive 32-bit TMEM columns
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int BN, bool A_MN, bool B_MN, typename OutT, int CL, int EW>
__global__ void __launch_bounds__(Epi<EW>::THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                    const __grid_constant__ CUtensorMap tma_o, const __grid_constant__ CUtensorMap tma_s, const TcParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::STAGES * A_STAGE_BYTES;
  uint8_t* smem_e = smem + C::STAGES * C::STAGE_BYTES;                       // epilogue staging boxes (1024 B aligned)
  constexpr int EPI_WARPS = EW;
  constexpr int EPI_BOX_BYTES = Epi<EW>::BOX_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_e + EPI_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::STAGES;
  uint64_t* tmem_full = bars + 2 * C::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* epi_bar = tmem_empty + 2;                                        // [EPI_WARPS] staged-input arrival
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(epi_bar + MAX_EPI_WARPS);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int rank = CL > 1 ? (int)cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / CL, num_clusters = gridDim.x / CL;
  const int m_blocks = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int m_groups = (m_blocks + CL - 1) / CL;        // CL consecutive M-blocks share one B tile
  const int n_blocks = (p.N + BN - 1) / BN;
  const int num_tiles = m_groups * n_blocks;
  const int num_work = num_tiles * p.splits;
  const int kb_total = (p.Kr + BLOCK_K - 1) / BLOCK_K;
  constexpr uint16_t MC_MASK = (uint16_t)((1u << CL) - 1);

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tma_a);
    prefetch_tensormap(&tma_b);
    if (sizeof(OutT) == 2) { prefetch_tensormap(&tma_o); if (p.staged_mode) prefetch_tensormap(&tma_s); }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::STAGES; i++) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], CL); }
    for (int i = 0; i < 2; i++) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], EPI_WARPS); }
    for (int i = 0; i < EPI_WARPS; i++) mbar_init(&epi_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  if (CL > 1) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: this grid may have been started while the previous kernel of the stream was still
  // running (its CTAs take an SM as soon as one is free and run the prologue above); nothing of global memory has been
  // touched so far.  Let the NEXT kernel do the same, then wait for the previous one to complete and flush.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (EW == 16) {      // 20 warps x 96 registers at launch: the producer group keeps 56, each epilogue thread gets 104
    if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
  }

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int w = cluster_id; w < num_work; w += num_clusters) {
        const int tile = w % num_tiles, split = w / num_tiles;
        const int m_blk = (tile / n_blocks) * CL + rank, n_blk = tile % n_blocks;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; kb++) {
          mbar_wait(&empty_bar[stage], phase ^ 1);       // both CTAs of the cluster have drained this stage
          mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
          uint8_t* sa = smem_a + stage * A_STAGE_BYTES;
          uint8_t* sb = smem_b + stage * C::B_STAGE_BYTES;
          if (!A_MN) {
            tma_load_2d(sa, &tma_a, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
          } else {
#pragma unroll
            for (int i = 0; i < BLOCK_M / 64; i++)
              tma_load_2d(sa + i * ATOM_BYTES, &tma_a, &full_bar[stage], m_blk * BLOCK_M + i * 64, kb * BLOCK_K);
          }
          // B: this CTA fetches its 1/CL share of the tile and multicasts it to the whole cluster
          if (!B_MN) {
            constexpr int ROWS = BN / CL;
            uint8_t* dst = sb + rank * ROWS * 128;
            if (CL > 1) tma_load_2d_mc(dst, &tma_b, &full_bar[stage], kb * BLOCK_K, n_blk * BN + rank * ROWS, MC_MASK);
            else tma_load_2d(dst, &tma_b, &full_bar[stage], kb * BLOCK_K, n_blk * BN);
          } else {
            constexpr int ATOMS = BN / 64 / CL;
#pragma unroll
            for (int i = 0; i < ATOMS; i++) {
              const int a = rank * ATOMS + i;
              if (CL > 1) tma_load_2d_mc(sb + a * ATOM_BYTES, &tma_b, &full_bar[stage], n_blk * BN + a * 64, kb * BLOCK_K, MC_MASK);
              else tma_load_2d(sb + a * ATOM_BYTES, &tma_b, &full_bar[stage], n_blk * BN + a * 64, kb * BLOCK_K);
            }
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int w = cluster_id; w < num_work; w += num_clusters, it++) {
        const int split = w / num_tiles;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb_total, kb0 + p.kb_per_split);
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = kb0; kb < kb1; kb++) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem_a + stage * A_STAGE_BYTES);
          const uint32_t sb = smem_u32(smem_b + stage * C::B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; k++) {
            const uint64_t adesc = A_MN ? make_desc_mnmajor(sa, k, ATOM_BYTES) : make_desc_kmajor(sa, k);
            const uint64_t bdesc = B_MN ? make_desc_mnmajor(sb, k, ATOM_BYTES) : make_desc_kmajor(sb, k);
            umma_f16(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          // free the stage in EVERY CTA of the cluster (the peer multicasts into our smem too)
          if (CL > 1) umma_commit_mc(&empty_bar[stage], MC_MASK); else umma_commit(&empty_bar[stage]);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[as]);
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int ew = warp - 4;                // epilogue warp index
    const int half = ew >> 2;               // column half (EW = 8) / quarter (EW = 16) of the tile
    const bool do_drop = (p.flags & MAR_EPI_DROPOUT) && p.p_drop > 0.f;
    const bool idx32 = ((uint64_t)(p.M + BLOCK_M) * (uint64_t)p.N >> 1) < 0xffffffffull;   // incl. the rows of a ragged last tile
    DropKey dk;
    if (do_drop) dk = make_drop_key(p.rng, p.site, p.p_drop);
    uint8_t* box = smem_e + ew * EPI_BOX_BYTES;
    uint64_t* sbar = &epi_bar[ew];
    uint32_t sphase = 0;
    constexpr int CHUNKS = BN * 4 / (32 * EW);   // 32-column TMEM chunks handled by this warp
    // Bias of the tile's columns, lane <-> column (one coalesced 128 B load per 32-column chunk), fetched ONE TILE AHEAD
    // and handed out with shuffles: the per-chunk 16 B bias loads sat in every chunk's dependent chain (ncu: the FADDs
    // behind them were the largest long-scoreboard stall of the epilogue warps)
    float bias_nxt[CHUNKS];
    auto load_bias = [&](int w2) {
#pragma unroll
      for (int cc = 0; cc < CHUNKS; cc++) {
        const int col = ((w2 % num_tiles) % n_blocks) * BN + (half * CHUNKS + cc) * 32 + lane;
        bias_nxt[cc] = (p.bias != nullptr && w2 < num_work && w2 / num_tiles == 0 && col < p.N) ? __ldg(p.bias + col) : 0.f;
      }
    };
    load_bias(cluster_id);
    int it = 0;
    for (int w = cluster_id; w < num_work; w += num_clusters, it++) {
      const int tile = w % num_tiles, split = w / num_tiles;
      const int m_blk = (tile / n_blocks) * CL + rank, n_blk = tile % n_blocks;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int row0 = m_blk * BLOCK_M + q * 32;
      const int row = row0 + lane;
      float bias_cur[CHUNKS];
#pragma unroll
      for (int cc = 0; cc < CHUNKS; cc++) bias_cur[cc] = bias_nxt[cc];
      load_bias(w + num_clusters);
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      if constexpr (EW == 16 && sizeof(OutT) == 2) {
        // one 2 KB box per 32-column chunk: [32 rows][32 bf16], 16 B pieces XOR-swizzled by (row / 2) % 4 (TMA 64 B swizzle)
#pragma unroll 1
        for (int cc = 0; cc < CHUNKS; cc++) {
          const int c = half * CHUNKS + cc;
          const int col0 = n_blk * BN + c * 32;
          const bool active = col0 < p.N && row0 < p.M;        // warp-uniform
          if (active) {
            if (lane == 0) {
              tma_store_wait_read();                           // the previous store has finished reading the box
              if (p.staged_mode) {
                mbar_expect_tx(sbar, EPI_BOX_BYTES);
                tma_load_2d(box, &tma_s, sbar, col0, row0);
              }
            }
            __syncwarp();
          }
          // two 16-column halves: 16 accumulator registers live at a time (104 registers per epilogue thread)
          uint8_t* rowp = box + lane * 64;
          const int swz = (lane >> 1) & 3;
#pragma unroll 1
          for (int hh = 0; hh < 2; hh++) {
            uint32_t r[16];
            tmem_ld_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + c * 32 + hh * 16), r);
            tmem_ld_wait();
            if (cc == CHUNKS - 1 && hh == 1) {     // accumulator fully read: hand the TMEM stage back to the MMA warp now
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tmem_empty[as]);
            }
            if (!active) continue;
            const int colh = col0 + hh * 16;
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; j++) v[j] = __uint_as_float(r[j]);
            if (p.bias != nullptr) {
              const float bl = cc == 0 ? bias_cur[0] : bias_cur[CHUNKS - 1];
#pragma unroll
              for (int j = 0; j < 16; j++) v[j] += __shfl_sync(0xffffffffu, bl, hh * 16 + j);
            }
            if (p.flags & MAR_EPI_RELU_PRE) {
#pragma unroll
              for (int j = 0; j < 16; j++) v[j] = fmaxf(v[j], 0.f);
            }
            if (do_drop) {
              const uint64_t e0 = (uint64_t)row * (uint64_t)p.N + (uint64_t)colh;
              if (idx32) {
                const uint32_t p0 = (uint32_t)(e0 >> 1);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                  bool k0, k1;
                  drop_keep2_32(dk, p0 + (uint32_t)j, k0, k1);
                  v[2 * j] = k0 ? v[2 * j] * dk.scale : 0.f;
                  v[2 * j + 1] = k1 ? v[2 * j + 1] * dk.scale : 0.f;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 8; j++) {
                  bool k0, k1;
                  drop_keep2(dk, (e0 >> 1) + j, k0, k1);
                  v[2 * j] = k0 ? v[2 * j] * dk.scale : 0.f;
                  v[2 * j + 1] = k1 ? v[2 * j + 1] * dk.scale : 0.f;
                }
              }
            }
            if (p.flags & MAR_EPI_RELU_POST) {
#pragma unroll
              for (int j = 0; j < 16; j++) v[j] = fmaxf(v[j], 0.f);
            }
            if (p.staged_mode) {
              if (hh == 0) { mbar_wait(sbar, sphase); sphase ^= 1; }
#pragma unroll
              for (int g = 0; g < 2; g++) {
                const uint4 u = *reinterpret_cast<const uint4*>(rowp + (((hh * 2 + g) ^ swz) << 4));
                const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                  const float2 f = __bfloat1622float2(h2[j]);
                  if (p.staged_mode == 1) { v[g * 8 + 2 * j] += f.x; v[g * 8 + 2 * j + 1] += f.y; }
                  else {
                    v[g * 8 + 2 * j] = f.x > 0.f ? v[g * 8 + 2 * j] * p.aux_scale : 0.f;
                    v[g * 8 + 2 * j + 1] = f.y > 0.f ? v[g * 8 + 2 * j + 1] * p.aux_scale : 0.f;
                  }
                }
              }
            }
#pragma unroll
            for (int g = 0; g < 2; g++) {
              uint4 u;
              u.x = pack_bf16x2(v[g * 8 + 0], v[g * 8 + 1]);
              u.y = pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]);
              u.z = pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]);
              u.w = pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]);
              if (p.colsum != nullptr && row >= p.M) u = make_uint4(0u, 0u, 0u, 0u);   // rows past M are clipped by the store
              *reinterpret_cast<uint4*>(rowp + (((hh * 2 + g) ^ swz) << 4)) = u;
            }
          }
          if (!active) continue;
          if (p.colsum != nullptr) {
            // column sums of the ROUNDED block (what a pass over the written tensor would sum), read back lane <-> column:
            // 32 independent 4 B loads instead of the 31-shuffle butterfly
            __syncwarp();
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int rr = 0; rr < 32; rr++) {
              const uint32_t w = *reinterpret_cast<const uint32_t*>(box + rr * 64 + (((lane >> 3) ^ ((rr >> 1) & 3)) << 4) + ((lane & 7) >> 1) * 4);
              const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
              if (rr & 1) s1 += (lane & 1) ? f.y : f.x; else s0 += (lane & 1) ? f.y : f.x;
            }
            if (col0 + lane < p.N) atomicAdd(p.colsum + col0 + lane, s0 + s1);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) { tma_store_2d(&tma_o, box, col0, row0); tma_store_commit(); }
        }
        continue;
      }
#pragma unroll 1
      for (int cc = 0; cc < CHUNKS; cc++) {
        const int c = half * CHUNKS + cc;
        const int col0 = n_blk * BN + c * 32;
        const bool col_ok = col0 < p.N;                        // warp-uniform
        const bool box_start = (cc & 1) == 0 || CHUNKS == 1;
        const int bcol0 = (CHUNKS == 1) ? col0 : col0 - (cc & 1) * 32;   // first column of the 64-wide staging box
        if (sizeof(OutT) == 2 && box_start && col_ok && row0 < p.M) {
          // the box is free once the previous TMA store has finished reading it; then fetch the staged input
          if (lane == 0) {
            tma_store_wait_read();
            if (p.staged_mode) {
              mbar_expect_tx(sbar, EPI_BOX_BYTES);
              tma_load_2d(box, &tma_s, sbar, bcol0, row0);
            }
          }
          __syncwarp();
        }
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + c * 32), r);
        tmem_ld_wait();
        if (cc == CHUNKS - 1) {             // accumulator fully read: hand the TMEM stage back to the MMA warp now
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[as]);
        }
        if (!col_ok || row0 >= p.M) continue;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = __uint_as_float(r[j]);
        if (p.bias != nullptr) {
          float bl = bias_cur[0];
#pragma unroll
          for (int k = 1; k < CHUNKS; k++) bl = cc == k ? bias_cur[k] : bl;
#pragma unroll
          for (int j = 0; j < 32; j++) v[j] += __shfl_sync(0xffffffffu, bl, j);
        }
        if (p.flags & MAR_EPI_RELU_PRE) {
#pragma unroll
          for (int j = 0; j < 32; j++) v[j] = fmaxf(v[j], 0.f);
        }
        if (do_drop) {
          const uint64_t e0 = (uint64_t)row * (uint64_t)p.N + (uint64_t)col0;
          if (idx32) {                       // every pair index of this launch fits 32 bits (warp-uniform)
            const uint32_t p0 = (uint32_t)(e0 >> 1);
#pragma unroll
            for (int j = 0; j < 16; j++) {
              bool k0, k1;
              drop_keep2_32(dk, p0 + (uint32_t)j, k0, k1);
              v[2 * j] = k0 ? v[2 * j] * dk.scale : 0.f;
              v[2 * j + 1] = k1 ? v[2 * j + 1] * dk.scale : 0.f;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; j++) {
              bool k0, k1;
              drop_keep2(dk, (e0 >> 1) + j, k0, k1);
              v[2 * j] = k0 ? v[2 * j] * dk.scale : 0.f;
              v[2 * j + 1] = k1 ? v[2 * j + 1] * dk.scale : 0.f;
            }
          }
        }
        if (p.flags & MAR_EPI_RELU_POST) {
#pragma unroll
          for (int j = 0; j < 32; j++) v[j] = fmaxf(v[j], 0.f);
        }
        if (sizeof(OutT) == 2) {
          // this thread's row of the staging box: 8 x 16 B chunks, physical chunk = logical ^ (row & 7)
          uint8_t* rowp = box + lane * 128;
          const int ch0 = (CHUNKS == 1) ? 0 : (cc & 1) * 4;
          if (p.staged_mode) {
            if (box_start) { mbar_wait(sbar, sphase); sphase ^= 1; }
#pragma unroll
            for (int g = 0; g < 4; g++) {
              const uint4 u = *reinterpret_cast<const uint4*>(rowp + (((ch0 + g) ^ (lane & 7)) << 4));
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
              for (int j = 0; j < 4; j++) {
                const float2 f = __bfloat1622float2(h2[j]);
                if (p.staged_mode == 1) { v[g * 8 + 2 * j] += f.x; v[g * 8 + 2 * j + 1] += f.y; }
                else {
                  v[g * 8 + 2 * j] = f.x > 0.f ? v[g * 8 + 2 * j] * p.aux_scale : 0.f;
                  v[g * 8 + 2 * j + 1] = f.y > 0.f ? v[g * 8 + 2 * j + 1] * p.aux_scale : 0.f;
                }
              }
            }
          }
          if (p.colsum != nullptr) {
            // column sums of this warp's 32 x 32 block: butterfly transpose-reduce (31 shuffles), after which lane l
            // holds the sum of column l over the 32 rows; one 128 B red.add per block
            float s[32];
#pragma unroll
            for (int j = 0; j < 32; j++) s[j] = row < p.M ? v[j] : 0.f;
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
              const bool upper = (lane & off) != 0;
#pragma unroll
              for (int j = 0; j < off; j++) {
                const float send = upper ? s[j] : s[j + off];
                const float keep = upper ? s[j + off] : s[j];
                s[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
              }
            }
            if (col0 + lane < p.N) atomicAdd(p.colsum + col0 + lane, s[0]);
          }
#pragma unroll
          for (int g = 0; g < 4; g++) {
            uint4 u;
            u.x = pack_bf16x2(v[g * 8 + 0], v[g * 8 + 1]);
            u.y = pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]);
            u.z = pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]);
            u.w = pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]);
            *reinterpret_cast<uint4*>(rowp + (((ch0 + g) ^ (lane & 7)) << 4)) = u;
          }
          const bool box_end = (cc & 1) == 1 || CHUNKS == 1 || col0 + 32 >= p.N;
          if (box_end) {
            fence_proxy_async();            // generic-proxy smem writes -> visible to the TMA (async proxy)
            __syncwarp();
            if (lane == 0) { tma_store_2d(&tma_o, box, bcol0, row0); tma_store_commit(); }
          }
        } else {
          if (row >= p.M) continue;
          float* op = reinterpret_cast<float*>(p.out) + (int64_t)row * p.ldo + col0;
#pragma unroll
          for (int g = 0; g < 8; g++) {
            if (col0 + g * 4 < p.N) {
              if (p.atomic_out) {
                red_add_v4(op + g * 4, v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
              } else {
                float4 o = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
                if (p.accumulate) {
                  const float4 old = *reinterpret_cast<const float4*>(op + g * 4);
                  o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                }
                *reinterpret_cast<float4*>(op + g * 4) = o;
              }
            }
          }
        }
      }
    }
    if (sizeof(OutT) == 2 && lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  if (CL > 1) cluster_sync_all(); else __syncthreads();   // nobody leaves while a peer may still write into its smem
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 2-D bf16 tensor map over a row-major (rows, cols) matrix with leading dimension ld (elements);
// box = (box_cols = 64 elements = 128 B, box_rows), 128 B swizzle, zero fill out of bounds.
int make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols = 64) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) { mar_set_error("cuTensorMapEncodeTiled not available from the driver"); return MAR_ERR_CUDA; }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};   // 64 columns = 128 B rows / 32 columns = 64 B rows
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    mar_set_error("cuTensorMapEncodeTiled failed (%d): rows=%lld cols=%lld ld=%lld base=%p", (int)r, (long long)rows,
                  (long long)cols, (long long)ld, base);
    return MAR_ERR_CUDA;
  }
  return MAR_OK;
}

template <int BN, bool A_MN, bool B_MN, typename OutT, int CL, int EW>
int launch(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo, const CUtensorMap& ms, const TcParams& p,
           cudaStream_t st) {
  using C = Cfg<BN>;
  auto kern = gemm_tcgen05_kernel<BN, A_MN, B_MN, OutT, CL, EW>;
  static bool configured = false;
  if (!configured) {
    MAR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    configured = true;
  }
  const int m_groups = (int)ceil_div(ceil_div(p.M, BLOCK_M), CL), n_blocks = (p.N + BN - 1) / BN;
  const int64_t work = (int64_t)m_groups * n_blocks * p.splits;
  const int max_clusters = mar_sm_count() / CL;
  const int clusters = (int)(work < max_clusters ? work : max_clusters);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(clusters * CL), 1, 1);
  cfg.blockDim = dim3(Epi<EW>::THREADS, 1, 1);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;     // MAR_PDL bit 1 (common.cuh)
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = mar_pdl_enabled(1) ? 2 : 1;
  MAR_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mb, mo, ms, p));
  MAR_LAUNCH_CHECK("gemm_tcgen05");
  return MAR_OK;
}

template <int BN, bool A_MN, bool B_MN, typename OutT>
int launch_cl(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo, const CUtensorMap& ms, const TcParams& p,
              int cl, int ew, cudaStream_t st) {
  if constexpr (sizeof(OutT) == 2) {
    if (ew == 16)
      return cl == 2 ? launch<BN, A_MN, B_MN, OutT, 2, 16>(ma, mb, mo, ms, p, st) : launch<BN, A_MN, B_MN, OutT, 1, 16>(ma, mb, mo, ms, p, st);
  }
  return cl == 2 ? launch<BN, A_MN, B_MN, OutT, 2, 8>(ma, mb, mo, ms, p, st) : launch<BN, A_MN, B_MN, OutT, 1, 8>(ma, mb, mo, ms, p, st);
}

}  // namespace

bool gemm_tcgen05_supported(const TcGemmArgs& a) {
  if (a.M < 1 || a.N < 8 || a.Kr < 8) return false;
  if (a.N % 8 != 0) return false;
  // TMA: 16 B aligned bases and row pitches
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) % 16) == 0; };
  if (!al16(a.A) || !al16(a.B) || !al16(a.out)) return false;
  if ((a.lda * 2) % 16 != 0 || (a.ldb * 2) % 16 != 0) return false;
  if (a.residual != nullptr && (!al16(a.residual) || (a.ldr * 2) % 16 != 0)) return false;
  if (a.aux != nullptr && (!al16(a.aux) || (a.ldaux * 2) % 16 != 0)) return false;
  if (a.aux != nullptr && a.residual != nullptr) return false;   // one staged input per launch
  if ((a.aux != nullptr || a.residual != nullptr || a.colsum != nullptr) && a.out_fp32) return false;
  const int64_t osz = a.out_fp32 ? 4 : 2;
  if ((a.ldo * osz) % 16 != 0) return false;
  // TN (forward, dgrad on a transposed weight copy), NT-on-rows (wgrad, fp32 out) and dgrad on the weight itself
  // (A = dz K-major, B = W (N,K) row-major = MN-major over the reduction, bf16 out)
  if (a.a_mn_major && !a.b_mn_major) return false;
  if (a.a_mn_major && !a.out_fp32) return false;
  if (!a.a_mn_major && a.b_mn_major && a.out_fp32) return false;
  if (a.M >= (1ll << 31) || a.N >= (1ll << 31) || a.Kr >= (1ll << 31)) return false;
  return true;
}

int gemm_tcgen05(const TcGemmArgs& a, cudaStream_t st) {
  if (!gemm_tcgen05_supported(a)) MAR_UNSUPPORTED("gemm_tcgen05: unsupported problem M=%lld N=%lld K=%lld", (long long)a.M, (long long)a.N, (long long)a.Kr);
  mar_set_engine(MAR_ENGINE_TCGEN05);
  TcParams p;
  p.M = (int)a.M; p.N = (int)a.N; p.Kr = (int)a.Kr;
  p.out = a.out; p.ldo = a.ldo; p.bias = a.bias;
  p.staged_mode = a.residual != nullptr ? 1 : (a.aux != nullptr ? 2 : 0);
  p.aux_scale = a.aux_scale;
  p.flags = a.flags; p.p_drop = a.p_drop; p.rng = a.rng; p.site = a.site;
  p.atomic_out = 0; p.accumulate = a.accumulate;
  p.colsum = a.out_fp32 ? nullptr : a.colsum;
  const int BN = (a.N > 128) ? 256 : 128;
  const int64_t m_blocks = ceil_div(a.M, BLOCK_M);
  // cluster of 2 when there are at least two M-blocks to pair (MAR_TC_CLUSTER=1 disables, for A/B measurements)
  static int env_cl = -1;
  if (env_cl < 0) { const char* e = getenv("MAR_TC_CLUSTER"); env_cl = e ? atoi(e) : 2; }
  const int cl = (m_blocks >= 2 && env_cl >= 2) ? 2 : 1;
  const int kb_total = (int)ceil_div(a.Kr, BLOCK_K);
  p.splits = 1; p.kb_per_split = kb_total;
  if (a.allow_split && a.out_fp32 && a.flags == 0 && a.residual == nullptr && a.aux == nullptr) {
    const int64_t tiles = ceil_div(m_blocks, cl) * ceil_div(a.N, BN);
    const int64_t clusters = mar_sm_count() / cl;
    const int64_t max_split = kb_total / 8 > 0 ? kb_total / 8 : 1;   // at least 8 k-blocks (512 rows) per split
    // Work items = tiles x splits are dealt round-robin to the persistent clusters: pick the split count whose last
    // wave is fullest (tiles x splits just under a multiple of the cluster count), looking at up to 4 waves; the
    // smallest such count wins ties (every split adds one fp32 red.add pass over the output).
    int64_t splits = 1;
    double best = 0.0;
    for (int64_t waves = 1; waves <= 4; waves++) {
      int64_t sp = waves * clusters / tiles;
      if (sp < 1) sp = 1;
      if (sp > max_split) sp = max_split;
      const int64_t kbs = ceil_div(kb_total, sp);
      sp = ceil_div(kb_total, kbs);                              // splits actually produced by this k-block count
      const int64_t items = tiles * sp;
      const double util = (double)items / (double)(ceil_div(items, clusters) * clusters);
      if (util > best + 0.02) { best = util; splits = sp; }
    }
    if (splits > 1) {
      p.kb_per_split = (int)ceil_div(kb_total, splits);
      p.splits = (int)ceil_div(kb_total, p.kb_per_split);
      p.atomic_out = 1;
      if (!a.accumulate) {
        if (a.ldo == a.N) MAR_CUDA(cudaMemsetAsync(a.out, 0, (size_t)a.M * a.N * 4, st));
        else MAR_CUDA(cudaMemset2DAsync(a.out, (size_t)a.ldo * 4, 0, (size_t)a.N * 4, (size_t)a.M, st));
      }
    }
  }
  CUtensorMap ma, mb, mo, ms;
  int rc;
  if (!a.a_mn_major) {
    rc = make_map(&ma, a.A, a.M, a.Kr, a.lda, BLOCK_M); if (rc) return rc;
    if (a.b_mn_major) rc = make_map(&mb, a.B, a.Kr, a.N, a.ldb, BLOCK_K);
    else rc = make_map(&mb, a.B, a.N, a.Kr, a.ldb, BN / cl);
    if (rc) return rc;
  } else {
    rc = make_map(&ma, a.A, a.Kr, a.M, a.lda, BLOCK_K); if (rc) return rc;
    rc = make_map(&mb, a.B, a.Kr, a.N, a.ldb, BLOCK_K); if (rc) return rc;
  }
  // Epilogue warps: 16 for the launches that also take column sums (linear2's dgrad: mask + bias gradient of linear1),
  // else 8.  Measured INSIDE the step (profiles/r02b_gemm_in_step_tables.txt): with column sums 16 warps 244.6 us against
  // 277.6 us (M = 80384); for the dropout epilogues of out_proj / linear1 16 warps are 8-18 % SLOWER in the step (257 vs
  // 218 us) although the isolated microbenchmark had them 2-4 % ahead (profiles/r02b_gemm_epilogue_warps.jsonl).
  // MAR_TC_EPI_WARPS=8|16 forces one layout (A/B runs).
  static int env_ew = -1;
  if (env_ew < 0) { const char* e = getenv("MAR_TC_EPI_WARPS"); env_ew = e != nullptr ? atoi(e) : 0; }
  const bool heavy = a.colsum != nullptr;
  const int ew = a.out_fp32 ? 8 : (env_ew == 8 || env_ew == 16 ? env_ew : (heavy ? 16 : 8));
  if (!a.out_fp32) {
    const int bc = ew == 16 ? 32 : 64;
    rc = make_map(&mo, a.out, a.M, a.N, a.ldo, 32, bc); if (rc) return rc;
    const void* sp = a.residual != nullptr ? a.residual : a.aux;
    if (sp != nullptr) { rc = make_map(&ms, sp, a.M, a.N, a.residual != nullptr ? a.ldr : a.ldaux, 32, bc); if (rc) return rc; }
    else ms = mo;
  } else {
    mo = ma; ms = ma;   // unused by the fp32 epilogue
  }
  if (!a.a_mn_major) {
    if (a.b_mn_major) return BN == 256 ? launch_cl<256, false, true, bf16>(ma, mb, mo, ms, p, cl, ew, st) : launch_cl<128, false, true, bf16>(ma, mb, mo, ms, p, cl, ew, st);
    if (a.out_fp32) return BN == 256 ? launch_cl<256, false, false, float>(ma, mb, mo, ms, p, cl, 8, st) : launch_cl<128, false, false, float>(ma, mb, mo, ms, p, cl, 8, st);
    return BN == 256 ? launch_cl<256, false, false, bf16>(ma, mb, mo, ms, p, cl, ew, st) : launch_cl<128, false, false, bf16>(ma, mb, mo, ms, p, cl, ew, st);
  }
  return BN == 256 ? launch_cl<256, true, true, float>(ma, mb, mo, ms, p, cl, 8, st) : launch_cl<128, true, true, float>(ma, mb, mo, ms, p, cl, 8, st);
}
