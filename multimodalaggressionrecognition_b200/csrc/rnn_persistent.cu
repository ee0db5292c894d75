// Persistent cluster GRU forward (bf16): the whole T-step recurrence in ONE launch.
//
//   gh_t = h_{t-1}·W_hhᵀ + b_hh ;  r = σ(gi_r + gh_r), z = σ(gi_z + gh_z), n = tanh(gi_n + r·gh_n),
//   h_t = (1 − z)·n + z·h_{t-1}                       (nn.GRU, gate order r,z,n, h0 = 0; models.py:110,122)
//
// W_hh (3H×H bf16, 1.5 MB at H = 512) does not fit one SM, so a CLUSTER of 16 CTAs shares it: CTA c keeps the
// 3·U rows (U = H/16 hidden units × gates r,z,n) of W_hh resident in shared memory for all T steps (96 KB at
// H = 512) and owns those U units of the state.  A cluster serves 8 or 16 batch rows (one or two n = 8 tiles of
// mma.m16n8k16); the batch is spread over ceil(B/16) clusters (B = 64 → 4 clusters: measured, at most 4-5
// clusters of 16 are co-resident on the 148 SMs, so 8 clusters of 8 rows would run in two waves).
// Per step and CTA:
//   1. wait for h_{t-1} (8 × H bf16) to have arrived in the local double buffer (mbarrier, transaction bytes);
//   2. 12 warps: gh slice (3U × 8) = W_slice · h_{t-1}ᵀ with ldmatrix + mma.sync m16n8k16 (warp = one 16-row
//      tile × one half of K), fp32 partials to shared memory;
//   3. 8 warps (warp ↔ batch row, lane ↔ hidden unit): gates in fp32 (the state is carried in fp32 registers;
//      only the GEMM operand is rounded to bf16, as in the step engine), hseq / saved / hprev to HBM;
//   4. the new (batch × U) slice is staged in shared memory and pushed into the h buffer of ALL 16 CTAs with one
//      cp.async.bulk shared::cta → shared::cluster per peer (DSMEM write + remote mbarrier complete_tx in one
//      instruction; 16 mbarrier transactions per step instead of one per 16 B) — no cluster-wide barrier on
//      the critical path; double buffering makes the data arrival the only synchronisation needed.
// gi = x·W_ihᵀ + b_ih for all steps comes from the tcgen05 GEMM (mar_linear_fwd); gi_{t+1} is prefetched into
// registers during step t.
//
// Bound: latency of the dependent chain (per step: smem-resident GEMV-like product, 96 KB of ldmatrix traffic,
// one DSMEM exchange); algorithmic work T·2·B·3H·H FLOP.
#include <cuda_runtime.h>
#include "common.cuh"
#include "rnn.cuh"

namespace {

constexpr int CL = 16;          // CTAs per cluster
constexpr int PAD = 8;          // bf16 elements of row padding (conflict-free ldmatrix)

__device__ __forceinline__ uint32_t s_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t (&r)[2], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t local_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_addr(bar)), "r"(bytes) : "memory");
}
// completion of the phase makes the peers' bulk-copied bytes visible (same contract as a TMA load); the default
// CTA-scope acquire avoids the L1 invalidate a cluster-scope acquire costs on every step
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok) : "r"(s_addr(bar)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

enum { CELL_GRU = 0, CELL_LSTM = 1 };

struct GruParams {     // G = 3 (GRU: r,z,n) or 4 (LSTM: i,f,g,o) gate blocks
  const bf16* gi;      // (B,T,G·H)
  const bf16* w_hh;    // (G·H,H)
  const float* b_hh;   // (G·H)
  bf16* hseq;          // (B,T,H)
  bf16* hprev;         // (B,T,H) or null
  float* saved;        // (B,T,5H) or null   GRU: r,z,n,hn,hp   LSTM: i,f,g,o,c
  int B, T, H;
};

// one bulk copy of `bytes` from local shared memory into a peer CTA's shared memory; the peer's mbarrier
// receives complete_tx(bytes) when the data has landed
__device__ __forceinline__ void bulk_push(uint32_t remote_dst, uint32_t local_src, uint32_t bytes, uint32_t remote_bar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(remote_dst), "r"(local_src), "r"(bytes), "r"(remote_bar) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// U = hidden units per CTA = H / 16 (16 or 32);  NT = 8-row batch tiles per cluster (BG = 8·NT batch rows)
template <int CELL, int U, int NT>
__global__ void __launch_bounds__((CELL == CELL_GRU ? 12 : 16) * 32, 1)
rnn_fwd_persistent_kernel(const GruParams p) {
  constexpr int G = CELL == CELL_GRU ? 3 : 4;
  constexpr int NWARPS = CELL == CELL_GRU ? 12 : 16;
  constexpr int H = U * CL;
  constexpr int BGR = 8 * NT;               // batch rows per cluster
  constexpr int LD = H + PAD;               // W smem row pitch in elements
  constexpr int UP = U + 8;                 // h-slice row pitch in elements (conflict-free ldmatrix)
  constexpr int SLICE = BGR * UP;           // elements of one CTA's (BGR x U) slice block
  constexpr int ROWS = G * U;               // W_hh rows of this CTA (gate blocks of U)
  constexpr int MT = ROWS / 16;             // 16-row tiles
  constexpr int KSPLIT = NWARPS / MT;       // warps per tile along K
  constexpr int KSTEPS = H / 16 / KSPLIT;   // k-steps per warp
  static_assert(NWARPS % MT == 0 && (H / 16) % KSPLIT == 0, "tile split");
  static_assert(NT == 1 || NT == 2, "batch tiles");

  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* sW = reinterpret_cast<bf16*>(smem_raw);                 // [ROWS][LD]
  bf16* sH = sW + ROWS * LD;                                    // [2][CL][BGR][UP]  h_{t-1}, slice-major
  bf16* sOut = sH + 2 * CL * SLICE;                             // [2][BGR][UP]      this CTA's new slice (bulk-copy source)
  float* sG = reinterpret_cast<float*>(sOut + 2 * SLICE);       // [KSPLIT][ROWS][BGR] partial gh
  uint64_t* bars = reinterpret_cast<uint64_t*>(sG + KSPLIT * ROWS * BGR);   // [2]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t c = cluster_rank();
  const int group = blockIdx.x / CL;
  const int b0 = group * BGR;
  const int T = p.T;

  // ---- resident weights: rows g*H + c*U + u  ->  smem row g*U + u
  for (int i = threadIdx.x; i < ROWS * (H / 8); i += blockDim.x) {
    const int r = i / (H / 8), ch = i % (H / 8);
    const int g = r / U, u = r % U;
    const uint4 v = *reinterpret_cast<const uint4*>(p.w_hh + ((int64_t)g * H + c * U + u) * H + ch * 8);
    *reinterpret_cast<uint4*>(sW + r * LD + ch * 8) = v;
  }
  for (int i = threadIdx.x; i < 2 * SLICE; i += blockDim.x) sOut[i] = __float2bfloat16_rn(0.f);   // padding is copied too
  if (threadIdx.x == 0) {
    mbar_init(bars + 0, 1);
    mbar_init(bars + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync();                 // every CTA's barriers exist before anyone pushes data at them

  // ---- gate-thread identity (warps 0 .. U*8/32-1): warp-major over batch rows, lane-major over units;
  //      each gate thread owns NT batch rows (gb, gb + 8)
  constexpr int GATE_WARPS = U * 8 / 32;
  constexpr int ROWS_PER_WARP = 32 / U;                 // batch rows of one 8-row tile handled by one gate warp
  const bool gate_thread = warp < GATE_WARPS;
  const int gb = warp * ROWS_PER_WARP + lane / U;       // batch row within an 8-row tile
  const int gu = lane % U;                              // unit within the CTA slice
  const int j = c * U + gu;                             // hidden unit
  float bh[G];
  float hp[NT];                                    // GRU: fp32 h_{t-1};  LSTM: fp32 cell state c_{t-1}
  bf16 hprev_b[NT];                                // LSTM: bf16 h_{t-1} (what backward's wgrad consumes)
  unsigned short gi_raw[NT][G];                    // raw bf16 bits: converted at use, so the prefetch stays in flight
  bool bvalid[NT];
#pragma unroll
  for (int g = 0; g < G; g++) bh[g] = gate_thread ? p.b_hh[g * H + j] : 0.f;
#pragma unroll
  for (int nt = 0; nt < NT; nt++) {
    hp[nt] = 0.f; hprev_b[nt] = __float2bfloat16_rn(0.f);
    bvalid[nt] = gate_thread && (b0 + nt * 8 + gb) < p.B;
#pragma unroll
    for (int g = 0; g < G; g++) gi_raw[nt][g] = 0;
    if (bvalid[nt]) {
      const unsigned short* gp = reinterpret_cast<const unsigned short*>(p.gi) + ((int64_t)(b0 + nt * 8 + gb) * T) * G * H;
#pragma unroll
      for (int g = 0; g < G; g++) gi_raw[nt][g] = __ldg(gp + g * H + j);
    }
  }
  const int mt = warp % MT, kq = warp / MT;
  const int g8 = lane >> 2, t4 = lane & 3;
  // ldmatrix lane addresses: A = 16 rows of W_slice; B = h rows (lane&7) of batch tile (lane>>4), k chunk (lane>>3)&1
  const uint32_t a_base = s_addr(sW + (mt * 16 + (lane & 15)) * LD + (lane >> 4) * 8);
  const int b_row = (NT == 2 ? (lane >> 4) * 8 : 0) + (lane & 7);
  const int b_chunk = ((lane >> 3) & 1) * 8;

  for (int t = 0; t < T; t++) {
    const int cur = t & 1;                 // h_{t-1} sits in buffer cur^1; the new h_t goes to buffer cur
    // arm the barrier that will collect h_t from all 16 CTAs
    if (threadIdx.x == 0 && t + 1 < T) mbar_expect_tx(bars + cur, CL * SLICE * 2);

    float gh[NT][G];
#pragma unroll
    for (int nt = 0; nt < NT; nt++)
#pragma unroll
      for (int g = 0; g < G; g++) gh[nt][g] = bh[g];
    if (t > 0) {
      mbar_wait_cluster(bars + (cur ^ 1), ((t - 1) >> 1) & 1);
      // ---- partial product: tile mt (16 rows of W_slice) x K range kq, all batch tiles
      const bf16* hbuf = sH + (cur ^ 1) * CL * SLICE;
      float acc[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; nt++) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ks++) {
        const int k0 = (kq * KSTEPS + ks) * 16;
        uint32_t a[4];
        ldsm_x4(a, a_base + k0 * 2);
        const uint32_t b_addr = s_addr(hbuf + (k0 / U) * SLICE + b_row * UP + (k0 % U) + b_chunk);
        if (NT == 2) {
          uint32_t b[4];
          ldsm_x4(b, b_addr);
          const uint32_t b0v[2] = {b[0], b[1]}, b1v[2] = {b[2], b[3]};
          mma16816(acc[0], a, b0v);
          mma16816(acc[NT - 1], a, b1v);
        } else {
          uint32_t b[2];
          ldsm_x2(b, b_addr);
          mma16816(acc[0], a, b);
        }
      }
      float* part = sG + (kq * ROWS + mt * 16) * BGR;
#pragma unroll
      for (int nt = 0; nt < NT; nt++) {
        *reinterpret_cast<float2*>(part + g8 * BGR + nt * 8 + t4 * 2) = make_float2(acc[nt][0], acc[nt][1]);
        *reinterpret_cast<float2*>(part + (g8 + 8) * BGR + nt * 8 + t4 * 2) = make_float2(acc[nt][2], acc[nt][3]);
      }
      if (threadIdx.x < CL) bulk_wait_read_1();      // the push of step t-2 has finished reading sOut[cur]
      __syncthreads();
      if (gate_thread) {
#pragma unroll
        for (int nt = 0; nt < NT; nt++)
#pragma unroll
          for (int q = 0; q < KSPLIT; q++)
#pragma unroll
            for (int g = 0; g < G; g++) gh[nt][g] += sG[(q * ROWS + g * U + gu) * BGR + nt * 8 + gb];
      }
    }

    if (gate_thread) {
#pragma unroll
      for (int nt = 0; nt < NT; nt++) {
        float x[G];
#pragma unroll
        for (int g = 0; g < G; g++) x[g] = __uint_as_float((uint32_t)gi_raw[nt][g] << 16);
        const int64_t row = (int64_t)(b0 + nt * 8 + gb) * T + t;
        bf16 hb;
        if (CELL == CELL_GRU) {
          const float r = sigmoidf_(x[0] + gh[nt][0]);
          const float z = sigmoidf_(x[1] + gh[nt][1]);
          const float n = tanhf(x[2] + r * gh[nt][2]);
          const float h = (1.f - z) * n + z * hp[nt];
          hb = __float2bfloat16_rn(h);
          if (bvalid[nt] && p.saved != nullptr) {
            float* sp = p.saved + row * 5 * H;
            sp[j] = r; sp[H + j] = z; sp[2 * H + j] = n; sp[3 * H + j] = gh[nt][2]; sp[4 * H + j] = hp[nt];
            p.hprev[row * H + j] = __float2bfloat16_rn(hp[nt]);
          }
          hp[nt] = h;
        } else {
          const float ig = sigmoidf_(x[0] + gh[nt][0]);
          const float fg = sigmoidf_(x[1] + gh[nt][1]);
          const float gg = tanhf(x[2] + gh[nt][2]);
          const float og = sigmoidf_(x[G - 1] + gh[nt][G - 1]);
          const float cc = fg * hp[nt] + ig * gg;
          hb = __float2bfloat16_rn(og * tanhf(cc));
          if (bvalid[nt] && p.saved != nullptr) {
            float* sp = p.saved + row * 5 * H;
            sp[j] = ig; sp[H + j] = fg; sp[2 * H + j] = gg; sp[3 * H + j] = og; sp[4 * H + j] = cc;
            p.hprev[row * H + j] = hprev_b[nt];
          }
          hp[nt] = cc;
          hprev_b[nt] = hb;
        }
        if (bvalid[nt]) {
          p.hseq[row * H + j] = hb;
          if (t + 1 < T) {                 // prefetch the next step's input projection
            const unsigned short* gp = reinterpret_cast<const unsigned short*>(p.gi) + (row + 1) * G * H;
#pragma unroll
            for (int g = 0; g < G; g++) gi_raw[nt][g] = __ldg(gp + g * H + j);
          }
        }
        sOut[cur * SLICE + (nt * 8 + gb) * UP + gu] = hb;
      }
      fence_async_smem();                  // the bulk copies below read sOut through the async proxy
    }
    __syncthreads();
    // ---- push this CTA's (BGR x U) slice into the h buffer of all 16 CTAs: one bulk copy + one mbarrier
    //      transaction per peer (DSMEM write and remote complete_tx in one instruction)
    if (threadIdx.x < CL && t + 1 < T) {
      const uint32_t dst = threadIdx.x;
      const uint32_t local_dst = s_addr(sH + (cur * CL + c) * SLICE);
      bulk_push(mapa(local_dst, dst), s_addr(sOut + cur * SLICE), SLICE * 2, mapa(s_addr(bars + cur), dst));
      bulk_commit();
    }
  }
  if (threadIdx.x < CL) bulk_wait_all();
  cluster_sync();                 // no CTA exits while a peer may still write into its shared memory
}

template <int CELL, int U, int NT>
int launch(const GruParams& prm, cudaStream_t st) {
  constexpr int G = CELL == CELL_GRU ? 3 : 4, NWARPS = CELL == CELL_GRU ? 12 : 16;
  constexpr int H = U * CL, LD = H + PAD, ROWS = G * U, MT = ROWS / 16, KSPLIT = NWARPS / MT, BGR = 8 * NT, SLICE = BGR * (U + 8);
  constexpr int SMEM = ROWS * LD * 2 + 2 * CL * SLICE * 2 + 2 * SLICE * 2 + KSPLIT * ROWS * BGR * 4 + 16;
  static_assert(SMEM <= 232448, "persistent GRU: shared memory budget");
  static bool cfg = false;
  if (!cfg) {
    MAR_CUDA(cudaFuncSetAttribute(rnn_fwd_persistent_kernel<CELL, U, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    MAR_CUDA(cudaFuncSetAttribute(rnn_fwd_persistent_kernel<CELL, U, NT>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cfg = true;
  }
  const int groups = (int)ceil_div(prm.B, BGR);
  cudaLaunchConfig_t cfgl = {};
  cfgl.gridDim = dim3((unsigned)(groups * CL));
  cfgl.blockDim = dim3(NWARPS * 32);
  cfgl.dynamicSmemBytes = SMEM;
  cfgl.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfgl.attrs = attr;
  cfgl.numAttrs = 1;
  MAR_CUDA(cudaLaunchKernelEx(&cfgl, rnn_fwd_persistent_kernel<CELL, U, NT>, prm));
  MAR_LAUNCH_CHECK(CELL == CELL_GRU ? "gru_fwd_persistent" : "lstm_fwd_persistent");
  return MAR_OK;
}

// ================================================================================================
// Backward through time, same cluster decomposition.  Per step t (descending), CTA c (units k in its slice):
//   dh    = dhseq[t] + dh_direct + Σ_src partial_src[·, k]        (partials of step t+1, received over DSMEM)
//   dn~ = dh(1−z)(1−n²), dz~ = dh(hp−n)z(1−z), dr~ = dn~·hn·r(1−r);  dgi = (dr~,dz~,dn~), dgh = (dr~,dz~,dn~·r)
//   dh_direct = dh·z ;   partial_c[b, k'] = Σ_{j ∈ slice c} dgh[b,j]·W_hh[j,k']  for ALL k' (the resident W slice
//   read transposed with ldmatrix.trans), then reduce-scattered: block k' ∈ slice c' is bulk-pushed to CTA c'.
// The carry stays fp32 end to end (fp32 partials on the wire); dgi / dgh leave as bf16 for the wgrad / dgrad GEMMs.
// ================================================================================================
constexpr int NWARPS_B = 16;

__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void bulk_wait_read_0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

struct GruBwdParams {
  const bf16* dhseq;   // (B,T,H)
  const float* saved;  // (B,T,5H)   GRU: r, z, n, hn, hp    LSTM: i, f, g, o, c
  const bf16* w_hh;    // (G·H,H)
  bf16* dgi;           // (B,T,G·H)  GRU: d(gi);  LSTM: d(gates)
  bf16* dgh;           // (B,T,3H)   GRU only: d(gh) (differs from dgi in the n block)
  int B, T, H;
};

// LSTM backward through time uses the same decomposition: per step  dc += dh·o·(1−tanh²c);  d_i = dc·g·i(1−i),
// d_f = dc·c_{t-1}·f(1−f),  d_g = dc·i·(1−g²),  d_o = dh·tanh(c)·o(1−o);  the cell carry dc·f stays in a register of
// the owning thread, dh_{t-1} = d(gates)·W_hh is reduce-scattered like the GRU's.
template <int CELL, int U, int NT>
__global__ void __launch_bounds__(NWARPS_B * 32, 1)
rnn_bwd_persistent_kernel(const GruBwdParams p) {
  constexpr int G = CELL == CELL_GRU ? 3 : 4;
  constexpr int H = U * CL;
  constexpr int BGR = 8 * NT;
  constexpr int LD = H + PAD;
  constexpr int ROWS = G * U;               // local gate rows j (K of the transposed product)
  constexpr int DLD = ROWS + 8;             // dgh slice row pitch (bf16)
  constexpr int BLK = BGR * U;              // fp32 elements of one (batch x unit-slice) block
  constexpr int MTILES = H / 16;            // 16-column tiles of the output k'
  constexpr int MPW = MTILES / NWARPS_B;    // tiles per warp
  static_assert(MTILES % NWARPS_B == 0, "tile split");

  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* sW = reinterpret_cast<bf16*>(smem_raw);                    // [ROWS][LD]
  bf16* sD = sW + ROWS * LD;                                       // [BGR][DLD]  dgh slice of this step (MMA B operand)
  float* sP = reinterpret_cast<float*>(sD + BGR * DLD);            // [CL][BGR][U] partials, block per destination CTA
  float* sR = sP + CL * BLK;                                       // [2][CL][BGR][U] received partial blocks
  uint64_t* bars = reinterpret_cast<uint64_t*>(sR + 2 * CL * BLK); // [2]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t c = cluster_rank();
  const int b0 = (blockIdx.x / CL) * BGR;
  const int T = p.T;

  for (int i = threadIdx.x; i < ROWS * (H / 8); i += blockDim.x) {
    const int r = i / (H / 8), ch = i % (H / 8);
    const int g = r / U, u = r % U;
    *reinterpret_cast<uint4*>(sW + r * LD + ch * 8) =
        *reinterpret_cast<const uint4*>(p.w_hh + ((int64_t)g * H + c * U + u) * H + ch * 8);
  }
  if (threadIdx.x == 0) {
    mbar_init(bars + 0, 1);
    mbar_init(bars + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync();

  constexpr int GATE_WARPS = U * 8 / 32;
  constexpr int ROWS_PER_WARP = 32 / U;
  const bool gate_thread = warp < GATE_WARPS;
  const int gb = warp * ROWS_PER_WARP + lane / U;
  const int gu = lane % U;
  const int j = c * U + gu;
  float dh_direct[NT];      // GRU: dh·z carried to the previous step;  LSTM: the cell-state carry dc·f
  float sv[NT][5];
  bool bvalid[NT];
#pragma unroll
  for (int nt = 0; nt < NT; nt++) {
    dh_direct[nt] = 0.f;
    bvalid[nt] = gate_thread && (b0 + nt * 8 + gb) < p.B;
#pragma unroll
    for (int q = 0; q < 5; q++) sv[nt][q] = 0.f;
    if (bvalid[nt]) {
      const float* s = p.saved + ((int64_t)(b0 + nt * 8 + gb) * T + (T - 1)) * 5 * H;
#pragma unroll
      for (int q = 0; q < 5; q++) sv[nt][q] = __ldg(s + q * H + j);
    }
  }
  const int g8 = lane >> 2, t4 = lane & 3;
  // ldmatrix.trans lane address of A = W_sliceᵀ: matrix q = lane>>3: K rows (q>>1)*8 + (lane&7), M columns (q&1)*8
  const int a_krow = ((lane >> 4) & 1) * 8 + (lane & 7);
  const int a_mcol = ((lane >> 3) & 1) * 8;
  const int b_row = (NT == 2 ? (lane >> 4) * 8 : 0) + (lane & 7);
  const int b_chunk = ((lane >> 3) & 1) * 8;

  for (int t = T - 1; t >= 0; t--) {
    const int it = T - 1 - t;              // iteration counter
    const int cur = it & 1;                // partials produced in this iteration land in sR[cur] of the peers
    if (threadIdx.x == 0 && t > 0) mbar_expect_tx(bars + cur, CL * BLK * 4);
    if (gate_thread) {
      if (it > 0) mbar_wait_cluster(bars + (cur ^ 1), ((it - 1) >> 1) & 1);
#pragma unroll
      for (int nt = 0; nt < NT; nt++) {
        float dh = CELL == CELL_GRU ? dh_direct[nt] : 0.f;
        if (it > 0) {
          const float* rb = sR + (cur ^ 1) * CL * BLK + (nt * 8 + gb) * U + gu;
#pragma unroll
          for (int src = 0; src < CL; src++) dh += rb[src * BLK];
        }
        const int64_t row = (int64_t)(b0 + nt * 8 + gb) * T + t;
        if (bvalid[nt]) dh += __bfloat162float(p.dhseq[row * H + j]);
        bf16 dq[G];                                  // what the recurrent product consumes (d gh / d gates)
        if (CELL == CELL_GRU) {
          const float r = sv[nt][0], z = sv[nt][1], n = sv[nt][2], hn = sv[nt][3], hp = sv[nt][4];
          const float dn_pre = dh * (1.f - z) * (1.f - n * n);
          const float dz_pre = dh * (hp - n) * z * (1.f - z);
          const float dr_pre = dn_pre * hn * r * (1.f - r);
          dq[0] = __float2bfloat16_rn(dr_pre); dq[1] = __float2bfloat16_rn(dz_pre); dq[2] = __float2bfloat16_rn(dn_pre * r);
          dh_direct[nt] = dh * z;
          if (bvalid[nt]) {
            bf16* a = p.dgi + row * 3 * H;
            bf16* q = p.dgh + row * 3 * H;
            a[j] = dq[0]; a[H + j] = dq[1]; a[2 * H + j] = __float2bfloat16_rn(dn_pre);
            q[j] = dq[0]; q[H + j] = dq[1]; q[2 * H + j] = dq[2];
          }
        } else {
          const float ig = sv[nt][0], fg = sv[nt][1], gg = sv[nt][2], og = sv[nt][3], cc = sv[nt][4];
          float cp = 0.f;
          if (bvalid[nt] && t > 0) cp = __ldg(p.saved + (row - 1) * 5 * H + 4 * H + j);
          const float tc = tanhf(cc);
          const float dc = dh_direct[nt] + dh * og * (1.f - tc * tc);
          dq[0] = __float2bfloat16_rn(dc * gg * ig * (1.f - ig));
          dq[1] = __float2bfloat16_rn(dc * cp * fg * (1.f - fg));
          dq[2] = __float2bfloat16_rn(dc * ig * (1.f - gg * gg));
          dq[G - 1] = __float2bfloat16_rn(dh * tc * og * (1.f - og));
          dh_direct[nt] = dc * fg;
          if (bvalid[nt]) {
            bf16* a = p.dgi + row * G * H;
#pragma unroll
            for (int g = 0; g < G; g++) a[g * H + j] = dq[g];
          }
        }
        if (bvalid[nt] && t > 0) {                   // prefetch the saved gates of the next (earlier) step
          const float* sp = p.saved + (row - 1) * 5 * H;
#pragma unroll
          for (int e = 0; e < 5; e++) sv[nt][e] = __ldg(sp + e * H + j);
        }
        bf16* d = sD + (nt * 8 + gb) * DLD;
#pragma unroll
        for (int g = 0; g < G; g++) d[g * U + gu] = dq[g];
      }
    }
    if (t == 0) break;                     // no earlier step to carry into (uniform)
    if (threadIdx.x < CL) bulk_wait_read_0();      // the previous pushes have finished reading sP
    __syncthreads();
    // ---- partialᵀ (k' x batch) = W_sliceᵀ (k' x 3U) · dghᵀ (3U x batch): warp owns MPW 16-column tiles of k'
    {
      float acc[MPW][NT][4];
#pragma unroll
      for (int mi = 0; mi < MPW; mi++)
#pragma unroll
        for (int nt = 0; nt < NT; nt++) { acc[mi][nt][0] = acc[mi][nt][1] = acc[mi][nt][2] = acc[mi][nt][3] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < ROWS / 16; ks++) {
        uint32_t b[4];
        const uint32_t b_addr = s_addr(sD + b_row * DLD + ks * 16 + b_chunk);
        if (NT == 2) ldsm_x4(b, b_addr);
        else { uint32_t b2[2]; ldsm_x2(b2, b_addr); b[0] = b2[0]; b[1] = b2[1]; b[2] = b[3] = 0; }
        const uint32_t b0v[2] = {b[0], b[1]}, b1v[2] = {b[2], b[3]};
#pragma unroll
        for (int mi = 0; mi < MPW; mi++) {
          const int m0 = (warp * MPW + mi) * 16;
          uint32_t a[4];
          ldsm_x4_t(a, s_addr(sW + (ks * 16 + a_krow) * LD + m0 + a_mcol));
          mma16816(acc[mi][0], a, b0v);
          if (NT == 2) mma16816(acc[mi][NT - 1], a, b1v);
        }
      }
#pragma unroll
      for (int mi = 0; mi < MPW; mi++) {
        const int m0 = (warp * MPW + mi) * 16;
#pragma unroll
        for (int nt = 0; nt < NT; nt++) {
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const int m = m0 + g8 + (e >> 1) * 8;            // output unit k'
            const int bb = nt * 8 + t4 * 2 + (e & 1);        // batch row
            sP[((m / U) * BGR + bb) * U + (m % U)] = acc[mi][nt][e];
          }
        }
      }
    }
    fence_async_smem();
    __syncthreads();
    if (threadIdx.x < CL) {
      const uint32_t dst = threadIdx.x;
      const uint32_t local_dst = s_addr(sR + (cur * CL + c) * BLK);
      bulk_push(mapa(local_dst, dst), s_addr(sP + dst * BLK), BLK * 4, mapa(s_addr(bars + cur), dst));
      bulk_commit();
    }
  }
  if (threadIdx.x < CL) bulk_wait_all();
  cluster_sync();
}

template <int CELL, int U, int NT>
int launch_bwd(const GruBwdParams& prm, cudaStream_t st) {
  constexpr int G = CELL == CELL_GRU ? 3 : 4;
  constexpr int H = U * CL, LD = H + PAD, ROWS = G * U, BGR = 8 * NT, BLK = BGR * U;
  constexpr int SMEM = ROWS * LD * 2 + BGR * (ROWS + 8) * 2 + 3 * CL * BLK * 4 + 16;
  static_assert(SMEM <= 232448, "persistent GRU backward: shared memory budget");
  static bool cfg = false;
  if (!cfg) {
    MAR_CUDA(cudaFuncSetAttribute(rnn_bwd_persistent_kernel<CELL, U, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    MAR_CUDA(cudaFuncSetAttribute(rnn_bwd_persistent_kernel<CELL, U, NT>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cfg = true;
  }
  const int groups = (int)ceil_div(prm.B, BGR);
  cudaLaunchConfig_t cfgl = {};
  cfgl.gridDim = dim3((unsigned)(groups * CL));
  cfgl.blockDim = dim3(NWARPS_B * 32);
  cfgl.dynamicSmemBytes = SMEM;
  cfgl.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfgl.attrs = attr;
  cfgl.numAttrs = 1;
  MAR_CUDA(cudaLaunchKernelEx(&cfgl, rnn_bwd_persistent_kernel<CELL, U, NT>, prm));
  MAR_LAUNCH_CHECK(CELL == CELL_GRU ? "gru_bwd_persistent" : "lstm_bwd_persistent");
  return MAR_OK;
}

}  // namespace

bool gru_persistent_supported(int64_t B, int64_t T, int64_t H, int dtype) {
  if (dtype != MAR_BF16) return false;
  if (!(H == 512 || H == 256)) return false;
  if (B < 1 || T < 1 || B > (1 << 20) || T > (1 << 24)) return false;
  return true;
}

int gru_fwd_persistent(const void* gi, const void* w_hh, const float* b_hh, void* hseq, void* hprev, float* saved, int64_t B,
                       int64_t T, int64_t H, cudaStream_t st) {
  MAR_CHECK_ARG(((uintptr_t)w_hh % 16 == 0), "gru (persistent engine): w_hh must be 16 B aligned");
  GruParams prm;
  prm.gi = (const bf16*)gi; prm.w_hh = (const bf16*)w_hh; prm.b_hh = b_hh; prm.hseq = (bf16*)hseq; prm.hprev = (bf16*)hprev;
  prm.saved = saved; prm.B = (int)B; prm.T = (int)T; prm.H = (int)H;
  // 16 batch rows per cluster once the batch would otherwise need more clusters than fit the chip at once
  const bool wide = B > 32;
  if (H == 512) return wide ? launch<CELL_GRU, 32, 2>(prm, st) : launch<CELL_GRU, 32, 1>(prm, st);
  if (H == 256) return wide ? launch<CELL_GRU, 16, 2>(prm, st) : launch<CELL_GRU, 16, 1>(prm, st);
  MAR_UNSUPPORTED("gru (persistent engine): hidden size %lld", (long long)H);
}

int gru_bwd_persistent(const void* dhseq, const float* saved, const void* w_hh, void* dgi, void* dgh, int64_t B, int64_t T,
                       int64_t H, cudaStream_t st) {
  MAR_CHECK_ARG(((uintptr_t)w_hh % 16 == 0), "gru (persistent engine): w_hh must be 16 B aligned");
  GruBwdParams prm;
  prm.dhseq = (const bf16*)dhseq; prm.saved = saved; prm.w_hh = (const bf16*)w_hh; prm.dgi = (bf16*)dgi; prm.dgh = (bf16*)dgh;
  prm.B = (int)B; prm.T = (int)T; prm.H = (int)H;
  const bool wide = B > 32;
  if (H == 512) return wide ? launch_bwd<CELL_GRU, 32, 2>(prm, st) : launch_bwd<CELL_GRU, 32, 1>(prm, st);
  if (H == 256) return wide ? launch_bwd<CELL_GRU, 16, 2>(prm, st) : launch_bwd<CELL_GRU, 16, 1>(prm, st);
  MAR_UNSUPPORTED("gru backward (persistent engine): hidden size %lld", (long long)H);
}

int lstm_fwd_persistent(const void* gi, const void* w_hh, const float* b_hh, void* hseq, void* hprev, float* saved, int64_t B,
                        int64_t T, int64_t H, cudaStream_t st) {
  MAR_CHECK_ARG(((uintptr_t)w_hh % 16 == 0), "lstm (persistent engine): w_hh must be 16 B aligned");
  GruParams prm;
  prm.gi = (const bf16*)gi; prm.w_hh = (const bf16*)w_hh; prm.b_hh = b_hh; prm.hseq = (bf16*)hseq; prm.hprev = (bf16*)hprev;
  prm.saved = saved; prm.B = (int)B; prm.T = (int)T; prm.H = (int)H;
  const bool wide = B > 32;
  if (H == 512) return wide ? launch<CELL_LSTM, 32, 2>(prm, st) : launch<CELL_LSTM, 32, 1>(prm, st);
  if (H == 256) return wide ? launch<CELL_LSTM, 16, 2>(prm, st) : launch<CELL_LSTM, 16, 1>(prm, st);
  MAR_UNSUPPORTED("lstm (persistent engine): hidden size %lld", (long long)H);
}

int lstm_bwd_persistent(const void* dhseq, const float* saved, const void* w_hh, void* dgates, int64_t B, int64_t T, int64_t H,
                        cudaStream_t st) {
  MAR_CHECK_ARG(((uintptr_t)w_hh % 16 == 0), "lstm (persistent engine): w_hh must be 16 B aligned");
  GruBwdParams prm;
  prm.dhseq = (const bf16*)dhseq; prm.saved = saved; prm.w_hh = (const bf16*)w_hh; prm.dgi = (bf16*)dgates; prm.dgh = nullptr;
  prm.B = (int)B; prm.T = (int)T; prm.H = (int)H;
  const bool wide = B > 32;
  (void)wide;
  if (H == 512) return launch_bwd<CELL_LSTM, 32, 1>(prm, st);   // 4·32 rows of W_hh + the fp32 exchange buffers of 16 batch rows exceed 227 KB
  if (H == 256) return wide ? launch_bwd<CELL_LSTM, 16, 2>(prm, st) : launch_bwd<CELL_LSTM, 16, 1>(prm, st);
  MAR_UNSUPPORTED("lstm backward (persistent engine): hidden size %lld", (long long)H);
}
