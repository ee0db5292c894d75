// Data-parallel gradient exchange fused with the optimizer, over NVLink peer memory (one process per GPU, one node).
//
//   grad (fp32, local)  ->  bf16 wire copy  ->  reduce-scatter: rank r sums shard r over all ranks' wire copies
//   (peer LOADS)  ->  all-gather: the bf16 sum of shard r is STORED into every rank's `red` buffer  ->  per-parameter
//   Adam over the whole flat buffer from `red`, refreshing the bf16 mirror of the parameters.
//
// ONE kernel per step replaces: the fp32->bf16 cast, ncclAllReduce (39 MB for the reference's fusion model), the
// bf16->fp32 cast, the Adam tick and the Adam kernel.  Every rank applies the same bf16-rounded sum (the owner of a
// shard uses the value it pushed, not its fp32 sum), so parameters stay bit-identical across ranks, as with NCCL's
// bf16 all-reduce; shard sums are taken in rank order 0..N-1 in fp32 (deterministic).
//
// Each rank owns one SYMMETRIC block (cudaMalloc + CUDA IPC, mar_peer_*):
//   [ flags: 2 x MAX_RANKS u32 (arrival epochs A, B) | pad to 1 KB | wire: n bf16 | red: n bf16 ]
// Synchronisation (epoch e = number of exchanges so far + 1, kept on the device so a captured graph replays):
//   A: "my wire copy is complete, my red buffer is free"  — written into every peer's flagsA[me] with st.release.sys
//   B: "my shard's sums are in your red buffer, I am done reading your wire copy"
// Inside a rank the phases are separated by grid barriers (all CTAs co-resident: cooperative launch, one CTA per SM).
// Every spin has a time-out (a dead peer must not hang the GPU): on expiry the kernel raises ctrl.error and runs on.
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

namespace {

constexpr int MAX_RANKS = 16;
constexpr int FLAG_BYTES = 1024;
constexpr int THREADS = 512;
// Spins give up after this long (a dead peer must not hang the GPU).  Ranks legitimately arrive seconds apart (graph
// capture, a slow data loader): the default is generous; MAR_DP_TIMEOUT_S overrides it.
constexpr uint64_t DEFAULT_TIMEOUT_S = 60;

struct Ctrl {            // local device memory (zero-initialised by mar_dp_ctrl_init)
  uint32_t epoch;        // exchanges completed
  uint32_t error;        // 1: a spin timed out
  unsigned long long grid_arrivals;   // monotone grid-barrier counter
  unsigned long long readers;         // CTAs that have read seg_steps this epoch (monotone)
};
__constant__ uint64_t c_timeout_ns = DEFAULT_TIMEOUT_S * 1000000000ull;

struct Peers {
  uint8_t* base[MAX_RANKS];
};

__device__ __forceinline__ uint64_t now_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// peer memory: bypass L1 (another GPU writes it), 16 B
__device__ __forceinline__ uint4 ld_peer(const void* p) {
  uint4 v;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_peer(void* p, const uint4& v) {
  asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// all CTAs of the grid (co-resident) meet; `target` = arrivals expected once everybody is here
__device__ __forceinline__ void grid_barrier(Ctrl* ctrl, unsigned long long target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();        // this CTA's stores (some into peer memory) before its arrival
    atomicAdd(&ctrl->grid_arrivals, 1ull);
    const uint64_t t0 = now_ns();
    while (ld_acquire_gpu(&ctrl->grid_arrivals) < target) {
      if (now_ns() - t0 > c_timeout_ns) { ctrl->error = 1; break; }
    }
  }
  __syncthreads();
}

// thread 0 of CTA 0, after a grid barrier: everything this rank wrote (also into peer memory) is visible system-wide,
// then raise my flag in every rank's block
__device__ __forceinline__ void signal_all(const Peers& peers, int world, int rank, int which, uint32_t epoch) {
  __threadfence_system();
  for (int r = 0; r < world; r++)
    st_release_sys(reinterpret_cast<uint32_t*>(peers.base[r]) + which * MAX_RANKS + rank, epoch);
}
// one thread per CTA polls this rank's OWN flags (peers wrote them) — every CTA waits for itself, no extra grid barrier
__device__ __forceinline__ void wait_all(const Peers& peers, Ctrl* ctrl, int world, int rank, int which, uint32_t epoch) {
  if (threadIdx.x < world) {
    const uint32_t* f = reinterpret_cast<const uint32_t*>(peers.base[rank]) + which * MAX_RANKS + threadIdx.x;
    const uint64_t t0 = now_ns();
    while ((int32_t)(ld_acquire_sys(f) - epoch) < 0) {
      if (now_ns() - t0 > c_timeout_ns) { ctrl->error = 1; break; }
    }
  }
  __syncthreads();
}

struct AdamArgs {
  float* p; float* m; float* v;
  const int32_t* chunk_seg;
  float* seg_steps;
  int chunk_shift, nseg;
  float lr, b1, b2, eps;
  bf16* mirror;
  float* grad_out;     // nullable: the reduced gradient, widened, back into the fp32 gradient buffer (for callers that read .grad)
};

__global__ void __launch_bounds__(THREADS, 1)
dp_allreduce_adam_kernel(const float* grad, int64_t n, int world, int rank, const Peers peers, Ctrl* ctrl,
                         const AdamArgs ad) {
  extern __shared__ float2 s_coef[];        // [nseg] (step size, 1/sqrt(bias correction 2)); x == 0: parameter skipped
  const uint32_t epoch = ctrl->epoch + 1;    // read before anybody can bump it (bumped after the last grid barrier)
  const unsigned long long G = gridDim.x;
  const unsigned long long bar_base = (unsigned long long)(epoch - 1) * 2ull * G;
  const int64_t tid = (int64_t)blockIdx.x * THREADS + threadIdx.x;
  const int64_t nthreads = (int64_t)G * THREADS;
  const int64_t nvec = n / 8;                // n is a multiple of 64
  bf16* wire_me = reinterpret_cast<bf16*>(peers.base[rank] + FLAG_BYTES);
  const bf16* red_me = wire_me + n;

  // ---- phase 0: fp32 gradient -> bf16 wire copy (local)
  for (int64_t i = tid; i < nvec; i += nthreads) {
    float v[8];
    Vec8<float>::load(grad + i * 8, v);
    Vec8<bf16>::store(wire_me + i * 8, v);
  }
  grid_barrier(ctrl, bar_base + G);
  if (blockIdx.x == 0 && threadIdx.x == 0) signal_all(peers, world, rank, 0, epoch);
  wait_all(peers, ctrl, world, rank, 0, epoch);

  // ---- phase 1: reduce my shard over all ranks' wire copies (peer loads), push the bf16 sum to every rank's red buffer
  const int64_t per = (nvec + world - 1) / world;
  const int64_t v0 = per * rank, v1 = min(nvec, v0 + per);
  for (int64_t i = v0 + tid; i < v1; i += nthreads) {
    uint4 in[MAX_RANKS];
#pragma unroll
    for (int r = 0; r < MAX_RANKS; r++)
      if (r < world) in[r] = ld_peer(peers.base[r] + FLAG_BYTES + i * 16);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = 0.f;
#pragma unroll
    for (int r = 0; r < MAX_RANKS; r++) {
      if (r < world) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&in[r]);
#pragma unroll
        for (int j = 0; j < 4; j++) { const float2 f = __bfloat1622float2(h[j]); acc[2 * j] += f.x; acc[2 * j + 1] += f.y; }
      }
    }
    uint4 out;
    out.x = pack_bf16x2(acc[0], acc[1]); out.y = pack_bf16x2(acc[2], acc[3]);
    out.z = pack_bf16x2(acc[4], acc[5]); out.w = pack_bf16x2(acc[6], acc[7]);
#pragma unroll
    for (int r = 0; r < MAX_RANKS; r++)
      if (r < world) st_peer(peers.base[r] + FLAG_BYTES + (n + i * 8) * 2, out);
  }
  grid_barrier(ctrl, bar_base + 2 * G);
  if (blockIdx.x == 0 && threadIdx.x == 0) signal_all(peers, world, rank, 1, epoch);
  wait_all(peers, ctrl, world, rank, 1, epoch);

  // ---- phase 2: per-parameter Adam over the whole buffer from the reduced gradient (torch.optim.Adam semantics as in
  // adam_seg_kernel: a parameter whose reduced "received a gradient" flag is 0 is skipped and keeps its step count)
  for (int s = threadIdx.x; s < ad.nseg; s += THREADS) {
    float2 c = make_float2(0.f, 0.f);
    if (__bfloat162float(red_me[s]) > 0.f) {
      const float step = ad.seg_steps[s] + 1.f;
      c = make_float2(ad.lr / (1.f - powf(ad.b1, step)), rsqrtf(1.f - powf(ad.b2, step)));
    }
    s_coef[s] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence(); atomicAdd(&ctrl->readers, 1ull); }
  for (int64_t i = tid; i < nvec; i += nthreads) {
    const int64_t e = i * 8;
    const int seg = ad.chunk_seg[e >> ad.chunk_shift];
    if (seg < 0) continue;
    const float2 coef = s_coef[seg];
    if (coef.x == 0.f) continue;
    float g[8], p[8], m[8], v[8];
    Vec8<bf16>::load(red_me + e, g);
    Vec8<float>::load(ad.p + e, p);
    Vec8<float>::load(ad.m + e, m);
    Vec8<float>::load(ad.v + e, v);
#pragma unroll
    for (int j = 0; j < 8; j++) {
      m[j] = ad.b1 * m[j] + (1.f - ad.b1) * g[j];
      v[j] = ad.b2 * v[j] + (1.f - ad.b2) * g[j] * g[j];
      p[j] -= coef.x * m[j] / (sqrtf(v[j]) * coef.y + ad.eps);
    }
    Vec8<float>::store(ad.p + e, p);
    Vec8<float>::store(ad.m + e, m);
    Vec8<float>::store(ad.v + e, v);
    if (ad.mirror != nullptr) Vec8<bf16>::store(ad.mirror + e, p);
    if (ad.grad_out != nullptr) Vec8<float>::store(ad.grad_out + e, g);
  }
  // CTA 0 advances the step counts once every CTA has read the old ones, then closes the epoch
  if (blockIdx.x == 0) {
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint64_t t0 = now_ns();
      while (ld_acquire_gpu(&ctrl->readers) < (unsigned long long)epoch * G) {
        if (now_ns() - t0 > c_timeout_ns) { ctrl->error = 1; break; }
      }
    }
    __syncthreads();
    for (int s = threadIdx.x; s < ad.nseg; s += THREADS)
      if (s_coef[s].x != 0.f) ad.seg_steps[s] += 1.f;
    if (threadIdx.x == 0) ctrl->epoch = epoch;
  }
}

inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace

extern "C" {

int64_t mar_dp_block_bytes(int64_t n) { return n < 0 ? 0 : (int64_t)FLAG_BYTES + 4 * n; }
int64_t mar_dp_ctrl_bytes(void) { return (int64_t)sizeof(Ctrl); }

int mar_peer_alloc(void** ptr, int64_t bytes) {
  MAR_CHECK_ARG(ptr && bytes > 0, "mar_peer_alloc: bad arguments");
  MAR_CUDA(cudaMalloc(ptr, (size_t)bytes));
  MAR_CUDA(cudaMemset(*ptr, 0, (size_t)bytes));
  return MAR_OK;
}
int mar_peer_free(void* ptr) {
  if (ptr) MAR_CUDA(cudaFree(ptr));
  return MAR_OK;
}
int mar_peer_export(void* ptr, void* handle64) {
  MAR_CHECK_ARG(ptr && handle64, "mar_peer_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  cudaIpcMemHandle_t h;
  MAR_CUDA(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle64, &h, 64);
  return MAR_OK;
}
int mar_peer_import(const void* handle64, void** ptr) {
  MAR_CHECK_ARG(ptr && handle64, "mar_peer_import: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  MAR_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return MAR_OK;
}
int mar_peer_close(void* ptr) {
  if (ptr) MAR_CUDA(cudaIpcCloseMemHandle(ptr));
  return MAR_OK;
}

int mar_dp_allreduce_adam(float* grad, int write_back, int64_t n, int world, int rank, void* const* blocks, void* ctrl,
                          float* param, float* exp_avg, float* exp_avg_sq, const int32_t* chunk_seg, float* seg_steps,
                          int chunk, int nseg, float lr, float beta1, float beta2, float eps, void* bf16_mirror,
                          void* stream) {
  MAR_CHECK_ARG(grad && blocks && ctrl && param && exp_avg && exp_avg_sq && chunk_seg && seg_steps && n > 0 && nseg > 0,
                "mar_dp_allreduce_adam: bad arguments");
  MAR_CHECK_ARG(world >= 2 && world <= MAX_RANKS && rank >= 0 && rank < world, "mar_dp_allreduce_adam: 2..%d ranks", MAX_RANKS);
  MAR_CHECK_ARG(chunk >= 8 && (chunk & (chunk - 1)) == 0 && n % chunk == 0 && nseg <= n,
                "mar_dp_allreduce_adam: chunk must be a power of two >= 8 that divides n");
  MAR_CHECK_ARG(((uintptr_t)grad % 32 == 0) && ((uintptr_t)param % 32 == 0) && ((uintptr_t)exp_avg % 32 == 0) &&
                    ((uintptr_t)exp_avg_sq % 32 == 0) && ((uintptr_t)bf16_mirror % 16 == 0),
                "mar_dp_allreduce_adam: buffers must be 32 B aligned");
  Peers peers;
  for (int r = 0; r < MAX_RANKS; r++) peers.base[r] = nullptr;
  for (int r = 0; r < world; r++) {
    MAR_CHECK_ARG(blocks[r] != nullptr && ((uintptr_t)blocks[r] % 256) == 0, "mar_dp_allreduce_adam: block %d missing or misaligned", r);
    peers.base[r] = reinterpret_cast<uint8_t*>(blocks[r]);
  }
  AdamArgs ad;
  ad.p = param; ad.m = exp_avg; ad.v = exp_avg_sq; ad.chunk_seg = chunk_seg; ad.seg_steps = seg_steps;
  ad.chunk_shift = 0;
  while ((1 << ad.chunk_shift) < chunk) ad.chunk_shift++;
  ad.nseg = nseg; ad.lr = lr; ad.b1 = beta1; ad.b2 = beta2; ad.eps = eps; ad.mirror = reinterpret_cast<bf16*>(bf16_mirror);
  ad.grad_out = write_back ? grad : nullptr;
  const size_t smem = (size_t)nseg * sizeof(float2);
  MAR_CHECK_ARG(smem <= 48 * 1024, "mar_dp_allreduce_adam: more than 6144 parameters");
  Ctrl* c = reinterpret_cast<Ctrl*>(ctrl);
  static bool timeout_set = false;
  if (!timeout_set) {
    timeout_set = true;
    const char* e = getenv("MAR_DP_TIMEOUT_S");
    if (e != nullptr && atof(e) > 0) {
      const uint64_t ns = (uint64_t)(atof(e) * 1e9);
      MAR_CUDA(cudaMemcpyToSymbol(c_timeout_ns, &ns, sizeof(ns)));
    }
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)mar_sm_count(), 1, 1);      // one CTA per SM, all co-resident (the grid barriers need it)
  cfg.blockDim = dim3(THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = S(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  MAR_CUDA(cudaLaunchKernelEx(&cfg, dp_allreduce_adam_kernel, (const float*)grad, n, world, rank, peers, c, ad));
  MAR_LAUNCH_CHECK("dp_allreduce_adam");
  return MAR_OK;
}

/* ctrl[1] != 0: a spin inside the exchange kernel timed out (a peer never arrived).  Synchronises the stream. */
int mar_dp_check(void* ctrl, void* stream) {
  MAR_CHECK_ARG(ctrl, "mar_dp_check: null");
  Ctrl h;
  MAR_CUDA(cudaMemcpyAsync(&h, ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, S(stream)));
  MAR_CUDA(cudaStreamSynchronize(S(stream)));
  if (h.error != 0) {
    mar_set_error("data-parallel exchange: a wait on a peer timed out (after exchange %u; MAR_DP_TIMEOUT_S, default %llu s)", h.epoch, (unsigned long long)DEFAULT_TIMEOUT_S);
    return MAR_ERR_CUDA;
  }
  return MAR_OK;
}

}  // extern "C"
