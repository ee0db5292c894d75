// SIMT fp32-math attention (forward, dQ, dK/dV) over the packed in-projection output.
// This is the fp32-mode engine (parity 1e-4 needs fp32 products) and the checker for the
// tensor-core engine.  Flash-style: scores never reach HBM; one warp owns a query (forward, dQ) or
// a key (dK/dV), lanes run over the 32 keys (queries) of the staged tile.
#include "common.cuh"
#include "attention.cuh"

namespace {

constexpr int TILE = 32;     // keys (or queries) staged per step
constexpr int ROWS = 32;     // queries (or keys) owned by a CTA: 4 warps x 8
constexpr int RPW = 8;       // rows per warp
constexpr int NJMAX = 4;     // dh <= 128

struct AttnDims {
  int64_t B, T, H, dh;
};

template <typename T>
__device__ __forceinline__ void load_rows_to_smem(float* dst, int ld, const T* src, int64_t row_stride, int64_t t0,
                                                   int64_t Tn, int dh, float scale) {
  // dst[r][c] = src[(t0+r)*row_stride + c] * scale for r < 32, zero beyond Tn.
  for (int idx = threadIdx.x; idx < TILE * dh; idx += blockDim.x) {
    int r = idx / dh, c = idx % dh;
    int64_t t = t0 + r;
    dst[r * ld + c] = t < Tn ? to_f32<T>(src[t * row_stride + c]) * scale : 0.f;
  }
}

// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128)
attn_fwd_simt_kernel(const T* __restrict__ qkv, const uint8_t* __restrict__ key_mask, T* __restrict__ out,
                     float* __restrict__ lse, AttnDims dm, float p_drop, const uint32_t* __restrict__ dbits) {
  extern __shared__ float sm[];
  const int dh = (int)dm.dh, ldk = dh + 1;
  float* Ks = sm;
  float* Vs = Ks + TILE * ldk;
  float* Qs = Vs + TILE * ldk;
  float* Ps = Qs + ROWS * dh;
  const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
  const int64_t bh = blockIdx.y, b = bh / dm.H, h = bh % dm.H;
  const int64_t d = dm.H * dm.dh, rs = 3 * d;
  const int64_t q0 = (int64_t)blockIdx.x * ROWS;
  const T* qbase = qkv + b * dm.T * rs + h * dm.dh;
  const float scale = rsqrtf((float)dh);
  const bool drop = p_drop > 0.f;
  const DropBits dk = make_drop_bits(dbits, dm.T, p_drop);

  load_rows_to_smem<T>(Qs, dh, qbase, rs, q0, dm.T, dh, scale);

  float m[RPW], l[RPW], o[RPW][NJMAX];
#pragma unroll
  for (int i = 0; i < RPW; i++) {
    m[i] = -INFINITY; l[i] = 0.f;
#pragma unroll
    for (int j = 0; j < NJMAX; j++) o[i][j] = 0.f;
  }

  for (int64_t k0 = 0; k0 < dm.T; k0 += TILE) {
    __syncthreads();
    load_rows_to_smem<T>(Ks, ldk, qbase + d, rs, k0, dm.T, dh, 1.f);
    load_rows_to_smem<T>(Vs, ldk, qbase + 2 * d, rs, k0, dm.T, dh, 1.f);
    __syncthreads();
    const int64_t kk = k0 + lane;
    const bool valid = kk < dm.T && !(key_mask != nullptr && key_mask[b * dm.T + kk]);
#pragma unroll
    for (int i = 0; i < RPW; i++) {
      const int qi = warp * RPW + i;
      const int64_t q = q0 + qi;
      if (q >= dm.T) break;   // warp-uniform
      float s = 0.f;
      for (int c = 0; c < dh; c++) s = fmaf(Qs[qi * dh + c], Ks[lane * ldk + c], s);
      s = valid ? s : -INFINITY;
      const float m_new = fmaxf(m[i], warp_max(s));
      float p, corr;
      if (m_new == -INFINITY) { p = 0.f; corr = 1.f; }
      else { p = valid ? expf(s - m_new) : 0.f; corr = expf(m[i] - m_new); }
      l[i] = l[i] * corr + warp_sum(p);
      m[i] = m_new;
      if (drop) p = dropbit(dk, (int64_t)bh * dm.T + q, (int)kk) ? p * dk.scale : 0.f;
      Ps[warp * TILE + lane] = p;
      __syncwarp();
#pragma unroll
      for (int j = 0; j < NJMAX; j++) {
        const int c = lane + 32 * j;
        if (c < dh) {
          float a = o[i][j] * corr;
#pragma unroll 8
          for (int k = 0; k < TILE; k++) a = fmaf(Ps[warp * TILE + k], Vs[k * ldk + c], a);
          o[i][j] = a;
        }
      }
      __syncwarp();
    }
  }
#pragma unroll
  for (int i = 0; i < RPW; i++) {
    const int64_t q = q0 + warp * RPW + i;
    if (q >= dm.T) break;
    const float inv = l[i] > 0.f ? 1.f / l[i] : 0.f;
#pragma unroll
    for (int j = 0; j < NJMAX; j++) {
      const int c = lane + 32 * j;
      if (c < dh) out[(b * dm.T + q) * d + h * dm.dh + c] = from_f32<T>(o[i][j] * inv);
    }
    if (lane == 0) lse[bh * dm.T + q] = l[i] > 0.f ? m[i] + logf(l[i]) : -INFINITY;
  }
}

// ------------------------------------------------------------------------------------------
// delta[b,h,q] = Σ_c dO·O : one warp per (b,q,h)
template <typename T>
__global__ void attn_delta_kernel(const T* __restrict__ out, const T* __restrict__ dout, float* __restrict__ delta, AttnDims dm) {
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  if (w >= dm.B * dm.T * dm.H) return;
  const int64_t h = w % dm.H, bq = w / dm.H, b = bq / dm.T, q = bq % dm.T;
  const int64_t base = bq * dm.H * dm.dh + h * dm.dh;
  float s = 0.f;
  for (int c = lane; c < dm.dh; c += 32) s += to_f32<T>(out[base + c]) * to_f32<T>(dout[base + c]);
  s = warp_sum(s);
  if (lane == 0) delta[(b * dm.H + h) * dm.T + q] = s;
}

// ------------------------------------------------------------------------------------------
// dQ: CTA owns 32 queries, streams key tiles.
template <typename T>
__global__ void __launch_bounds__(128)
attn_bwd_dq_simt_kernel(const T* __restrict__ qkv, const uint8_t* __restrict__ key_mask, const T* __restrict__ dout,
                        const float* __restrict__ lse, const float* __restrict__ delta, T* __restrict__ dqkv,
                        AttnDims dm, float p_drop, const uint32_t* __restrict__ dbits) {
  extern __shared__ float sm[];
  const int dh = (int)dm.dh, ldk = dh + 1;
  float* Ks = sm;
  float* Vs = Ks + TILE * ldk;
  float* Qs = Vs + TILE * ldk;     // scaled queries
  float* dOs = Qs + ROWS * dh;
  float* Ps = dOs + ROWS * dh;
  const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
  const int64_t bh = blockIdx.y, b = bh / dm.H, h = bh % dm.H;
  const int64_t d = dm.H * dm.dh, rs = 3 * d;
  const int64_t q0 = (int64_t)blockIdx.x * ROWS;
  const T* qbase = qkv + b * dm.T * rs + h * dm.dh;
  const float scale = rsqrtf((float)dh);
  const bool drop = p_drop > 0.f;
  const DropBits dk = make_drop_bits(dbits, dm.T, p_drop);

  load_rows_to_smem<T>(Qs, dh, qbase, rs, q0, dm.T, dh, scale);
  load_rows_to_smem<T>(dOs, dh, dout + b * dm.T * d + h * dm.dh, d, q0, dm.T, dh, 1.f);

  float acc[RPW][NJMAX], ls[RPW], dl[RPW];
#pragma unroll
  for (int i = 0; i < RPW; i++) {
    const int64_t q = q0 + warp * RPW + i;
    ls[i] = q < dm.T ? lse[bh * dm.T + q] : -INFINITY;
    dl[i] = q < dm.T ? delta[bh * dm.T + q] : 0.f;
#pragma unroll
    for (int j = 0; j < NJMAX; j++) acc[i][j] = 0.f;
  }

  for (int64_t k0 = 0; k0 < dm.T; k0 += TILE) {
    __syncthreads();
    load_rows_to_smem<T>(Ks, ldk, qbase + d, rs, k0, dm.T, dh, 1.f);
    load_rows_to_smem<T>(Vs, ldk, qbase + 2 * d, rs, k0, dm.T, dh, 1.f);
    __syncthreads();
    const int64_t kk = k0 + lane;
    const bool valid = kk < dm.T && !(key_mask != nullptr && key_mask[b * dm.T + kk]);
#pragma unroll
    for (int i = 0; i < RPW; i++) {
      const int qi = warp * RPW + i;
      const int64_t q = q0 + qi;
      if (q >= dm.T) break;
      float s = 0.f, dp = 0.f;
      for (int c = 0; c < dh; c++) {
        s = fmaf(Qs[qi * dh + c], Ks[lane * ldk + c], s);
        dp = fmaf(dOs[qi * dh + c], Vs[lane * ldk + c], dp);
      }
      float p = (valid && ls[i] != -INFINITY) ? expf(s - ls[i]) : 0.f;
      if (drop) dp = dropbit(dk, (int64_t)bh * dm.T + q, (int)kk) ? dp * dk.scale : 0.f;
      const float ds = p * (dp - dl[i]);
      Ps[warp * TILE + lane] = ds;
      __syncwarp();
#pragma unroll
      for (int j = 0; j < NJMAX; j++) {
        const int c = lane + 32 * j;
        if (c < dh) {
          float a = acc[i][j];
#pragma unroll 8
          for (int k = 0; k < TILE; k++) a = fmaf(Ps[warp * TILE + k], Ks[k * ldk + c], a);
          acc[i][j] = a;
        }
      }
      __syncwarp();
    }
  }
#pragma unroll
  for (int i = 0; i < RPW; i++) {
    const int64_t q = q0 + warp * RPW + i;
    if (q >= dm.T) break;
#pragma unroll
    for (int j = 0; j < NJMAX; j++) {
      const int c = lane + 32 * j;
      if (c < dh) dqkv[(b * dm.T + q) * rs + h * dm.dh + c] = from_f32<T>(acc[i][j] * scale);
    }
  }
}

// ------------------------------------------------------------------------------------------
// dK, dV: CTA owns 32 keys, streams query tiles.
template <typename T>
__global__ void __launch_bounds__(128)
attn_bwd_dkv_simt_kernel(const T* __restrict__ qkv, const uint8_t* __restrict__ key_mask, const T* __restrict__ dout,
                         const float* __restrict__ lse, const float* __restrict__ delta, T* __restrict__ dqkv,
                         AttnDims dm, float p_drop, const uint32_t* __restrict__ dbits) {
  extern __shared__ float sm[];
  const int dh = (int)dm.dh, ldq = dh + 1;
  float* Qs = sm;                   // [TILE][dh+1] scaled queries
  float* dOs = Qs + TILE * ldq;     // [TILE][dh+1]
  float* Ks = dOs + TILE * ldq;     // [ROWS][dh]
  float* Vs = Ks + ROWS * dh;       // [ROWS][dh]
  float* Ps = Vs + ROWS * dh;       // [4][32] p~
  float* Ds = Ps + 4 * TILE;        // [4][32] ds
  float* Ls = Ds + 4 * TILE;        // [32] lse
  float* Dl = Ls + TILE;            // [32] delta
  const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
  const int64_t bh = blockIdx.y, b = bh / dm.H, h = bh % dm.H;
  const int64_t d = dm.H * dm.dh, rs = 3 * d;
  const int64_t k0 = (int64_t)blockIdx.x * ROWS;
  const T* qbase = qkv + b * dm.T * rs + h * dm.dh;
  const float scale = rsqrtf((float)dh);
  const bool drop = p_drop > 0.f;
  const DropBits dk = make_drop_bits(dbits, dm.T, p_drop);

  load_rows_to_smem<T>(Ks, dh, qbase + d, rs, k0, dm.T, dh, 1.f);
  load_rows_to_smem<T>(Vs, dh, qbase + 2 * d, rs, k0, dm.T, dh, 1.f);

  float dK[RPW][NJMAX], dV[RPW][NJMAX];
#pragma unroll
  for (int i = 0; i < RPW; i++)
#pragma unroll
    for (int j = 0; j < NJMAX; j++) { dK[i][j] = 0.f; dV[i][j] = 0.f; }

  for (int64_t q0 = 0; q0 < dm.T; q0 += TILE) {
    __syncthreads();
    load_rows_to_smem<T>(Qs, ldq, qbase, rs, q0, dm.T, dh, scale);
    load_rows_to_smem<T>(dOs, ldq, dout + b * dm.T * d + h * dm.dh, d, q0, dm.T, dh, 1.f);
    if (threadIdx.x < TILE) {
      int64_t q = q0 + threadIdx.x;
      Ls[threadIdx.x] = q < dm.T ? lse[bh * dm.T + q] : -INFINITY;
      Dl[threadIdx.x] = q < dm.T ? delta[bh * dm.T + q] : 0.f;
    }
    __syncthreads();
    const int64_t q = q0 + lane;
    const float lq = Ls[lane], dq = Dl[lane];
    const bool qvalid = q < dm.T && lq != -INFINITY;
#pragma unroll
    for (int i = 0; i < RPW; i++) {
      const int ki = warp * RPW + i;
      const int64_t kk = k0 + ki;
      if (kk >= dm.T) break;
      const bool kvalid = !(key_mask != nullptr && key_mask[b * dm.T + kk]);   // warp-uniform
      if (!kvalid) continue;
      float s = 0.f, dp = 0.f;
      for (int c = 0; c < dh; c++) {
        s = fmaf(Qs[lane * ldq + c], Ks[ki * dh + c], s);
        dp = fmaf(dOs[lane * ldq + c], Vs[ki * dh + c], dp);
      }
      float p = qvalid ? expf(s - lq) : 0.f;
      float pd = p;
      if (drop) {
        const bool keep = dropbit(dk, (int64_t)bh * dm.T + q, (int)kk);
        pd = keep ? p * dk.scale : 0.f;
        dp = keep ? dp * dk.scale : 0.f;
      }
      const float ds = p * (dp - dq);
      Ps[warp * TILE + lane] = pd;
      Ds[warp * TILE + lane] = ds;
      __syncwarp();
#pragma unroll
      for (int j = 0; j < NJMAX; j++) {
        const int c = lane + 32 * j;
        if (c < dh) {
          float av = dV[i][j], ak = dK[i][j];
#pragma unroll 8
          for (int r = 0; r < TILE; r++) {
            av = fmaf(Ps[warp * TILE + r], dOs[r * ldq + c], av);
            ak = fmaf(Ds[warp * TILE + r], Qs[r * ldq + c], ak);   // Qs carries 1/sqrt(dh)
          }
          dV[i][j] = av; dK[i][j] = ak;
        }
      }
      __syncwarp();
    }
  }
#pragma unroll
  for (int i = 0; i < RPW; i++) {
    const int64_t kk = k0 + warp * RPW + i;
    if (kk >= dm.T) break;
#pragma unroll
    for (int j = 0; j < NJMAX; j++) {
      const int c = lane + 32 * j;
      if (c < dh) {
        dqkv[(b * dm.T + kk) * rs + d + h * dm.dh + c] = from_f32<T>(dK[i][j]);
        dqkv[(b * dm.T + kk) * rs + 2 * d + h * dm.dh + c] = from_f32<T>(dV[i][j]);
      }
    }
  }
}

template <typename T>
int fwd_impl(const void* qkv, const uint8_t* key_mask, void* out, float* lse, int64_t B, int64_t Tn, int64_t H,
             int64_t dh, float p, const uint32_t* dbits, cudaStream_t st) {
  AttnDims dm{B, Tn, H, dh};
  size_t smem = (size_t)(2 * TILE * (dh + 1) + ROWS * dh + 4 * TILE) * sizeof(float);
  static size_t cfg_fwd = 0;
  if (smem > cfg_fwd) { cudaFuncSetAttribute(attn_fwd_simt_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); cfg_fwd = smem; }
  dim3 grid((unsigned)ceil_div(Tn, ROWS), (unsigned)(B * H));
  attn_fwd_simt_kernel<T><<<grid, 128, smem, st>>>((const T*)qkv, key_mask, (T*)out, lse, dm, p, dbits);
  MAR_LAUNCH_CHECK("attn_fwd_simt");
  return MAR_OK;
}

template <typename T>
int bwd_impl(const void* qkv, const uint8_t* key_mask, const void* out, const void* dout, const float* lse,
             float* delta, void* dqkv, int64_t B, int64_t Tn, int64_t H, int64_t dh, float p, const uint32_t* dbits,
             cudaStream_t st) {
  AttnDims dm{B, Tn, H, dh};
  {
    int64_t warps = B * Tn * H;
    attn_delta_kernel<T><<<(unsigned)ceil_div(warps, 8), 256, 0, st>>>((const T*)out, (const T*)dout, delta, dm);
    MAR_LAUNCH_CHECK("attn_delta");
  }
  dim3 grid((unsigned)ceil_div(Tn, ROWS), (unsigned)(B * H));
  {
    size_t smem = (size_t)(2 * TILE * (dh + 1) + 2 * ROWS * dh + 4 * TILE) * sizeof(float);
    static size_t cfg_dq = 0;
    if (smem > cfg_dq) { cudaFuncSetAttribute(attn_bwd_dq_simt_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); cfg_dq = smem; }
    attn_bwd_dq_simt_kernel<T><<<grid, 128, smem, st>>>((const T*)qkv, key_mask, (const T*)dout, lse, delta, (T*)dqkv, dm,
                                                        p, dbits);
    MAR_LAUNCH_CHECK("attn_bwd_dq_simt");
  }
  {
    size_t smem = (size_t)(2 * TILE * (dh + 1) + 2 * ROWS * dh + 8 * TILE + 2 * TILE) * sizeof(float);
    static size_t cfg_dkv = 0;
    if (smem > cfg_dkv) { cudaFuncSetAttribute(attn_bwd_dkv_simt_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); cfg_dkv = smem; }
    attn_bwd_dkv_simt_kernel<T><<<grid, 128, smem, st>>>((const T*)qkv, key_mask, (const T*)dout, lse, delta, (T*)dqkv, dm,
                                                         p, dbits);
    MAR_LAUNCH_CHECK("attn_bwd_dkv_simt");
  }
  return MAR_OK;
}

}  // namespace

int attention_fwd_simt(const void* qkv, const uint8_t* key_mask, void* out, float* lse, int64_t B, int64_t T, int64_t H,
                       int64_t dh, int dtype, float p, const uint32_t* dbits, cudaStream_t st) {
  if (dh > 32 * NJMAX) MAR_UNSUPPORTED("attention (SIMT): head dim %lld > 128", (long long)dh);
  if (dtype == MAR_BF16) return fwd_impl<bf16>(qkv, key_mask, out, lse, B, T, H, dh, p, dbits, st);
  if (dtype == MAR_F32) return fwd_impl<float>(qkv, key_mask, out, lse, B, T, H, dh, p, dbits, st);
  MAR_UNSUPPORTED("attention: dtype %d", dtype);
}

int attention_bwd_simt(const void* qkv, const uint8_t* key_mask, const void* out, const void* dout, const float* lse,
                       float* delta, void* dqkv, int64_t B, int64_t T, int64_t H, int64_t dh, int dtype, float p,
                       const uint32_t* dbits, cudaStream_t st) {
  if (dh > 32 * NJMAX) MAR_UNSUPPORTED("attention (SIMT): head dim %lld > 128", (long long)dh);
  if (dtype == MAR_BF16) return bwd_impl<bf16>(qkv, key_mask, out, dout, lse, delta, dqkv, B, T, H, dh, p, dbits, st);
  if (dtype == MAR_F32) return bwd_impl<float>(qkv, key_mask, out, dout, lse, delta, dqkv, B, T, H, dh, p, dbits, st);
  MAR_UNSUPPORTED("attention: dtype %d", dtype);
}
