// LayerNorm forward / backward over the last dimension (HBM-bound).
// One warp per row; a lane owns NCH chunks of 8 consecutive elements (16 B bf16 / 32 B fp32 loads,
// a warp covers 256 contiguous elements per chunk round => fully coalesced).  Statistics in fp32.
// Algorithmic bytes: fwd = rows*D*(read+write)*sizeof(T); bwd = rows*D*(2 reads + 1 write)*sizeof(T).
#define MAR_PDL_CLASS 2
#include "common.cuh"

namespace {

inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Where a row of the (rows, D) problem lives on the "mapped" side (the forward's y, the backward's dy): rows are
// (b, t) = (row / Tin, row % Tin); segment k covers t in [t0, t1) and places the row at
//   ptr[k] + ((b * Tout[k]) + tout0[k] + (t - t0[k])) * D.
// nseg = 0: the plain contiguous (rows, D) layout.  This is how an encoder's final LayerNorm writes straight into its
// slice of the fused (B, T_a+T_v, d) sequence (one segment, Tout = T_a+T_v) and how the fusion encoder's final LayerNorm
// writes one contiguous tensor per modality (one segment each) — torch.cat / the per-modality slices of
// models.py:419,430 without a copy; the backward reads its incoming gradient through the same map.
struct RowMap {
  int nseg, Tin;
  int t0[4], t1[4], Tout[4], tout0[4];
  void* ptr[4];
};

template <typename T>
__device__ __forceinline__ T* mapped_row(const RowMap& m, T* plain, int64_t row, int D) {
  if (m.nseg == 0) return plain + row * D;
  const int64_t b = row / m.Tin;
  const int t = (int)(row - b * m.Tin);
  // constant indices only (a dynamic index into a by-value kernel parameter puts the struct on the local-memory stack)
  int t0 = m.t0[0], Tout = m.Tout[0], tout0 = m.tout0[0];
  void* ptr = m.ptr[0];
#pragma unroll
  for (int i = 1; i < 4; i++)
    if (i < m.nseg && t >= m.t0[i]) { t0 = m.t0[i]; Tout = m.Tout[i]; tout0 = m.tout0[i]; ptr = m.ptr[i]; }
  return reinterpret_cast<T*>(ptr) + (b * Tout + tout0 + (t - t0)) * (int64_t)D;
}

template <typename T, int NCH>
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const T* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     T* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd,
                     const uint8_t* __restrict__ zero_rows, int64_t rows, int D, float eps, const RowMap map) {
  pdl_entry();
  const int lane = threadIdx.x % 32;
  const int warps_per_block = blockDim.x / 32;
  const int64_t warp_global = (int64_t)blockIdx.x * warps_per_block + threadIdx.x / 32;
  const int64_t warp_stride = (int64_t)gridDim.x * warps_per_block;
  const float invD = 1.f / (float)D;

  for (int64_t row = warp_global; row < rows; row += warp_stride) {
    float v[NCH][8];
    const bool zero = zero_rows != nullptr && zero_rows[row] != 0;
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; c++) {
      int col = (c * 32 + lane) * 8;
      if (col < D && !zero) Vec8<T>::load(x + row * D + col, v[c]);
      else {
#pragma unroll
        for (int j = 0; j < 8; j++) v[c][j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; j++) s += v[c][j];
    }
    const float mu = warp_sum(s) * invD;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; c++) {
      int col = (c * 32 + lane) * 8;
      if (col < D) {
#pragma unroll
        for (int j = 0; j < 8; j++) { float d = v[c][j] - mu; q += d * d; }
      }
    }
    const float rs = rsqrtf(warp_sum(q) * invD + eps);
    if (lane == 0 && mean != nullptr) { mean[row] = mu; rstd[row] = rs; }
    T* yrow = mapped_row<T>(map, y, row, D);
#pragma unroll
    for (int c = 0; c < NCH; c++) {
      int col = (c * 32 + lane) * 8;
      if (col < D) {
        float g[8], b[8], o[8];
        Vec8<float>::load(gamma + col, g);
        Vec8<float>::load(beta + col, b);
#pragma unroll
        for (int j = 0; j < 8; j++) o[j] = (v[c][j] - mu) * rs * g[j] + b[j];
        Vec8<T>::store(yrow + col, o);
      }
    }
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g*xhat)),  g = dy*gamma ; dgamma += Σ dy*xhat ; dbeta += Σ dy
// Register diet (ncu: the first version held gamma, g and xhat in registers -> 158 regs, ONE 8-warp block per
// SM, 24 KB of loads in flight per SM and 1/3 of HBM speed): the row is kept as the raw 16 B vectors it was loaded
// as and decoded twice, gamma is re-read through L1 (3 KB, hot) -> <= 128 regs, 2 blocks = 16 warps per SM.
template <typename T> struct Raw8;
template <> struct Raw8<bf16> {
  uint4 u;
  __device__ __forceinline__ void load(const bf16* p) { u = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void get(float (&v)[8]) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; i++) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
};
template <> struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = *reinterpret_cast<const float4*>(p); b = *reinterpret_cast<const float4*>(p + 4);
  }
  __device__ __forceinline__ void get(float (&v)[8]) const {
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};

// DROP: x is the output of a linear whose epilogue was  x = residual + dropout(z)  (out_proj / linear2 of the encoder
// layer, transformer.py:953-956): the kernel ALSO writes dz = dx ⊙ keep / (1 - p) — the gradient that linear's dgrad and
// wgrad GEMMs consume — and accumulates that linear's bias gradient (column sums of dz), so the separate dropout-backward
// pass over the tensor (one read + one write) disappears.  The mask is regenerated from (rng, site, element index) like
// everywhere else (common.cuh).  dz's column sums live in per-warp shared-memory rows (a lane owns its columns: plain
// read-modify-write, no atomics) — a third set of register accumulators would cost the second resident block.
struct LnDrop {
  void* dz;
  float* dbias;
  const uint64_t* rng;
  uint32_t site;
  float p;
};

template <typename T, int NCH, bool DROP>
__global__ void __launch_bounds__(256, 2)
layernorm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ mean,
                     const float* __restrict__ rstd, const float* __restrict__ gamma, T* __restrict__ dx,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t rows, int D, const RowMap map,
                     const LnDrop drp) {
  pdl_entry();
  const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
  const int warps_per_block = blockDim.x / 32;
  const int64_t warp_global = (int64_t)blockIdx.x * warps_per_block + warp;
  const int64_t warp_stride = (int64_t)gridDim.x * warps_per_block;
  const float invD = 1.f / (float)D;
  extern __shared__ __align__(16) float red[];  // [warps][D] reused for dgamma then dbeta ; DROP: + [warps][NCH][2][32] float4 of dz sums
  float4* zacc = reinterpret_cast<float4*>(red + warps_per_block * D) + warp * NCH * 64;
  DropKey dk;
  T* dzp = nullptr;
  [[maybe_unused]] const bool idx32 = ((uint64_t)rows * (uint64_t)D >> 1) < 0xffffffffull;   // every pair index fits 32 bits
  if (DROP) {
    dk = make_drop_key(drp.rng, drp.site, drp.p);
    dzp = reinterpret_cast<T*>(drp.dz);
#pragma unroll
    for (int i = 0; i < NCH * 2; i++) zacc[i * 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
  }

  float dg[NCH][8], db[NCH][8];
#pragma unroll
  for (int c = 0; c < NCH; c++) {
#pragma unroll
    for (int j = 0; j < 8; j++) { dg[c][j] = 0.f; db[c][j] = 0.f; }
  }

  for (int64_t row = warp_global; row < rows; row += warp_stride) {
    Raw8<T> ra[NCH], rb[NCH];
    const T* dyrow = mapped_row<const T>(map, dy, row, D);
#pragma unroll
    for (int c = 0; c < NCH; c++) {
      const int col = (c * 32 + lane) * 8;
      if (col < D) { ra[c].load(dyrow + col); rb[c].load(x + row * D + col); }
    }
    const float mu = mean[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; c++) {
      const int col = (c * 32 + lane) * 8;
      if (col < D) {
        float a[8], b[8], gm[8];
        ra[c].get(a); rb[c].get(b);
        Vec8<float>::load(gamma + col, gm);
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const float xh = (b[j] - mu) * rs;
          const float g = a[j] * gm[j];
          dg[c][j] += a[j] * xh;
          db[c][j] += a[j];
          s1 += g;
          s2 += g * xh;
        }
      }
    }
    s1 = warp_sum(s1) * invD;
    s2 = warp_sum(s2) * invD;
#pragma unroll
    for (int c = 0; c < NCH; c++) {
      const int col = (c * 32 + lane) * 8;
      if (col < D) {
        float a[8], b[8], gm[8], o[8];
        ra[c].get(a); rb[c].get(b);
        Vec8<float>::load(gamma + col, gm);
#pragma unroll
        for (int j = 0; j < 8; j++) o[j] = rs * (a[j] * gm[j] - s1 - (b[j] - mu) * rs * s2);
        Vec8<T>::store(dx + row * D + col, o);
        if (DROP) {
          const uint64_t e0 = (uint64_t)row * (uint64_t)D + (uint64_t)col;     // even: one hash per pair of elements
#pragma unroll
          for (int j = 0; j < 4; j++) {
            bool k0, k1;
            if (idx32) drop_keep2_32(dk, (uint32_t)(e0 >> 1) + (uint32_t)j, k0, k1);
            else drop_keep2(dk, (e0 >> 1) + j, k0, k1);
            o[2 * j] = k0 ? o[2 * j] * dk.scale : 0.f;
            o[2 * j + 1] = k1 ? o[2 * j + 1] * dk.scale : 0.f;
          }
          Vec8<T>::store(dzp + row * D + col, o);
          float4 z0 = zacc[(c * 2) * 32 + lane], z1 = zacc[(c * 2 + 1) * 32 + lane];
          z0.x += o[0]; z0.y += o[1]; z0.z += o[2]; z0.w += o[3];
          z1.x += o[4]; z1.y += o[5]; z1.z += o[6]; z1.w += o[7];
          zacc[(c * 2) * 32 + lane] = z0; zacc[(c * 2 + 1) * 32 + lane] = z1;
        }
      }
    }
  }

  // block reduction of the column partials: 8 warps -> 1, then one vector atomicAdd per 4 columns per block
  for (int pass = 0; pass < 2; pass++) {
#pragma unroll
    for (int c = 0; c < NCH; c++) {
      int col = (c * 32 + lane) * 8;
      if (col < D) {
#pragma unroll
        for (int j = 0; j < 8; j++) red[warp * D + col + j] = pass == 0 ? dg[c][j] : db[c][j];
      }
    }
    __syncthreads();
    for (int col = threadIdx.x * 4; col < D; col += blockDim.x * 4) {     // D % 8 == 0: 16 B vector atomics
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int w = 0; w < warps_per_block; w++) {
        const float4 t = *reinterpret_cast<const float4*>(&red[w * D + col]);
        s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
      }
      atomicAdd(reinterpret_cast<float4*>((pass == 0 ? dgamma : dbeta) + col), s);
    }
    __syncthreads();
  }
  if (DROP && drp.dbias != nullptr) {                                     // (the loop above ended with a __syncthreads)
    const float4* zall = reinterpret_cast<const float4*>(red + warps_per_block * D);
    for (int col = threadIdx.x * 4; col < D; col += blockDim.x * 4) {
      const int c = col / 256, l = (col % 256) / 8, half = (col % 8) / 4;
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int w = 0; w < warps_per_block; w++) {
        const float4 t = zall[w * NCH * 64 + (c * 2 + half) * 32 + l];
        s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
      }
      atomicAdd(reinterpret_cast<float4*>(drp.dbias + col), s);
    }
  }
}

template <typename T>
int launch_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
               const uint8_t* zero_rows, int64_t rows, int64_t D, float eps, const RowMap& map, cudaStream_t st) {
  const int nch = (int)ceil_div(D, 256);
  int64_t blocks = ceil_div(rows, 8);
  int64_t cap = (int64_t)mar_sm_count() * 8;
  if (blocks > cap) blocks = cap;
#define LN_FWD(N)                                                                                             \
  mar_launch(layernorm_fwd_kernel<T, N>, (unsigned)blocks, 256, 0, st, (const T*)x, gamma, beta, (T*)y, mean, rstd,   \
                                                               zero_rows, rows, (int)D, eps, map)
  switch (nch) {
    case 1: LN_FWD(1); break;
    case 2: LN_FWD(2); break;
    case 3: LN_FWD(3); break;
    case 4: LN_FWD(4); break;
    case 5: LN_FWD(5); break;
    case 6: LN_FWD(6); break;
    case 7: LN_FWD(7); break;
    case 8: LN_FWD(8); break;
    default: MAR_UNSUPPORTED("layernorm: D=%lld > 2048 not supported", (long long)D);
  }
#undef LN_FWD
  MAR_LAUNCH_CHECK("layernorm_fwd");
  return MAR_OK;
}

template <typename T>
int launch_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma, void* dx,
               float* dgamma, float* dbeta, int64_t rows, int64_t D, const RowMap& map, const LnDrop& drp, cudaStream_t st) {
  const int nch = (int)ceil_div(D, 256);
  int64_t blocks = ceil_div(rows, 8 * 4);
  int64_t cap = (int64_t)mar_sm_count() * 2;    // one resident wave (2 blocks/SM): the column partials end in one atomic per block
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const bool drop = drp.dz != nullptr;
  size_t smem = (size_t)8 * D * sizeof(float) + (drop ? (size_t)8 * nch * 64 * sizeof(float4) : 0);
#define LN_BWD_T(N, DR)                                                                                           \
  do {                                                                                                            \
    static bool cfg = false;                                                                                      \
    if (smem > 48 * 1024 && !cfg) {                                                                               \
      cudaFuncSetAttribute(layernorm_bwd_kernel<T, N, DR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * (N * 256) * 4 * 2); \
      cfg = true;                                                                                                 \
    }                                                                                                             \
    mar_launch(layernorm_bwd_kernel<T, N, DR>, (unsigned)blocks, 256, smem, st, (const T*)dy, (const T*)x, mean, rstd, gamma, \
                                                                        (T*)dx, dgamma, dbeta, rows, (int)D, map, drp); \
  } while (0)
#define LN_BWD(N) do { if (drop) LN_BWD_T(N, true); else LN_BWD_T(N, false); } while (0)
  switch (nch) {
    case 1: LN_BWD(1); break;
    case 2: LN_BWD(2); break;
    case 3: LN_BWD(3); break;
    case 4: LN_BWD(4); break;
    case 5: LN_BWD(5); break;
    case 6: LN_BWD(6); break;
    case 7: LN_BWD(7); break;
    case 8: LN_BWD(8); break;
    default: MAR_UNSUPPORTED("layernorm: D=%lld > 2048 not supported", (long long)D);
  }
#undef LN_BWD
#undef LN_BWD_T
  MAR_LAUNCH_CHECK("layernorm_bwd");
  return MAR_OK;
}

}  // namespace

namespace {
int make_row_map(RowMap* m, int64_t rows, int nseg, int64_t Tin, const int64_t* t0, const int64_t* t1, const int64_t* Tout,
                 const int64_t* tout0, void* const* ptrs, const char* who) {
  m->nseg = 0; m->Tin = 1;
  if (nseg == 0) return MAR_OK;
  if (nseg < 0 || nseg > 4 || Tin <= 0 || Tin >= (1ll << 31) || rows % Tin != 0 || !t0 || !t1 || !Tout || !tout0 || !ptrs) {
    mar_set_error("%s: bad row map (1..4 segments, rows a multiple of Tin)", who);
    return MAR_ERR_INVALID;
  }
  int64_t expect = 0;
  for (int k = 0; k < nseg; k++) {
    if (t0[k] != expect || t1[k] <= t0[k] || !ptrs[k] || ((uintptr_t)ptrs[k] % 16) != 0 || Tout[k] < tout0[k] + (t1[k] - t0[k]) ||
        Tout[k] >= (1ll << 31)) {
      mar_set_error("%s: row-map segment %d is not contiguous with its predecessor, empty, unaligned or does not fit its target", who, k);
      return MAR_ERR_INVALID;
    }
    m->t0[k] = (int)t0[k]; m->t1[k] = (int)t1[k]; m->Tout[k] = (int)Tout[k]; m->tout0[k] = (int)tout0[k]; m->ptr[k] = ptrs[k];
    expect = t1[k];
  }
  if (expect != Tin) { mar_set_error("%s: row-map segments must cover [0, Tin)", who); return MAR_ERR_INVALID; }
  m->nseg = nseg; m->Tin = (int)Tin;
  return MAR_OK;
}
}  // namespace

extern "C" {

int mar_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                      const uint8_t* zero_rows, int64_t rows, int64_t D, float eps, int dtype, void* stream) {
  return mar_layernorm_fwd_mapped(x, gamma, beta, y, mean, rstd, zero_rows, rows, D, eps, dtype, 0, 0, nullptr, nullptr,
                                  nullptr, nullptr, nullptr, stream);
}

int mar_layernorm_fwd_mapped(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                             const uint8_t* zero_rows, int64_t rows, int64_t D, float eps, int dtype, int nseg, int64_t Tin,
                             const int64_t* t0, const int64_t* t1, const int64_t* Tout, const int64_t* tout0,
                             void* const* y_ptrs, void* stream) {
  MAR_CHECK_ARG(x && gamma && beta && (y || nseg > 0) && rows >= 0 && D > 0, "mar_layernorm_fwd: bad arguments");
  MAR_CHECK_ARG((mean == nullptr) == (rstd == nullptr), "mar_layernorm_fwd: mean and rstd go together");
  MAR_CHECK_ARG(D % 8 == 0, "mar_layernorm_fwd: D must be a multiple of 8 (got %lld)", (long long)D);
  if (rows == 0) return MAR_OK;
  RowMap map;
  int rc = make_row_map(&map, rows, nseg, Tin, t0, t1, Tout, tout0, y_ptrs, "mar_layernorm_fwd_mapped");
  if (rc) return rc;
  if (dtype == MAR_BF16) return launch_fwd<bf16>(x, gamma, beta, y, mean, rstd, zero_rows, rows, D, eps, map, S(stream));
  if (dtype == MAR_F32) return launch_fwd<float>(x, gamma, beta, y, mean, rstd, zero_rows, rows, D, eps, map, S(stream));
  MAR_UNSUPPORTED("mar_layernorm_fwd: dtype %d", dtype);
}

int mar_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                      void* dx, float* dgamma, float* dbeta, int64_t rows, int64_t D, int dtype, void* stream) {
  return mar_layernorm_bwd_mapped(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, rows, D, dtype, 0, 0, nullptr, nullptr, nullptr,
                                  nullptr, nullptr, stream);
}

int mar_layernorm_bwd_mapped(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                             void* dx, float* dgamma, float* dbeta, int64_t rows, int64_t D, int dtype, int nseg, int64_t Tin,
                             const int64_t* t0, const int64_t* t1, const int64_t* Tout, const int64_t* tout0,
                             void* const* dy_ptrs, void* stream) {
  MAR_CHECK_ARG((dy || nseg > 0) && x && mean && rstd && gamma && dx && dgamma && dbeta && rows >= 0 && D > 0,
                "mar_layernorm_bwd: bad arguments");
  MAR_CHECK_ARG(D % 8 == 0, "mar_layernorm_bwd: D must be a multiple of 8 (got %lld)", (long long)D);
  if (rows == 0) return MAR_OK;
  RowMap map;
  int rc = make_row_map(&map, rows, nseg, Tin, t0, t1, Tout, tout0, dy_ptrs, "mar_layernorm_bwd_mapped");
  if (rc) return rc;
  LnDrop drp = {nullptr, nullptr, nullptr, 0u, 0.f};
  if (dtype == MAR_BF16) return launch_bwd<bf16>(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, rows, D, map, drp, S(stream));
  if (dtype == MAR_F32) return launch_bwd<float>(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, rows, D, map, drp, S(stream));
  MAR_UNSUPPORTED("mar_layernorm_bwd: dtype %d", dtype);
}

int mar_layernorm_bwd_dropout(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                              void* dx, float* dgamma, float* dbeta, void* dz, float* dbias, int64_t rows, int64_t D,
                              int dtype, float p_drop, const uint64_t* rng_state, uint32_t site, void* stream) {
  MAR_CHECK_ARG(dy && x && mean && rstd && gamma && dx && dgamma && dbeta && dz && rows >= 0 && D > 0,
                "mar_layernorm_bwd_dropout: bad arguments");
  MAR_CHECK_ARG(D % 8 == 0, "mar_layernorm_bwd_dropout: D must be a multiple of 8 (got %lld)", (long long)D);
  MAR_CHECK_ARG(p_drop > 0.f && p_drop < 1.f && rng_state, "mar_layernorm_bwd_dropout: needs 0 < p_drop < 1 and rng_state");
  MAR_CHECK_ARG(((uintptr_t)dz % 16) == 0 && ((uintptr_t)dbias % 16) == 0, "mar_layernorm_bwd_dropout: dz / dbias must be 16 B aligned");
  if (rows == 0) return MAR_OK;
  RowMap map;
  map.nseg = 0; map.Tin = 1;
  LnDrop drp = {dz, dbias, rng_state, site, p_drop};
  if (dtype == MAR_BF16) return launch_bwd<bf16>(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, rows, D, map, drp, S(stream));
  if (dtype == MAR_F32) return launch_bwd<float>(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, rows, D, map, drp, S(stream));
  MAR_UNSUPPORTED("mar_layernorm_bwd_dropout: dtype %d", dtype);
}

}  // extern "C"
