// HBM-bound elementwise and reduction kernels: epilogue backward (+bias grad), casts, pooling,
// zero-row key mask, concat copies, classifier loss, Adam, RNG state.  All vectorised 16 B per
// thread where alignment allows, coalesced along the contiguous dimension, fp32 math.
#include "common.cuh"

namespace {

inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ------------------------------------------------------------------------------------------
// RNG state
// ------------------------------------------------------------------------------------------
__global__ void rng_init_kernel(uint64_t* st, uint64_t seed, uint64_t step) { st[0] = seed; st[1] = step; }
__global__ void rng_advance_kernel(uint64_t* st) { st[1] += 1; }

// ------------------------------------------------------------------------------------------
// dz = dout * d(epilogue)/dz ; dbias += colsum(dz)
// block = (32 column-vectors) x (8 row lanes); each thread owns VEC consecutive columns.
// ------------------------------------------------------------------------------------------
template <typename T, typename TO, int VEC>
__global__ void __launch_bounds__(256)
bwd_epilogue_kernel(const T* __restrict__ dout, const TO* __restrict__ out, T* __restrict__ dz,
                    float* __restrict__ dbias, int64_t M, int64_t N, int rows_per_block, int flags, float p,
                    const uint64_t* __restrict__ rng, uint32_t site, int bcast) {
  pdl_entry();
  // bcast > 0: dout has M / bcast rows and row r reads dout row r / bcast scaled by 1 / bcast — the backward of a mean
  // over `bcast` consecutive rows (SequenceAverageFeatures after an adaptor, models.py:693-699) without materialising it
  const int lane = threadIdx.x, ry = threadIdx.y;
  const int64_t c0 = ((int64_t)blockIdx.x * 32 + lane) * VEC;
  const int64_t r_begin = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r_end = min(M, r_begin + rows_per_block);
  const bool relu = flags & (MAR_EPI_RELU_PRE | MAR_EPI_RELU_POST);
  const bool drop = (flags & MAR_EPI_DROPOUT) && p > 0.f;
  DropKey dk;
  if (drop) dk = make_drop_key(rng, site, p);
  const float scale = (drop ? dk.scale : 1.f) * (bcast > 0 ? 1.f / (float)bcast : 1.f);
  const bool idx32 = ((uint64_t)M * (uint64_t)N >> 1) < 0xffffffffull;     // every pair index fits 32 bits
  float csum[VEC];
#pragma unroll
  for (int j = 0; j < VEC; j++) csum[j] = 0.f;

  if (c0 < N) {
    if (VEC == 8) {
      // 4 rows per trip: all loads of the trip are issued before the first use (memory-level parallelism), the
      // dropout mask comes from ONE hash per PAIR of elements (drop_keep2), as in the forward epilogue
      constexpr int U = 4;
      for (int64_t r = r_begin + ry; r < r_end; r += 8 * U) {
        uint4 gq[U], oq[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
          const int64_t rr = r + 8 * u;
          if (rr < r_end) {
            const int64_t rs = bcast > 0 ? rr / bcast : rr;
            gq[u] = *reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(dout) + (rs * N + c0) * sizeof(T));
            if (relu) oq[u] = *reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(out) + (rr * N + c0) * sizeof(TO));
          }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
          const int64_t rr = r + 8 * u;
          if (rr >= r_end) break;
          const int64_t e0 = rr * N + c0;
          float g[8], o[8];
          if (sizeof(T) == 2) {
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&gq[u]);
#pragma unroll
            for (int j = 0; j < 4; j++) { const float2 f2 = __bfloat1622float2(h2[j]); g[2 * j] = f2.x; g[2 * j + 1] = f2.y; }
          } else {
            Vec8<T>::load(dout + (bcast > 0 ? (rr / bcast) * N + c0 : e0), g);
          }
          if (relu) {
            if (sizeof(TO) == 2) {
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&oq[u]);
#pragma unroll
              for (int j = 0; j < 4; j++) { const float2 f2 = __bfloat1622float2(h2[j]); o[2 * j] = f2.x; o[2 * j + 1] = f2.y; }
            } else {
              Vec8<TO>::load(out + e0, o);
            }
#pragma unroll
            for (int j = 0; j < 8; j++) g[j] = o[j] > 0.f ? g[j] * scale : 0.f;
          } else if (drop) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
              bool k0, k1;                                           // N % 8 == 0 and c0 % 8 == 0: e0 is even
              if (idx32) drop_keep2_32(dk, (uint32_t)(e0 >> 1) + (uint32_t)j, k0, k1);
              else drop_keep2(dk, (uint64_t)(e0 >> 1) + j, k0, k1);
              g[2 * j] = k0 ? g[2 * j] * scale : 0.f;
              g[2 * j + 1] = k1 ? g[2 * j + 1] * scale : 0.f;
            }
          } else if (bcast > 0) {
#pragma unroll
            for (int j = 0; j < 8; j++) g[j] *= scale;
          }
#pragma unroll
          for (int j = 0; j < 8; j++) csum[j % VEC] += g[j];
          if (dz != nullptr) Vec8<T>::store(dz + e0, g);
        }
      }
    } else {
      for (int64_t r = r_begin + ry; r < r_end; r += 8) {
        float g[VEC], o[VEC];
        const int64_t e0 = r * N + c0;
#pragma unroll
        for (int j = 0; j < VEC; j++) {
          bool ok = c0 + j < N;
          g[j] = ok ? to_f32<T>(dout[(bcast > 0 ? (r / bcast) * N + c0 : e0) + j]) : 0.f;
          o[j] = (ok && relu) ? to_f32<TO>(out[e0 + j]) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < VEC; j++) {
          float f;
          if (relu) f = o[j] > 0.f ? scale : 0.f;
          else if (drop) f = drop_keep(dk, (uint64_t)(e0 + j)) ? scale : 0.f;
          else f = scale;
          g[j] *= f;
          csum[j] += g[j];
        }
        if (dz != nullptr) {
#pragma unroll
          for (int j = 0; j < VEC; j++)
            if (c0 + j < N) dz[e0 + j] = from_f32<T>(g[j]);
        }
      }
    }
  }
  if (dbias == nullptr) return;
  __shared__ float red[8][32 * VEC + 1];
#pragma unroll
  for (int j = 0; j < VEC; j++) red[ry][lane * VEC + j] = csum[j];
  __syncthreads();
  for (int idx = ry * 32 + lane; idx < 32 * VEC; idx += 256) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) s += red[k][idx];
    int64_t c = (int64_t)blockIdx.x * 32 * VEC + idx;
    if (c < N) atomicAdd(dbias + c, s);
  }
}

template <typename T, typename TO>
int launch_bwd_epilogue(const void* dout, const void* out, void* dz, float* dbias, int64_t M, int64_t N, int flags,
                        float p, const uint64_t* rng, uint32_t site, int bcast, cudaStream_t st) {
  const bool vec = (N % 8 == 0);
  const int VECW = vec ? 8 : 1;
  int64_t gx = ceil_div(N, 32 * VECW);
  int64_t want_y = ceil_div(4 * (int64_t)mar_sm_count(), gx);
  int64_t rows_per_block = ceil_div(M, want_y);
  rows_per_block = ceil_div(rows_per_block, 8) * 8;
  if (rows_per_block < 8) rows_per_block = 8;
  int64_t gy = ceil_div(M, rows_per_block);
  dim3 grid((unsigned)gx, (unsigned)gy), block(32, 8);
  if (vec)
    mar_launch(bwd_epilogue_kernel<T, TO, 8>, grid, block, 0, st, (const T*)dout, (const TO*)out, (T*)dz, dbias, M, N,
                                                          (int)rows_per_block, flags, p, rng, site, bcast);
  else
    mar_launch(bwd_epilogue_kernel<T, TO, 1>, grid, block, 0, st, (const T*)dout, (const TO*)out, (T*)dz, dbias, M, N,
                                                          (int)rows_per_block, flags, p, rng, site, bcast);
  MAR_LAUNCH_CHECK("bwd_epilogue");
  return MAR_OK;
}

// ------------------------------------------------------------------------------------------
// casts
// ------------------------------------------------------------------------------------------
template <typename TS, typename TD>
__global__ void cast_kernel(const TS* __restrict__ src, TD* __restrict__ dst, int64_t n) {
  pdl_entry();
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    float v[8];
    Vec8<TS>::load(src + i, v);
    Vec8<TD>::store(dst + i, v);
  } else {
    for (; i < n; i++) dst[i] = from_f32<TD>(to_f32<TS>(src[i]));
  }
}

// fp32 (N,K) -> T (N,K) and optional transpose T (K,N); 32x32 smem tile transpose.
template <typename T>
__global__ void cast_weight_kernel(const float* __restrict__ src, T* __restrict__ w, T* __restrict__ wt, int N, int K) {
  __shared__ float tile[32][33];
  int k = blockIdx.x * 32 + threadIdx.x;
  for (int i = threadIdx.y; i < 32; i += 8) {
    int n = blockIdx.y * 32 + i;
    float v = (n < N && k < K) ? src[(int64_t)n * K + k] : 0.f;
    tile[i][threadIdx.x] = v;
    if (w != nullptr && n < N && k < K) w[(int64_t)n * K + k] = from_f32<T>(v);
  }
  if (wt == nullptr) return;
  __syncthreads();
  int n = blockIdx.y * 32 + threadIdx.x;
  for (int i = threadIdx.y; i < 32; i += 8) {
    int kk = blockIdx.x * 32 + i;
    if (n < N && kk < K) wt[(int64_t)kk * N + n] = from_f32<T>(tile[threadIdx.x][i]);
  }
}

// ------------------------------------------------------------------------------------------
// mean over T
// ------------------------------------------------------------------------------------------
// grid (D/ (32*8)) x B ; block (32, 8): thread owns 8 columns, the 8 row lanes stride over T.
template <typename T>
__global__ void __launch_bounds__(256)
meanpool_fwd_kernel(const T* __restrict__ x, T* __restrict__ out, int64_t Tn, int64_t D) {
  pdl_entry();
  const int64_t b = blockIdx.y;
  const int64_t c0 = ((int64_t)blockIdx.x * 32 + threadIdx.x) * 8;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c0 < D) {
    for (int64_t t = threadIdx.y; t < Tn; t += 8) {
      float v[8];
      Vec8<T>::load(x + (b * Tn + t) * D + c0, v);
#pragma unroll
      for (int j = 0; j < 8; j++) s[j] += v[j];
    }
  }
  __shared__ float red[8][32 * 8 + 1];
#pragma unroll
  for (int j = 0; j < 8; j++) red[threadIdx.y][threadIdx.x * 8 + j] = s[j];
  __syncthreads();
  if (threadIdx.y == 0 && c0 < D) {
    float o[8];
    const float inv = 1.f / (float)Tn;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < 8; k++) a += red[k][threadIdx.x * 8 + j];
      o[j] = a * inv;
    }
    Vec8<T>::store(out + b * D + c0, o);
  }
}

template <typename T>
__global__ void meanpool_bwd_kernel(const T* __restrict__ dout, T* __restrict__ dx, int64_t Tn, int64_t D, int64_t total8) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // index of an 8-vector in (B,T,D)
  if (i >= total8) return;
  int64_t d8 = D / 8;
  int64_t c = (i % d8) * 8;
  int64_t b = i / (d8 * Tn);
  float v[8];
  Vec8<T>::load(dout + b * D + c, v);
  const float inv = 1.f / (float)Tn;
#pragma unroll
  for (int j = 0; j < 8; j++) v[j] *= inv;
  Vec8<T>::store(dx + i * 8, v);
}

// ------------------------------------------------------------------------------------------
// zero-row mask: one warp per row
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void rowzero_kernel(const T* __restrict__ x, uint8_t* __restrict__ mask, int64_t rows, int64_t D) {
  pdl_entry();
  int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  if (row >= rows) return;
  const int lane = threadIdx.x % 32;
  float s = 0.f;
  for (int64_t c = lane * 8; c < D; c += 256) {
    float v[8];
    Vec8<T>::load(x + row * D + c, v);
#pragma unroll
    for (int j = 0; j < 8; j++) s += v[j];
  }
  s = warp_sum(s);
  if (lane == 0) mask[row] = (s == 0.f) ? 1 : 0;
}

// ------------------------------------------------------------------------------------------
// concat copies (B,T,D) <-> (B,T_total,D)[:, t_off:t_off+T]
// ------------------------------------------------------------------------------------------
__global__ void concat_rows_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int64_t Tn, int64_t T_total,
                                   int64_t t_off, int64_t row_vecs, int64_t total, int to_concat) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int64_t v = i % row_vecs;
  int64_t r = i / row_vecs;
  int64_t b = r / Tn, t = r % Tn;
  int64_t j = (b * T_total + t_off + t) * row_vecs + v;
  if (to_concat) dst[j] = src[i];
  else dst[i] = src[j];
}

// ------------------------------------------------------------------------------------------
// focal loss (adeelh/pytorch-multi-class-focal-loss, the criterion train_multimodal.py:494-510 fetches from
// torch.hub): per row  L = -alpha[y] · (1 - p_y)^gamma · log p_y ;  'mean' = plain mean over the rows whose
// label is not ignored (NOT the alpha-weighted mean of nn.CrossEntropyLoss).  Single block, fp32.
// dL/dz_k = alpha[y] · [ gamma (1-p_y)^(gamma-1) p_y log p_y - (1-p_y)^gamma ] · (delta_ky - p_k) / n
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
focal_loss_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, const float* __restrict__ alpha,
                  float gamma, float* __restrict__ loss, float* __restrict__ dlogits, int64_t B, int64_t C) {
  __shared__ float s_num[256], s_cnt[256];
  float num = 0.f, cnt = 0.f;
  for (int64_t b = threadIdx.x; b < B; b += blockDim.x) {
    const int64_t y = labels[b];
    if (y < 0 || y >= C) continue;
    const float* l = logits + b * C;
    float m = l[0];
    for (int64_t c = 1; c < C; c++) m = fmaxf(m, l[c]);
    float se = 0.f;
    for (int64_t c = 0; c < C; c++) se += expf(l[c] - m);
    const float log_pt = l[y] - m - logf(se);
    const float pt = expf(log_pt);
    const float a = alpha ? alpha[y] : 1.f;
    num += -a * powf(fmaxf(1.f - pt, 0.f), gamma) * log_pt;
    cnt += 1.f;
  }
  s_num[threadIdx.x] = num;
  s_cnt[threadIdx.x] = cnt;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s_num[threadIdx.x] += s_num[threadIdx.x + o];
      s_cnt[threadIdx.x] += s_cnt[threadIdx.x + o];
    }
    __syncthreads();
  }
  const float n = s_cnt[0];
  if (threadIdx.x == 0) loss[0] = n > 0.f ? s_num[0] / n : 0.f;
  if (dlogits == nullptr) return;
  const float invn = n > 0.f ? 1.f / n : 0.f;
  for (int64_t b = threadIdx.x; b < B; b += blockDim.x) {
    const float* l = logits + b * C;
    float* g = dlogits + b * C;
    const int64_t y = labels[b];
    if (y < 0 || y >= C) {
      for (int64_t c = 0; c < C; c++) g[c] = 0.f;
      continue;
    }
    float m = l[0];
    for (int64_t c = 1; c < C; c++) m = fmaxf(m, l[c]);
    float se = 0.f;
    for (int64_t c = 0; c < C; c++) se += expf(l[c] - m);
    const float log_pt = l[y] - m - logf(se);
    const float pt = expf(log_pt);
    const float om = fmaxf(1.f - pt, 0.f);
    const float a = (alpha ? alpha[y] : 1.f) * invn;
    // d/dp_y of -(1-p)^g log p, times p_y (the softmax Jacobian's common factor)
    const float base = gamma > 0.f ? gamma * powf(om, gamma - 1.f) * pt * log_pt - powf(om, gamma) : -1.f;
    for (int64_t c = 0; c < C; c++) g[c] = a * base * ((c == y ? 1.f : 0.f) - expf(l[c] - m) / se);
  }
}

// ------------------------------------------------------------------------------------------
// cross entropy: single block (B is a few hundred), fp32
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cross_entropy_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels,
                     const float* __restrict__ cw, float* __restrict__ loss, float* __restrict__ dlogits,
                     int64_t* __restrict__ preds, int64_t B, int64_t C) {
  pdl_entry();
  __shared__ float s_num[256], s_den[256];
  float num = 0.f, den = 0.f;
  for (int64_t b = threadIdx.x; b < B; b += blockDim.x) {
    const float* l = logits + b * C;
    float m = l[0];
    int64_t am = 0;
    for (int64_t c = 1; c < C; c++)
      if (l[c] > m) { m = l[c]; am = c; }
    if (preds) preds[b] = am;
    int64_t y = labels[b];
    if (y < 0 || y >= C) continue;
    float se = 0.f;
    for (int64_t c = 0; c < C; c++) se += expf(l[c] - m);
    float w = cw ? cw[y] : 1.f;
    num += w * (m + logf(se) - l[y]);
    den += w;
  }
  s_num[threadIdx.x] = num;
  s_den[threadIdx.x] = den;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s_num[threadIdx.x] += s_num[threadIdx.x + o];
      s_den[threadIdx.x] += s_den[threadIdx.x + o];
    }
    __syncthreads();
  }
  const float total_w = s_den[0];
  if (threadIdx.x == 0) loss[0] = total_w > 0.f ? s_num[0] / total_w : 0.f;
  if (dlogits == nullptr) return;
  const float invw = total_w > 0.f ? 1.f / total_w : 0.f;
  for (int64_t b = threadIdx.x; b < B; b += blockDim.x) {
    const float* l = logits + b * C;
    float* g = dlogits + b * C;
    int64_t y = labels[b];
    if (y < 0 || y >= C) {
      for (int64_t c = 0; c < C; c++) g[c] = 0.f;
      continue;
    }
    float m = l[0];
    for (int64_t c = 1; c < C; c++) m = fmaxf(m, l[c]);
    float se = 0.f;
    for (int64_t c = 0; c < C; c++) se += expf(l[c] - m);
    float w = (cw ? cw[y] : 1.f) * invw;
    for (int64_t c = 0; c < C; c++) g[c] = w * (expf(l[c] - m) / se - (c == y ? 1.f : 0.f));
  }
}

// ------------------------------------------------------------------------------------------
// Adam
// ------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, const float* __restrict__ step_dev, int64_t n, float lr, float b1,
                            float b2, float eps) {
  const float step = step_dev[0];
  const float bc1 = 1.f - powf(b1, step);
  const float bc2 = 1.f - powf(b2, step);
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 4 <= n) {
    float4 pp = *reinterpret_cast<float4*>(p + i), gg = *reinterpret_cast<const float4*>(g + i);
    float4 mm = *reinterpret_cast<float4*>(m + i), vv = *reinterpret_cast<float4*>(v + i);
    float* P = &pp.x; float* G = &gg.x; float* Mm = &mm.x; float* V = &vv.x;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      Mm[j] = b1 * Mm[j] + (1.f - b1) * G[j];
      V[j] = b2 * V[j] + (1.f - b2) * G[j] * G[j];
      P[j] -= step_size * Mm[j] / (sqrtf(V[j]) * inv_sqrt_bc2 + eps);
    }
    *reinterpret_cast<float4*>(p + i) = pp;
    *reinterpret_cast<float4*>(m + i) = mm;
    *reinterpret_cast<float4*>(v + i) = vv;
  } else {
    for (; i < n; i++) {
      float mi = b1 * m[i] + (1.f - b1) * g[i];
      float vi = b2 * v[i] + (1.f - b2) * g[i] * g[i];
      m[i] = mi; v[i] = vi;
      p[i] -= step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
    }
  }
}
__global__ void adam_tick_kernel(float* step_dev) { step_dev[0] += 1.f; }

// ---- per-parameter Adam (torch.optim.Adam skips a parameter whose .grad is None: no moment decay, no step count) ----
// One segment per parameter.  seg_active[s] > 0 ⇔ the parameter received a gradient this step.  The tick kernel
// advances the step count of the active segments and leaves their bias-correction pair in seg_coef.
__global__ void adam_seg_tick_kernel(float* __restrict__ seg_steps, const float* __restrict__ seg_active,
                                     float2* __restrict__ seg_coef, int nseg, float lr, float b1, float b2) {
  pdl_entry();
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseg) return;
  if (!(seg_active[s] > 0.f)) { seg_coef[s] = make_float2(0.f, 0.f); return; }
  const float step = seg_steps[s] + 1.f;
  seg_steps[s] = step;
  const float bc1 = 1.f - powf(b1, step), bc2 = 1.f - powf(b2, step);
  seg_coef[s] = make_float2(lr / bc1, rsqrtf(bc2));       // x = step size, y = 1/sqrt(bias correction 2); x == 0 ⇔ skip
}
// chunk_seg[i / chunk]: segment of the chunk (every segment starts on a chunk boundary), < 0 for padding.
__global__ void __launch_bounds__(256)
adam_seg_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                const int32_t* __restrict__ chunk_seg, const float2* __restrict__ seg_coef, int64_t n, int chunk_shift,
                float b1, float b2, float eps, bf16* __restrict__ mirror) {
  pdl_entry();
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  const int seg = chunk_seg[i >> chunk_shift];
  if (seg < 0) return;
  const float2 coef = seg_coef[seg];
  if (coef.x == 0.f) return;
  float4 pp = *reinterpret_cast<float4*>(p + i), gg = *reinterpret_cast<const float4*>(g + i);
  float4 mm = *reinterpret_cast<float4*>(m + i), vv = *reinterpret_cast<float4*>(v + i);
  float* P = &pp.x; float* G = &gg.x; float* Mm = &mm.x; float* V = &vv.x;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    Mm[j] = b1 * Mm[j] + (1.f - b1) * G[j];
    V[j] = b2 * V[j] + (1.f - b2) * G[j] * G[j];
    P[j] -= coef.x * Mm[j] / (sqrtf(V[j]) * coef.y + eps);
  }
  *reinterpret_cast<float4*>(p + i) = pp;
  *reinterpret_cast<float4*>(m + i) = mm;
  *reinterpret_cast<float4*>(v + i) = vv;
  if (mirror != nullptr)       // the bf16 compute copy of the parameters, refreshed where they change
    *reinterpret_cast<uint2*>(mirror + i) = make_uint2(pack_bf16x2(pp.x, pp.y), pack_bf16x2(pp.z, pp.w));
}

// norm[0] = Σ_{rows with label >= 0} class_weight[label] (1 per row without weights): the denominator of
// nn.CrossEntropyLoss(reduction='mean') — what a data-parallel rank needs from its peers to weigh its gradient.
__global__ void label_weight_sum_kernel(const int64_t* __restrict__ labels, const float* __restrict__ w,
                                        float* __restrict__ out, int64_t B, int64_t C) {
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < B; i += blockDim.x) {
    const int64_t y = labels[i];
    if (y >= 0 && y < C) acc += w ? w[y] : 1.f;
  }
  __shared__ float part[32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) out[0] = t;
  }
}

}  // namespace

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" {

int mar_rng_init(uint64_t* rng_state, uint64_t seed, uint64_t step, void* stream) {
  MAR_CHECK_ARG(rng_state, "mar_rng_init: null state");
  rng_init_kernel<<<1, 1, 0, S(stream)>>>(rng_state, seed, step);
  MAR_LAUNCH_CHECK("rng_init");
  return MAR_OK;
}
int mar_rng_advance(uint64_t* rng_state, void* stream) {
  MAR_CHECK_ARG(rng_state, "mar_rng_advance: null state");
  rng_advance_kernel<<<1, 1, 0, S(stream)>>>(rng_state);
  MAR_LAUNCH_CHECK("rng_advance");
  return MAR_OK;
}

int mar_linear_bwd_epilogue(const void* dout, const void* out, void* dz, float* dbias, int64_t M, int64_t N,
                            int dtype, int out_dtype, int flags, float p_drop, const uint64_t* rng_state,
                            uint32_t site, int64_t pooled_rows, void* stream) {
  MAR_CHECK_ARG(dout && M >= 0 && N > 0, "mar_linear_bwd_epilogue: bad arguments");
  if (M == 0) return MAR_OK;
  const bool relu = flags & (MAR_EPI_RELU_PRE | MAR_EPI_RELU_POST);
  MAR_CHECK_ARG(!relu || out, "mar_linear_bwd_epilogue: ReLU backward needs the forward output");
  MAR_CHECK_ARG(!((flags & MAR_EPI_DROPOUT) && p_drop > 0.f) || rng_state, "mar_linear_bwd_epilogue: dropout needs rng_state");
  MAR_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, "mar_linear_bwd_epilogue: p_drop out of range");
  MAR_CHECK_ARG(pooled_rows >= 0 && pooled_rows < (1ll << 31) && (pooled_rows == 0 || (M % pooled_rows == 0 && dz)),
                "mar_linear_bwd_epilogue: pooled_rows must divide M (and dz must be given)");
  const int bcast = (int)pooled_rows;
  cudaStream_t st = S(stream);
  if (dtype == MAR_F32 && out_dtype == MAR_F32)
    return launch_bwd_epilogue<float, float>(dout, out, dz, dbias, M, N, flags, p_drop, rng_state, site, bcast, st);
  if (dtype == MAR_BF16 && out_dtype == MAR_BF16)
    return launch_bwd_epilogue<bf16, bf16>(dout, out, dz, dbias, M, N, flags, p_drop, rng_state, site, bcast, st);
  if (dtype == MAR_F32 && out_dtype == MAR_BF16)
    return launch_bwd_epilogue<float, bf16>(dout, out, dz, dbias, M, N, flags, p_drop, rng_state, site, bcast, st);
  if (dtype == MAR_BF16 && out_dtype == MAR_F32)
    return launch_bwd_epilogue<bf16, float>(dout, out, dz, dbias, M, N, flags, p_drop, rng_state, site, bcast, st);
  MAR_UNSUPPORTED("mar_linear_bwd_epilogue: dtype %d/%d", dtype, out_dtype);
}

int mar_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, void* stream) {
  MAR_CHECK_ARG(src && dst && n >= 0, "mar_cast: bad arguments");
  if (n == 0) return MAR_OK;
  MAR_CHECK_ARG(((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0), "mar_cast: pointers must be 16 B aligned");
  unsigned blocks = (unsigned)ceil_div(ceil_div(n, 8), 256);
  cudaStream_t st = S(stream);
  if (src_dtype == MAR_F32 && dst_dtype == MAR_BF16) mar_launch(cast_kernel<float, bf16>, blocks, 256, 0, st, (const float*)src, (bf16*)dst, n);
  else if (src_dtype == MAR_BF16 && dst_dtype == MAR_F32) mar_launch(cast_kernel<bf16, float>, blocks, 256, 0, st, (const bf16*)src, (float*)dst, n);
  else if (src_dtype == MAR_F32 && dst_dtype == MAR_F32) mar_launch(cast_kernel<float, float>, blocks, 256, 0, st, (const float*)src, (float*)dst, n);
  else if (src_dtype == MAR_BF16 && dst_dtype == MAR_BF16) mar_launch(cast_kernel<bf16, bf16>, blocks, 256, 0, st, (const bf16*)src, (bf16*)dst, n);
  else MAR_UNSUPPORTED("mar_cast: dtype %d -> %d", src_dtype, dst_dtype);
  MAR_LAUNCH_CHECK("cast");
  return MAR_OK;
}

int mar_cast_weight(const float* src, void* w, void* wt, int64_t N, int64_t K, int dtype, void* stream) {
  MAR_CHECK_ARG(src && (w || wt) && N > 0 && K > 0, "mar_cast_weight: bad arguments");
  dim3 grid((unsigned)ceil_div(K, 32), (unsigned)ceil_div(N, 32)), block(32, 8);
  if (dtype == MAR_BF16) cast_weight_kernel<bf16><<<grid, block, 0, S(stream)>>>(src, (bf16*)w, (bf16*)wt, (int)N, (int)K);
  else if (dtype == MAR_F32) cast_weight_kernel<float><<<grid, block, 0, S(stream)>>>(src, (float*)w, (float*)wt, (int)N, (int)K);
  else MAR_UNSUPPORTED("mar_cast_weight: dtype %d", dtype);
  MAR_LAUNCH_CHECK("cast_weight");
  return MAR_OK;
}

int mar_meanpool_fwd(const void* x, void* out, int64_t B, int64_t T, int64_t D, int dtype, void* stream) {
  MAR_CHECK_ARG(x && out && B >= 0 && T > 0 && D > 0, "mar_meanpool_fwd: bad arguments");
  MAR_CHECK_ARG(D % 8 == 0, "mar_meanpool_fwd: D must be a multiple of 8 (got %lld)", (long long)D);
  if (B == 0) return MAR_OK;
  dim3 grid((unsigned)ceil_div(D, 256), (unsigned)B), block(32, 8);
  if (dtype == MAR_BF16) mar_launch(meanpool_fwd_kernel<bf16>, grid, block, 0, S(stream), (const bf16*)x, (bf16*)out, T, D);
  else if (dtype == MAR_F32) mar_launch(meanpool_fwd_kernel<float>, grid, block, 0, S(stream), (const float*)x, (float*)out, T, D);
  else MAR_UNSUPPORTED("mar_meanpool_fwd: dtype %d", dtype);
  MAR_LAUNCH_CHECK("meanpool_fwd");
  return MAR_OK;
}

int mar_meanpool_bwd(const void* dout, void* dx, int64_t B, int64_t T, int64_t D, int dtype, void* stream) {
  MAR_CHECK_ARG(dout && dx && B >= 0 && T > 0 && D > 0, "mar_meanpool_bwd: bad arguments");
  MAR_CHECK_ARG(D % 8 == 0, "mar_meanpool_bwd: D must be a multiple of 8 (got %lld)", (long long)D);
  if (B == 0) return MAR_OK;
  int64_t total8 = B * T * D / 8;
  unsigned blocks = (unsigned)ceil_div(total8, 256);
  if (dtype == MAR_BF16) meanpool_bwd_kernel<bf16><<<blocks, 256, 0, S(stream)>>>((const bf16*)dout, (bf16*)dx, T, D, total8);
  else if (dtype == MAR_F32) meanpool_bwd_kernel<float><<<blocks, 256, 0, S(stream)>>>((const float*)dout, (float*)dx, T, D, total8);
  else MAR_UNSUPPORTED("mar_meanpool_bwd: dtype %d", dtype);
  MAR_LAUNCH_CHECK("meanpool_bwd");
  return MAR_OK;
}

int mar_rowzero_mask(const void* x, uint8_t* mask, int64_t rows, int64_t D, int dtype, void* stream) {
  MAR_CHECK_ARG(x && mask && rows >= 0 && D > 0, "mar_rowzero_mask: bad arguments");
  MAR_CHECK_ARG(D % 8 == 0, "mar_rowzero_mask: D must be a multiple of 8 (got %lld)", (long long)D);
  if (rows == 0) return MAR_OK;
  unsigned blocks = (unsigned)ceil_div(rows, 8);
  if (dtype == MAR_BF16) mar_launch(rowzero_kernel<bf16>, blocks, 256, 0, S(stream), (const bf16*)x, mask, rows, D);
  else if (dtype == MAR_F32) mar_launch(rowzero_kernel<float>, blocks, 256, 0, S(stream), (const float*)x, mask, rows, D);
  else MAR_UNSUPPORTED("mar_rowzero_mask: dtype %d", dtype);
  MAR_LAUNCH_CHECK("rowzero_mask");
  return MAR_OK;
}

int mar_concat_rows(const void* src, void* dst, int64_t B, int64_t T, int64_t T_total, int64_t t_off, int64_t D,
                    int dtype, int to_concat, void* stream) {
  MAR_CHECK_ARG(src && dst && B >= 0 && T > 0 && T_total >= t_off + T && t_off >= 0, "mar_concat_rows: bad arguments");
  const int64_t esz = dtype == MAR_BF16 ? 2 : 4;
  MAR_CHECK_ARG((D * esz) % 16 == 0, "mar_concat_rows: row bytes must be a multiple of 16");
  if (B == 0) return MAR_OK;
  int64_t row_vecs = D * esz / 16;
  int64_t total = B * T * row_vecs;
  concat_rows_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, S(stream)>>>((const uint4*)src, (uint4*)dst, T, T_total,
                                                                          t_off, row_vecs, total, to_concat);
  MAR_LAUNCH_CHECK("concat_rows");
  return MAR_OK;
}

int mar_cross_entropy_fwd(const float* logits, const int64_t* labels, const float* class_weight, float* loss,
                          float* dlogits, int64_t* preds, int64_t B, int64_t C, void* stream) {
  MAR_CHECK_ARG(logits && labels && loss && B > 0 && C > 0, "mar_cross_entropy_fwd: bad arguments");
  mar_launch(cross_entropy_kernel, 1, 256, 0, S(stream), logits, labels, class_weight, loss, dlogits, preds, B, C);
  MAR_LAUNCH_CHECK("cross_entropy");
  return MAR_OK;
}

int mar_focal_loss_fwd(const float* logits, const int64_t* labels, const float* alpha, float gamma, float* loss,
                       float* dlogits, int64_t B, int64_t C, void* stream) {
  MAR_CHECK_ARG(logits && labels && loss && B > 0 && C > 0, "mar_focal_loss_fwd: bad arguments");
  MAR_CHECK_ARG(gamma >= 0.f, "mar_focal_loss_fwd: gamma must be >= 0");
  focal_loss_kernel<<<1, 256, 0, S(stream)>>>(logits, labels, alpha, gamma, loss, dlogits, B, C);
  MAR_LAUNCH_CHECK("focal_loss");
  return MAR_OK;
}

int mar_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, const float* step_dev,
                  int64_t n, float lr, float beta1, float beta2, float eps, void* stream) {
  MAR_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && step_dev && n >= 0, "mar_adam_step: bad arguments");
  if (n == 0) return MAR_OK;
  MAR_CHECK_ARG(((uintptr_t)param % 16 == 0) && ((uintptr_t)grad % 16 == 0) && ((uintptr_t)exp_avg % 16 == 0) &&
                    ((uintptr_t)exp_avg_sq % 16 == 0), "mar_adam_step: buffers must be 16 B aligned");
  adam_kernel<<<(unsigned)ceil_div(ceil_div(n, 4), 256), 256, 0, S(stream)>>>(param, grad, exp_avg, exp_avg_sq, step_dev, n,
                                                                            lr, beta1, beta2, eps);
  MAR_LAUNCH_CHECK("adam");
  return MAR_OK;
}
int mar_adam_tick(float* step_dev, void* stream) {
  MAR_CHECK_ARG(step_dev, "mar_adam_tick: null");
  adam_tick_kernel<<<1, 1, 0, S(stream)>>>(step_dev);
  MAR_LAUNCH_CHECK("adam_tick");
  return MAR_OK;
}

int mar_adam_step_segments(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                           const int32_t* chunk_seg, float* seg_steps, const float* seg_active, float* seg_coef,
                           int64_t n, int chunk, int nseg, float lr, float beta1, float beta2, float eps, void* bf16_mirror,
                           void* stream) {
  MAR_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && chunk_seg && seg_steps && seg_active && seg_coef && n >= 0 &&
                    nseg > 0, "mar_adam_step_segments: bad arguments");
  MAR_CHECK_ARG(chunk >= 4 && (chunk & (chunk - 1)) == 0, "mar_adam_step_segments: chunk must be a power of two >= 4");
  MAR_CHECK_ARG(n % chunk == 0, "mar_adam_step_segments: n must be a multiple of chunk");
  if (n == 0) return MAR_OK;
  MAR_CHECK_ARG(((uintptr_t)param % 16 == 0) && ((uintptr_t)grad % 16 == 0) && ((uintptr_t)exp_avg % 16 == 0) &&
                    ((uintptr_t)exp_avg_sq % 16 == 0) && ((uintptr_t)seg_coef % 8 == 0) && ((uintptr_t)bf16_mirror % 8 == 0),
                "mar_adam_step_segments: buffers must be 16 B aligned");
  int shift = 0;
  while ((1 << shift) < chunk) shift++;
  mar_launch(adam_seg_tick_kernel, (unsigned)ceil_div(nseg, 128), 128, 0, S(stream), seg_steps, seg_active,
                                                                            reinterpret_cast<float2*>(seg_coef), nseg, lr, beta1, beta2);
  MAR_LAUNCH_CHECK("adam_seg_tick");
  mar_launch(adam_seg_kernel, (unsigned)ceil_div(ceil_div(n, 4), 256), 256, 0, S(stream), 
      param, grad, exp_avg, exp_avg_sq, chunk_seg, reinterpret_cast<const float2*>(seg_coef), n, shift, beta1, beta2, eps,
      reinterpret_cast<bf16*>(bf16_mirror));
  MAR_LAUNCH_CHECK("adam_seg");
  return MAR_OK;
}

int mar_label_weight_sum(const int64_t* labels, const float* class_weight, float* out, int64_t B, int64_t C, void* stream) {
  MAR_CHECK_ARG(labels && out && B >= 0 && C > 0, "mar_label_weight_sum: bad arguments");
  label_weight_sum_kernel<<<1, 256, 0, S(stream)>>>(labels, class_weight, out, B, C);
  MAR_LAUNCH_CHECK("label_weight_sum");
  return MAR_OK;
}

}  // extern "C"
