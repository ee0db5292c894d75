// GRU / LSTM recurrence, step-per-launch engine: one small GEMM (h·W_hhᵀ) plus one fused gate kernel
// per time step.  This is the fp32-mode engine and the checker for the persistent cluster kernel
// (rnn_persistent.cu).  Gate order and arithmetic follow nn.GRU / nn.LSTM (cuDNN order r,z,n and
// i,f,g,o), h0 = c0 = 0, no sequence lengths (the recurrence also runs over zero padding).
#include "common.cuh"
#include "gemm_simt.cuh"
#include "rnn.cuh"

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// ---- GRU ------------------------------------------------------------------------------------
template <typename T>
__global__ void gru_cell_fwd_kernel(const T* __restrict__ gi, const float* __restrict__ gh, const float* __restrict__ b_hh,
                                    const float* __restrict__ h_in, float* __restrict__ h_out, T* __restrict__ hseq,
                                    T* __restrict__ hprev, float* __restrict__ saved, int64_t B, int64_t Tn, int64_t H,
                                    int64_t t) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H) return;
  const int64_t b = idx / H, j = idx % H;
  const T* g = gi + (b * Tn + t) * 3 * H;
  float ghr, ghz, ghn, hp;
  if (t == 0) { ghr = b_hh[j]; ghz = b_hh[H + j]; ghn = b_hh[2 * H + j]; hp = 0.f; }
  else {
    const float* q = gh + b * 3 * H;
    ghr = q[j]; ghz = q[H + j]; ghn = q[2 * H + j];
    hp = h_in[idx];                       // fp32 carry: only the GEMM operand is rounded to T
  }
  const float r = sigmoidf_(to_f32<T>(g[j]) + ghr);
  const float z = sigmoidf_(to_f32<T>(g[H + j]) + ghz);
  const float n = tanhf(to_f32<T>(g[2 * H + j]) + r * ghn);
  const float h = (1.f - z) * n + z * hp;
  h_out[idx] = h;
  hseq[(b * Tn + t) * H + j] = from_f32<T>(h);
  if (saved != nullptr) {
    float* s = saved + (b * Tn + t) * 5 * H;
    s[j] = r; s[H + j] = z; s[2 * H + j] = n; s[3 * H + j] = ghn; s[4 * H + j] = hp;
    hprev[(b * Tn + t) * H + j] = from_f32<T>(hp);
  }
}

template <typename T>
__global__ void gru_cell_bwd_kernel(const T* __restrict__ dhseq, const float* __restrict__ dh_carry,
                                    const float* __restrict__ saved, T* __restrict__ dgi,
                                    T* __restrict__ dgh, float* __restrict__ dh_direct, int64_t B, int64_t Tn, int64_t H,
                                    int64_t t, int has_carry) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H) return;
  const int64_t b = idx / H, j = idx % H;
  const float* s = saved + (b * Tn + t) * 5 * H;
  const float r = s[j], z = s[H + j], n = s[2 * H + j], hn = s[3 * H + j], hp = s[4 * H + j];
  float dh = to_f32<T>(dhseq[(b * Tn + t) * H + j]);
  if (has_carry) dh += dh_carry[idx];
  const float dn_pre = dh * (1.f - z) * (1.f - n * n);
  const float dz_pre = dh * (hp - n) * z * (1.f - z);
  const float dr_pre = dn_pre * hn * r * (1.f - r);
  T* a = dgi + (b * Tn + t) * 3 * H;
  T* c = dgh + (b * Tn + t) * 3 * H;
  a[j] = from_f32<T>(dr_pre); a[H + j] = from_f32<T>(dz_pre); a[2 * H + j] = from_f32<T>(dn_pre);
  c[j] = from_f32<T>(dr_pre); c[H + j] = from_f32<T>(dz_pre); c[2 * H + j] = from_f32<T>(dn_pre * r);
  dh_direct[idx] = dh * z;
}

// ---- LSTM -----------------------------------------------------------------------------------
template <typename T>
__global__ void lstm_cell_fwd_kernel(const T* __restrict__ gi, const float* __restrict__ gh, const float* __restrict__ b_hh,
                                     T* __restrict__ hseq, T* __restrict__ hprev, float* __restrict__ saved,
                                     float* __restrict__ c_state, int64_t B, int64_t Tn, int64_t H, int64_t t) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H) return;
  const int64_t b = idx / H, j = idx % H;
  const T* g = gi + (b * Tn + t) * 4 * H;
  float a[4];
#pragma unroll
  for (int k = 0; k < 4; k++)
    a[k] = to_f32<T>(g[k * H + j]) + (t == 0 ? b_hh[k * H + j] : gh[b * 4 * H + k * H + j]);
  const float i = sigmoidf_(a[0]), f = sigmoidf_(a[1]), gg = tanhf(a[2]), o = sigmoidf_(a[3]);
  const float cp = t == 0 ? 0.f : c_state[idx];
  const float c = f * cp + i * gg;
  c_state[idx] = c;
  hseq[(b * Tn + t) * H + j] = from_f32<T>(o * tanhf(c));
  if (saved != nullptr) {
    float* s = saved + (b * Tn + t) * 5 * H;
    s[j] = i; s[H + j] = f; s[2 * H + j] = gg; s[3 * H + j] = o; s[4 * H + j] = c;
    hprev[(b * Tn + t) * H + j] = t == 0 ? from_f32<T>(0.f) : hseq[(b * Tn + t - 1) * H + j];
  }
}

template <typename T>
__global__ void lstm_cell_bwd_kernel(const T* __restrict__ dhseq, const float* __restrict__ dh_carry,
                                     float* __restrict__ dc_carry, const float* __restrict__ saved, T* __restrict__ dgates,
                                     int64_t B, int64_t Tn, int64_t H, int64_t t, int has_carry) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H) return;
  const int64_t b = idx / H, j = idx % H;
  const float* s = saved + (b * Tn + t) * 5 * H;
  const float i = s[j], f = s[H + j], g = s[2 * H + j], o = s[3 * H + j], c = s[4 * H + j];
  const float cp = t > 0 ? saved[(b * Tn + t - 1) * 5 * H + 4 * H + j] : 0.f;
  float dh = to_f32<T>(dhseq[(b * Tn + t) * H + j]);
  float dc = 0.f;
  if (has_carry) { dh += dh_carry[idx]; dc = dc_carry[idx]; }
  const float tc = tanhf(c);
  dc += dh * o * (1.f - tc * tc);
  T* d = dgates + (b * Tn + t) * 4 * H;
  d[j] = from_f32<T>(dc * g * i * (1.f - i));
  d[H + j] = from_f32<T>(dc * cp * f * (1.f - f));
  d[2 * H + j] = from_f32<T>(dc * i * (1.f - g * g));
  d[3 * H + j] = from_f32<T>(dh * tc * o * (1.f - o));
  dc_carry[idx] = dc * f;
}

template <typename T>
int gru_fwd_impl(const void* gi, const void* w_hh, const float* b_hh, void* hseq, void* hprev, float* saved, float* work,
                 int64_t B, int64_t Tn, int64_t H, int dtype, cudaStream_t st) {
  const unsigned blocks = (unsigned)ceil_div(B * H, 256);
  float* hbuf[2] = {work + B * 3 * H, work + B * 4 * H};    // fp32 h ping-pong
  for (int64_t t = 0; t < Tn; t++) {
    if (t > 0) {
      // gh (B,3H) fp32 = h_{t-1} (B,H; row stride T*H) · W_hhᵀ + b_hh
      SimtEpilogue epi;
      epi.bias = b_hh;
      int rc = gemm_simt(reinterpret_cast<const T*>(hseq) + (t - 1) * H, dtype, Tn * H, 1, w_hh, dtype, 1, H, work,
                         MAR_F32, 3 * H, B, 3 * H, H, epi, st);
      if (rc) return rc;
    }
    gru_cell_fwd_kernel<T><<<blocks, 256, 0, st>>>((const T*)gi, work, b_hh, hbuf[(t + 1) & 1], hbuf[t & 1], (T*)hseq,
                                                   (T*)hprev, saved, B, Tn, H, t);
    MAR_LAUNCH_CHECK("gru_cell_fwd");
  }
  return MAR_OK;
}

template <typename T>
int gru_bwd_impl(const void* dhseq, const float* saved, const void* w_hh, void* dgi, void* dgh,
                 float* work, int64_t B, int64_t Tn, int64_t H, int dtype, cudaStream_t st) {
  const unsigned blocks = (unsigned)ceil_div(B * H, 256);
  float* dh_carry = work;          // (B,H)
  float* dh_direct = work + B * H; // (B,H)
  for (int64_t t = Tn - 1; t >= 0; t--) {
    gru_cell_bwd_kernel<T><<<blocks, 256, 0, st>>>((const T*)dhseq, dh_carry, saved, (T*)dgi, (T*)dgh, dh_direct, B, Tn,
                                                   H, t, t != Tn - 1);
    MAR_LAUNCH_CHECK("gru_cell_bwd");
    if (t > 0) {
      // dh_carry (B,H) fp32 = dgh_t (B,3H; row stride T*3H) · W_hh (3H,H) + dh_direct
      SimtEpilogue epi;
      epi.residual = dh_direct; epi.ldr = H; epi.res_is_bf16 = 0;
      int rc = gemm_simt(reinterpret_cast<const T*>(dgh) + t * 3 * H, dtype, Tn * 3 * H, 1, w_hh, dtype, H, 1, dh_carry,
                         MAR_F32, H, B, H, 3 * H, epi, st);
      if (rc) return rc;
    }
  }
  return MAR_OK;
}

template <typename T>
int lstm_fwd_impl(const void* gi, const void* w_hh, const float* b_hh, void* hseq, void* hprev, float* saved, float* work,
                  int64_t B, int64_t Tn, int64_t H, int dtype, cudaStream_t st) {
  const unsigned blocks = (unsigned)ceil_div(B * H, 256);
  float* gh = work;                // (B,4H)
  float* c_state = work + B * 4 * H;  // (B,H)
  for (int64_t t = 0; t < Tn; t++) {
    if (t > 0) {
      SimtEpilogue epi;
      epi.bias = b_hh;
      int rc = gemm_simt(reinterpret_cast<const T*>(hseq) + (t - 1) * H, dtype, Tn * H, 1, w_hh, dtype, 1, H, gh, MAR_F32,
                         4 * H, B, 4 * H, H, epi, st);
      if (rc) return rc;
    }
    lstm_cell_fwd_kernel<T><<<blocks, 256, 0, st>>>((const T*)gi, gh, b_hh, (T*)hseq, (T*)hprev, saved, c_state, B, Tn, H, t);
    MAR_LAUNCH_CHECK("lstm_cell_fwd");
  }
  return MAR_OK;
}

template <typename T>
int lstm_bwd_impl(const void* dhseq, const float* saved, const void* w_hh, void* dgates, float* work, int64_t B,
                  int64_t Tn, int64_t H, int dtype, cudaStream_t st) {
  const unsigned blocks = (unsigned)ceil_div(B * H, 256);
  float* dh_carry = work;
  float* dc_carry = work + B * H;
  for (int64_t t = Tn - 1; t >= 0; t--) {
    lstm_cell_bwd_kernel<T><<<blocks, 256, 0, st>>>((const T*)dhseq, dh_carry, dc_carry, saved, (T*)dgates, B, Tn, H, t,
                                                    t != Tn - 1);
    MAR_LAUNCH_CHECK("lstm_cell_bwd");
    if (t > 0) {
      SimtEpilogue epi;
      int rc = gemm_simt(reinterpret_cast<const T*>(dgates) + t * 4 * H, dtype, Tn * 4 * H, 1, w_hh, dtype, H, 1,
                         dh_carry, MAR_F32, H, B, H, 4 * H, epi, st);
      if (rc) return rc;
    }
  }
  return MAR_OK;
}

}  // namespace

int gru_fwd_steps(const void* gi, const void* w_hh, const float* b_hh, void* hseq, void* hprev, float* saved, float* work,
                  int64_t B, int64_t T, int64_t H, int dtype, cudaStream_t st) {
  if (dtype == MAR_BF16) return gru_fwd_impl<bf16>(gi, w_hh, b_hh, hseq, hprev, saved, work, B, T, H, dtype, st);
  return gru_fwd_impl<float>(gi, w_hh, b_hh, hseq, hprev, saved, work, B, T, H, dtype, st);
}
int gru_bwd_steps(const void* dhseq, const float* saved, const void* w_hh, void* dgi, void* dgh, float* work, int64_t B,
                  int64_t T, int64_t H, int dtype, cudaStream_t st) {
  if (dtype == MAR_BF16) return gru_bwd_impl<bf16>(dhseq, saved, w_hh, dgi, dgh, work, B, T, H, dtype, st);
  return gru_bwd_impl<float>(dhseq, saved, w_hh, dgi, dgh, work, B, T, H, dtype, st);
}
int lstm_fwd_steps(const void* gi, const void* w_hh, const float* b_hh, void* hseq, void* hprev, float* saved, float* work,
                   int64_t B, int64_t T, int64_t H, int dtype, cudaStream_t st) {
  if (dtype == MAR_BF16) return lstm_fwd_impl<bf16>(gi, w_hh, b_hh, hseq, hprev, saved, work, B, T, H, dtype, st);
  return lstm_fwd_impl<float>(gi, w_hh, b_hh, hseq, hprev, saved, work, B, T, H, dtype, st);
}
int lstm_bwd_steps(const void* dhseq, const float* saved, const void* w_hh, void* dgates, float* work, int64_t B,
                   int64_t T, int64_t H, int dtype, cudaStream_t st) {
  if (dtype == MAR_BF16) return lstm_bwd_impl<bf16>(dhseq, saved, w_hh, dgates, work, B, T, H, dtype, st);
  return lstm_bwd_impl<float>(dhseq, saved, w_hh, dgates, work, B, T, H, dtype, st);
}
