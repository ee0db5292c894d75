// C ABI entry points (include/mar.h): argument validation, engine selection, error plumbing.
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include "common.cuh"
#include "gemm_simt.cuh"
#include "gemm_skinny.cuh"
#include "gemm_tcgen05.cuh"
#include "attention.cuh"
#include "rnn.cuh"

namespace {
thread_local char g_err[512] = "";
thread_local int g_engine = 0;
std::atomic<int64_t> g_launches{0};
int g_sm_count = 0;
inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
}  // namespace

void mar_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void mar_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
void mar_set_engine(int e) { g_engine = e; }
bool mar_debug_sync() {
  static const bool on = [] { const char* v = getenv("MAR_DEBUG_SYNC"); return v != nullptr && v[0] != '\0' && v[0] != '0'; }();
  return on;
}

bool mar_pdl_enabled(int kernel_class) {
  // default 1: the tcgen05 GEMM launches only — measured per class on B200 (profiles/r02_notes_measured_dead_ends.md): never
  // slower there, no gain for the attention / elementwise classes, LayerNorm launches get SLOWER (their CTAs become
  // resident beside the previous persistent GEMM)
  static const int mask = [] { const char* v = getenv("MAR_PDL"); return v != nullptr ? atoi(v) : 1; }();
  return (mask & kernel_class) != 0;
}

int mar_sm_count() {
  if (g_sm_count == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      g_sm_count = n;
    else
      return 148;
  }
  return g_sm_count;
}

static bool env_flag(const char* name) {
  const char* v = getenv(name);
  return v != nullptr && v[0] != '\0' && v[0] != '0';
}

extern "C" {

int mar_version(void) { return MAR_VERSION; }
const char* mar_last_error(void) { return g_err; }
int64_t mar_launch_count(void) { return g_launches.load(); }
void mar_launch_count_reset(void) { g_launches.store(0); }
int mar_last_engine(void) { return g_engine; }

int mar_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  MAR_CUDA(cudaGetDevice(&dev));
  int sm = 0, maj = 0, min = 0;
  MAR_CUDA(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev));
  MAR_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
  MAR_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count) *sm_count = sm;
  if (cc_major) *cc_major = maj;
  if (cc_minor) *cc_minor = min;
  if (maj != 10) {
    mar_set_error("libmar is built for sm_100a only; device is sm_%d%d", maj, min);
    return MAR_ERR_ARCH;
  }
  return MAR_OK;
}

// ------------------------------------------------------------------------------------------
int mar_linear_fwd(const void* x, int64_t ldx, const void* w, const float* bias, const void* residual, int64_t ldr,
                   void* out, int64_t ldo, int64_t M, int64_t N, int64_t K, int in_dtype, int out_dtype, int flags,
                   float p_drop, const uint64_t* rng_state, uint32_t site, int engine, void* stream) {
  MAR_CHECK_ARG(x && w && out, "mar_linear_fwd: null pointer");
  MAR_CHECK_ARG(M >= 0 && N > 0 && K > 0 && ldx >= K && ldo >= N, "mar_linear_fwd: bad shape M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
  MAR_CHECK_ARG(in_dtype == MAR_F32 || in_dtype == MAR_BF16, "mar_linear_fwd: bad in_dtype %d", in_dtype);
  MAR_CHECK_ARG(out_dtype == MAR_F32 || out_dtype == MAR_BF16, "mar_linear_fwd: bad out_dtype %d", out_dtype);
  MAR_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, "mar_linear_fwd: p_drop out of range");
  const bool drop = (flags & MAR_EPI_DROPOUT) && p_drop > 0.f;
  MAR_CHECK_ARG(!drop || rng_state, "mar_linear_fwd: dropout needs rng_state");
  MAR_CHECK_ARG(!residual || ldr >= N, "mar_linear_fwd: bad residual stride");
  if (M == 0) return MAR_OK;
  if (!drop) flags &= ~MAR_EPI_DROPOUT;

  TcGemmArgs a;
  a.A = x; a.lda = ldx; a.B = w; a.ldb = K; a.out = out; a.ldo = ldo; a.out_fp32 = out_dtype == MAR_F32;
  a.M = M; a.N = N; a.Kr = K; a.bias = bias; a.residual = residual; a.ldr = ldr; a.flags = flags; a.p_drop = p_drop;
  a.rng = rng_state; a.site = site;
  const bool tc_ok = in_dtype == MAR_BF16 && gemm_tcgen05_supported(a) && (bias == nullptr || ((uintptr_t)bias % 16) == 0) &&
                     !(residual != nullptr && out_dtype == MAR_F32);
  if (engine == MAR_ENGINE_TCGEN05 && !tc_ok) MAR_UNSUPPORTED("mar_linear_fwd: tcgen05 engine cannot take M=%lld N=%lld K=%lld dtype=%d", (long long)M, (long long)N, (long long)K, in_dtype);
  const bool use_tc = engine == MAR_ENGINE_TCGEN05 || (engine == MAR_ENGINE_AUTO && tc_ok && M >= 64 && N >= 64 && !env_flag("MAR_FORCE_SIMT"));
  if (use_tc) return gemm_tcgen05(a, S(stream));

  if (skinny_supported(N) && flags == 0 && residual == nullptr && !(in_dtype == MAR_F32 && out_dtype == MAR_BF16))
    return skinny_fwd(x, ldx, w, bias, out, ldo, M, N, K, in_dtype, out_dtype, S(stream));
  SimtEpilogue epi;
  epi.bias = bias; epi.residual = residual; epi.ldr = ldr; epi.res_is_bf16 = in_dtype == MAR_BF16;
  epi.flags = flags; epi.p = p_drop; epi.rng = rng_state; epi.site = site;
  return gemm_simt(x, in_dtype, ldx, 1, w, in_dtype, 1, K, out, out_dtype, ldo, M, N, K, epi, S(stream));
}

int mar_linear_dgrad(const void* dz, const void* w, const void* wt, const void* add, const void* act, float act_scale,
                     void* dx, int64_t lddx, float* dx_colsum, int64_t M, int64_t N, int64_t K, int dtype, int engine,
                     void* stream) {
  MAR_CHECK_ARG(dz && (w || wt) && dx, "mar_linear_dgrad: null pointer");
  MAR_CHECK_ARG(!dx_colsum || lddx == K, "mar_linear_dgrad: dx_colsum needs a contiguous dx");
  MAR_CHECK_ARG(M >= 0 && N > 0 && K > 0 && lddx >= K, "mar_linear_dgrad: bad shape");
  MAR_CHECK_ARG(dtype == MAR_F32 || dtype == MAR_BF16, "mar_linear_dgrad: bad dtype %d", dtype);
  MAR_CHECK_ARG(!(add && act), "mar_linear_dgrad: add and act are mutually exclusive");
  if (M == 0) return MAR_OK;
  TcGemmArgs a;
  a.A = dz; a.lda = N; a.out = dx; a.ldo = lddx; a.out_fp32 = 0;
  a.M = M; a.N = K; a.Kr = N; a.residual = add; a.ldr = lddx; a.aux = act; a.ldaux = lddx; a.aux_scale = act_scale;
  a.colsum = dx_colsum;
  // B operand: W itself, (N,K) row-major = MN-major over the reduction dimension N (no transposed copy needed);
  // a caller-provided Wᵀ (K,N) is used as a K-major operand when W is absent (MAR_DGRAD_WT=1 prefers it, for A/B runs)
  if (w != nullptr && !(wt != nullptr && env_flag("MAR_DGRAD_WT"))) { a.B = w; a.ldb = K; a.b_mn_major = 1; }
  else { a.B = wt; a.ldb = N; a.b_mn_major = 0; }
  const bool tc_ok = dtype == MAR_BF16 && a.B != nullptr && gemm_tcgen05_supported(a);
  if (engine == MAR_ENGINE_TCGEN05 && !tc_ok) MAR_UNSUPPORTED("mar_linear_dgrad: tcgen05 engine cannot take M=%lld N=%lld K=%lld (needs bf16)", (long long)M, (long long)N, (long long)K);
  const bool use_tc = engine == MAR_ENGINE_TCGEN05 || (engine == MAR_ENGINE_AUTO && tc_ok && M >= 64 && K >= 64 && !env_flag("MAR_FORCE_SIMT"));
  if (use_tc) return gemm_tcgen05(a, S(stream));   // column sums leave the GEMM's own epilogue
  int rc;
  if (skinny_supported(N) && w != nullptr && act == nullptr) {
    rc = skinny_dgrad(dz, w, add, dx, lddx, M, N, K, dtype, S(stream));
  } else {
    SimtEpilogue epi;
    epi.residual = add; epi.ldr = lddx; epi.res_is_bf16 = dtype == MAR_BF16;
    epi.aux = act; epi.ldaux = lddx; epi.aux_scale = act_scale;
    if (w != nullptr)   // dx(m,k) = Σ_n dz(m,n) W(n,k):  B(kr=n, col=k) = w[n*K + k]
      rc = gemm_simt(dz, dtype, N, 1, w, dtype, K, 1, dx, dtype, lddx, M, K, N, epi, S(stream));
    else
      rc = gemm_simt(dz, dtype, N, 1, wt, dtype, 1, N, dx, dtype, lddx, M, K, N, epi, S(stream));
  }
  if (rc == MAR_OK && dx_colsum != nullptr)        // the SIMT engines sum the columns in a pass of their own
    rc = mar_linear_bwd_epilogue(dx, nullptr, nullptr, dx_colsum, M, K, dtype, dtype, 0, 0.f, nullptr, 0, 0, stream);
  return rc;
}

int mar_linear_wgrad(const void* dz, const void* x, int64_t ldx, float* dw, int64_t M, int64_t N, int64_t K, int dtype,
                     int accumulate, int engine, void* stream) {
  MAR_CHECK_ARG(dz && x && dw, "mar_linear_wgrad: null pointer");
  MAR_CHECK_ARG(M >= 0 && N > 0 && K > 0 && ldx >= K, "mar_linear_wgrad: bad shape");
  MAR_CHECK_ARG(dtype == MAR_F32 || dtype == MAR_BF16, "mar_linear_wgrad: bad dtype %d", dtype);
  if (M == 0) {
    if (!accumulate) MAR_CUDA(cudaMemsetAsync(dw, 0, (size_t)N * K * 4, S(stream)));
    return MAR_OK;
  }
  // dW(n,k) = Σ_rows dz(row,n) x(row,k): "M"=N, "N"=K, reduction over the M token rows.
  TcGemmArgs a;
  a.A = dz; a.lda = N; a.a_mn_major = 1; a.B = x; a.ldb = ldx; a.b_mn_major = 1; a.out = dw; a.ldo = K; a.out_fp32 = 1;
  a.M = N; a.N = K; a.Kr = M; a.accumulate = accumulate; a.allow_split = 1;
  const bool tc_ok = dtype == MAR_BF16 && gemm_tcgen05_supported(a);
  if (engine == MAR_ENGINE_TCGEN05 && !tc_ok) MAR_UNSUPPORTED("mar_linear_wgrad: tcgen05 engine cannot take M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
  const bool use_tc = engine == MAR_ENGINE_TCGEN05 || (engine == MAR_ENGINE_AUTO && tc_ok && N >= 64 && K >= 64 && M >= 64 && !env_flag("MAR_FORCE_SIMT"));
  if (use_tc) return gemm_tcgen05(a, S(stream));
  if (skinny_supported(N)) return skinny_wgrad(dz, x, ldx, dw, M, N, K, dtype, accumulate, S(stream));
  SimtEpilogue epi;
  epi.accumulate = accumulate; epi.allow_split = 1;
  // A(m=n, kr=row) = dz[row*N + n] ; B(kr=row, col=k) = x[row*ldx + k]
  return gemm_simt(dz, dtype, 1, N, x, dtype, ldx, 1, dw, MAR_F32, K, N, K, M, epi, S(stream));
}

// ------------------------------------------------------------------------------------------
int mar_attention_fwd(const void* qkv, const uint8_t* key_mask, void* out, float* lse, int64_t B, int64_t T, int64_t H,
                      int64_t dh, int dtype, float p_drop, const uint64_t* rng_state, uint32_t site, uint32_t* drop_bits,
                      int engine, void* stream) {
  MAR_CHECK_ARG(qkv && out && lse, "mar_attention_fwd: null pointer");
  MAR_CHECK_ARG(B >= 0 && T > 0 && H > 0 && dh > 0, "mar_attention_fwd: bad shape");
  MAR_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, "mar_attention_fwd: p_drop out of range");
  MAR_CHECK_ARG(p_drop == 0.f || (rng_state && drop_bits), "mar_attention_fwd: dropout needs rng_state and the keep-bit buffer");
  MAR_CHECK_ARG((uintptr_t)drop_bits % 16 == 0, "mar_attention_fwd: drop_bits must be 16 B aligned");
  MAR_CHECK_ARG(B * H < (1ll << 31) && T < (1ll << 31), "mar_attention_fwd: shape too large");
  if (B == 0) return MAR_OK;
  if (p_drop > 0.f) {          // draw this call's keep bits once: every engine, and the backward pass, reads them
    int rc = attention_dropbits(drop_bits, B, T, H, p_drop, rng_state, site, S(stream));
    if (rc) return rc;
  }
  const uint32_t* dbits = p_drop > 0.f ? drop_bits : nullptr;
  const bool mma_ok = attention_mma_supported(T, dh, dtype);
  if (engine == MAR_ENGINE_TCGEN05 && !mma_ok) MAR_UNSUPPORTED("mar_attention_fwd: tensor-core engine cannot take dh=%lld dtype=%d", (long long)dh, dtype);
  if (engine == MAR_ENGINE_TCGEN05 || (engine == MAR_ENGINE_AUTO && mma_ok && !env_flag("MAR_FORCE_SIMT"))) {
    mar_set_engine(MAR_ENGINE_TCGEN05);
    // tcgen05/TMEM kernel where it applies; the mma.sync kernel covers the remaining head dims (MAR_ATTN_MMA=1 forces it)
    if (attention_tc_supported(B, T, H, dh, dtype) && !env_flag("MAR_ATTN_MMA"))
      return attention_fwd_tc(qkv, key_mask, out, lse, B, T, H, dh, p_drop, dbits, S(stream));
    return attention_fwd_mma(qkv, key_mask, out, lse, B, T, H, dh, p_drop, dbits, S(stream));
  }
  mar_set_engine(MAR_ENGINE_SIMT);
  return attention_fwd_simt(qkv, key_mask, out, lse, B, T, H, dh, dtype, p_drop, dbits, S(stream));
}

int64_t mar_attention_dropbits_words(int64_t B, int64_t T, int64_t H) {
  if (B <= 0 || T <= 0 || H <= 0) return 0;
  return attention_dropbits_words(B, T, H);
}

int mar_attention_bwd(const void* qkv, const uint8_t* key_mask, const void* out, const void* dout, const float* lse,
                      float* delta, void* dqkv, float* dqkv_colsum, int64_t B, int64_t T, int64_t H, int64_t dh, int dtype,
                      float p_drop, const uint32_t* drop_bits, int engine, void* stream) {
  MAR_CHECK_ARG(qkv && out && dout && lse && delta && dqkv, "mar_attention_bwd: null pointer");
  MAR_CHECK_ARG(B >= 0 && T > 0 && H > 0 && dh > 0, "mar_attention_bwd: bad shape");
  MAR_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, "mar_attention_bwd: p_drop out of range");
  MAR_CHECK_ARG(p_drop == 0.f || drop_bits, "mar_attention_bwd: dropout needs the keep bits of the forward call");
  MAR_CHECK_ARG((uintptr_t)drop_bits % 16 == 0, "mar_attention_bwd: drop_bits must be 16 B aligned");
  if (B == 0) return MAR_OK;
  const uint32_t* dbits = p_drop > 0.f ? drop_bits : nullptr;
  const bool mma_ok = attention_mma_supported(T, dh, dtype);
  if (engine == MAR_ENGINE_TCGEN05 && !mma_ok) MAR_UNSUPPORTED("mar_attention_bwd: tensor-core engine cannot take dh=%lld dtype=%d", (long long)dh, dtype);
  int rc;
  if (engine == MAR_ENGINE_TCGEN05 || (engine == MAR_ENGINE_AUTO && mma_ok && !env_flag("MAR_FORCE_SIMT"))) {
    mar_set_engine(MAR_ENGINE_TCGEN05);
    if (attention_tc_supported(B, T, H, dh, dtype) && !env_flag("MAR_ATTN_MMA") && !env_flag("MAR_ATTN_BWD_MMA"))   // sums inside the kernel
      return attention_bwd_tc(qkv, key_mask, out, dout, lse, delta, dqkv, dqkv_colsum, B, T, H, dh, p_drop, dbits, S(stream));
    rc = attention_bwd_mma(qkv, key_mask, out, dout, lse, delta, dqkv, B, T, H, dh, p_drop, dbits, S(stream));
  } else {
    mar_set_engine(MAR_ENGINE_SIMT);
    rc = attention_bwd_simt(qkv, key_mask, out, dout, lse, delta, dqkv, B, T, H, dh, dtype, p_drop, dbits, S(stream));
  }
  if (rc == MAR_OK && dqkv_colsum != nullptr)     // the other engines sum dqkv's columns in a pass of their own
    rc = mar_linear_bwd_epilogue(dqkv, nullptr, nullptr, dqkv_colsum, B * T, 3 * H * dh, dtype, dtype, 0, 0.f, nullptr, 0, 0, stream);
  return rc;
}

int64_t mar_attention_bwd_work_floats(int64_t B, int64_t T, int64_t H, int64_t dh) {
  if (B <= 0 || T <= 0 || H <= 0 || dh <= 0) return 0;
  return attention_bwd_tc_work_floats(B, T, H, dh);   // >= B*H*T, which is all the other engines need
}

// ------------------------------------------------------------------------------------------
int64_t mar_gru_work_floats(int64_t B, int64_t T, int64_t H) { (void)T; return B * 5 * H; }

int mar_gru_fwd(const void* gi, const void* w_hh, const float* b_hh, void* hseq, void* hprev, float* saved, float* work,
                int64_t B, int64_t T, int64_t H, int dtype, int engine, void* stream) {
  MAR_CHECK_ARG(gi && w_hh && b_hh && hseq && work, "mar_gru_fwd: null pointer");
  MAR_CHECK_ARG((saved == nullptr) == (hprev == nullptr), "mar_gru_fwd: saved and hprev go together");
  MAR_CHECK_ARG(B > 0 && T > 0 && H > 0, "mar_gru_fwd: bad shape");
  MAR_CHECK_ARG(dtype == MAR_F32 || dtype == MAR_BF16, "mar_gru_fwd: bad dtype");
  const bool pers_ok = gru_persistent_supported(B, T, H, dtype);
  if (engine == MAR_ENGINE_TCGEN05 && !pers_ok) MAR_UNSUPPORTED("mar_gru_fwd: persistent engine cannot take B=%lld T=%lld H=%lld dtype=%d", (long long)B, (long long)T, (long long)H, dtype);
  if (engine == MAR_ENGINE_TCGEN05 || (engine == MAR_ENGINE_AUTO && pers_ok && !env_flag("MAR_FORCE_SIMT"))) {
    mar_set_engine(MAR_ENGINE_TCGEN05);
    return gru_fwd_persistent(gi, w_hh, b_hh, hseq, hprev, saved, B, T, H, S(stream));
  }
  mar_set_engine(MAR_ENGINE_SIMT);
  return gru_fwd_steps(gi, w_hh, b_hh, hseq, hprev, saved, work, B, T, H, dtype, S(stream));
}

int mar_gru_bwd(const void* dhseq, const float* saved, const void* w_hh, void* dgi, void* dgh, float* work, int64_t B,
                int64_t T, int64_t H, int dtype, int engine, void* stream) {
  MAR_CHECK_ARG(dhseq && saved && w_hh && dgi && dgh && work, "mar_gru_bwd: null pointer");
  MAR_CHECK_ARG(B > 0 && T > 0 && H > 0, "mar_gru_bwd: bad shape");
  MAR_CHECK_ARG(dtype == MAR_F32 || dtype == MAR_BF16, "mar_gru_bwd: bad dtype");
  const bool pers_ok = gru_persistent_supported(B, T, H, dtype);
  if (engine == MAR_ENGINE_TCGEN05 && !pers_ok) MAR_UNSUPPORTED("mar_gru_bwd: persistent engine cannot take B=%lld T=%lld H=%lld dtype=%d", (long long)B, (long long)T, (long long)H, dtype);
  if (engine == MAR_ENGINE_TCGEN05 || (engine == MAR_ENGINE_AUTO && pers_ok && !env_flag("MAR_FORCE_SIMT"))) {
    mar_set_engine(MAR_ENGINE_TCGEN05);
    return gru_bwd_persistent(dhseq, saved, w_hh, dgi, dgh, B, T, H, S(stream));
  }
  mar_set_engine(MAR_ENGINE_SIMT);
  return gru_bwd_steps(dhseq, saved, w_hh, dgi, dgh, work, B, T, H, dtype, S(stream));
}

int mar_lstm_fwd(const void* gi, const void* w_hh, const float* b_hh, void* hseq, void* hprev, float* saved, float* work,
                 int64_t B, int64_t T, int64_t H, int dtype, int engine, void* stream) {
  MAR_CHECK_ARG(gi && w_hh && b_hh && hseq && work, "mar_lstm_fwd: null pointer");
  MAR_CHECK_ARG((saved == nullptr) == (hprev == nullptr), "mar_lstm_fwd: saved and hprev go together");
  MAR_CHECK_ARG(B > 0 && T > 0 && H > 0, "mar_lstm_fwd: bad shape");
  MAR_CHECK_ARG(dtype == MAR_F32 || dtype == MAR_BF16, "mar_lstm_fwd: bad dtype");
  const bool pers_ok = gru_persistent_supported(B, T, H, dtype);
  if (engine == MAR_ENGINE_TCGEN05 && !pers_ok) MAR_UNSUPPORTED("mar_lstm_fwd: persistent engine cannot take B=%lld T=%lld H=%lld dtype=%d", (long long)B, (long long)T, (long long)H, dtype);
  if (engine == MAR_ENGINE_TCGEN05 || (engine == MAR_ENGINE_AUTO && pers_ok && !env_flag("MAR_FORCE_SIMT"))) {
    mar_set_engine(MAR_ENGINE_TCGEN05);
    return lstm_fwd_persistent(gi, w_hh, b_hh, hseq, hprev, saved, B, T, H, S(stream));
  }
  mar_set_engine(MAR_ENGINE_SIMT);
  return lstm_fwd_steps(gi, w_hh, b_hh, hseq, hprev, saved, work, B, T, H, dtype, S(stream));
}

int mar_lstm_bwd(const void* dhseq, const float* saved, const void* w_hh, void* dgates, float* work, int64_t B,
                 int64_t T, int64_t H, int dtype, int engine, void* stream) {
  MAR_CHECK_ARG(dhseq && saved && w_hh && dgates && work, "mar_lstm_bwd: null pointer");
  MAR_CHECK_ARG(B > 0 && T > 0 && H > 0, "mar_lstm_bwd: bad shape");
  MAR_CHECK_ARG(dtype == MAR_F32 || dtype == MAR_BF16, "mar_lstm_bwd: bad dtype");
  const bool pers_ok = gru_persistent_supported(B, T, H, dtype);
  if (engine == MAR_ENGINE_TCGEN05 && !pers_ok) MAR_UNSUPPORTED("mar_lstm_bwd: persistent engine cannot take B=%lld T=%lld H=%lld dtype=%d", (long long)B, (long long)T, (long long)H, dtype);
  if (engine == MAR_ENGINE_TCGEN05 || (engine == MAR_ENGINE_AUTO && pers_ok && !env_flag("MAR_FORCE_SIMT"))) {
    mar_set_engine(MAR_ENGINE_TCGEN05);
    return lstm_bwd_persistent(dhseq, saved, w_hh, dgates, B, T, H, S(stream));
  }
  mar_set_engine(MAR_ENGINE_SIMT);
  return lstm_bwd_steps(dhseq, saved, w_hh, dgates, work, B, T, H, dtype, S(stream));
}

}  // extern "C"
