// SIMT fp32-accumulate GEMM with the fused linear epilogue.
//
// This is the fp32-mode engine (fp32 parity must hold 1e-4 against the CPU oracle, which rules out
// TF32 tensor cores) and the engine for shapes the tcgen05 kernel does not take (tiny N such as the
// 2-class classifier, unaligned leading dimensions).  Operands are addressed by (row, col) strides
// so that forward (x·Wᵀ), dgrad (dz·W) and wgrad (dzᵀ·x) all run through the same kernel.
#include "common.cuh"
#include "gemm_simt.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, TM = 8, TN = 8, PAD = 4;

template <typename TA, typename TB, typename TC>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const TA* __restrict__ A, int64_t sam, int64_t sak,
                 const TB* __restrict__ B, int64_t sbk, int64_t sbn,
                 TC* __restrict__ C, int64_t ldc, int M, int N, int K, int k_per_split, SimtEpilogue epi) {
  __shared__ float As[BK][BM + PAD];
  __shared__ float Bs[BK][BN + PAD];
  const int t = threadIdx.x;
  const int ty = t / 16, tx = t % 16;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; i++)
#pragma unroll
    for (int j = 0; j < TN; j++) acc[i][j] = 0.f;

  const bool a_kcontig = (sak == 1);
  const bool b_ncontig = (sbn == 1);
  float ra[8], rb[8];

  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      int e = t + 256 * i;
      int kk = a_kcontig ? (e % BK) : (e / BM);
      int mm = a_kcontig ? (e / BK) : (e % BM);
      int gm = m0 + mm, gk = k0 + kk;
      ra[i] = (gm < M && gk < kend) ? to_f32<TA>(A[(int64_t)gm * sam + (int64_t)gk * sak]) : 0.f;
      int kb = b_ncontig ? (e / BN) : (e % BK);
      int nb = b_ncontig ? (e % BN) : (e / BK);
      int gn = n0 + nb;
      gk = k0 + kb;
      rb[i] = (gn < N && gk < kend) ? to_f32<TB>(B[(int64_t)gk * sbk + (int64_t)gn * sbn]) : 0.f;
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      int e = t + 256 * i;
      int kk = a_kcontig ? (e % BK) : (e / BM);
      int mm = a_kcontig ? (e / BK) : (e % BM);
      As[kk][mm] = ra[i];
      int kb = b_ncontig ? (e / BN) : (e % BK);
      int nb = b_ncontig ? (e % BN) : (e / BK);
      Bs[kb][nb] = rb[i];
    }
  };

  if (kbeg < kend) load_tiles(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    store_tiles();
    __syncthreads();
    if (k0 + BK < kend) load_tiles(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; kk++) {
      float a[TM], b[TN];
      float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * TM]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * TM + 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
      for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  DropKey dk;
  const bool do_drop = (epi.flags & MAR_EPI_DROPOUT) && epi.p > 0.f;
  if (do_drop) dk = make_drop_key(epi.rng, epi.site, epi.p);
  const bool first_split = (blockIdx.z == 0);

#pragma unroll
  for (int i = 0; i < TM; i++) {
    int gm = m0 + ty * TM + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; j++) {
      int gn = n0 + tx * TN + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (epi.atomic) {
        // split-K partial: bias / residual belong to the first split only; no nonlinear epilogue here
        if (first_split && epi.bias) v += epi.bias[gn];
        atomicAdd(reinterpret_cast<float*>(C) + (int64_t)gm * ldc + gn, v);
        continue;
      }
      if (epi.bias) v += epi.bias[gn];
      if (epi.flags & MAR_EPI_RELU_PRE) v = fmaxf(v, 0.f);
      if (do_drop) v = drop_keep(dk, (uint64_t)gm * (uint64_t)N + (uint64_t)gn) ? v * dk.scale : 0.f;
      if (epi.flags & MAR_EPI_RELU_POST) v = fmaxf(v, 0.f);
      if (epi.aux) {
        const float a = epi.res_is_bf16 ? __bfloat162float(reinterpret_cast<const bf16*>(epi.aux)[(int64_t)gm * epi.ldaux + gn])
                                        : reinterpret_cast<const float*>(epi.aux)[(int64_t)gm * epi.ldaux + gn];
        v = a > 0.f ? v * epi.aux_scale : 0.f;
      }
      if (epi.residual) {
        v += epi.res_is_bf16 ? __bfloat162float(reinterpret_cast<const bf16*>(epi.residual)[(int64_t)gm * epi.ldr + gn])
                             : reinterpret_cast<const float*>(epi.residual)[(int64_t)gm * epi.ldr + gn];
      }
      TC* cp = C + (int64_t)gm * ldc + gn;
      if (epi.accumulate) v += to_f32<TC>(*cp);
      *cp = from_f32<TC>(v);
    }
  }
}

template <typename TA, typename TB, typename TC>
int launch(const void* A, int64_t sam, int64_t sak, const void* B, int64_t sbk, int64_t sbn, void* C,
           int64_t ldc, int64_t M, int64_t N, int64_t K, const SimtEpilogue& epi_in, cudaStream_t st) {
  SimtEpilogue epi = epi_in;
  dim3 grid((unsigned)ceil_div(N, BN), (unsigned)ceil_div(M, BM), 1);
  int64_t k_per_split = K;
  // split-K only for plain fp32 (+)= outputs (wgrad): long reductions with few output tiles
  const bool can_split = sizeof(TC) == 4 && epi.flags == 0 && epi.residual == nullptr && epi.allow_split;
  if (can_split) {
    int64_t tiles = (int64_t)grid.x * grid.y;
    int64_t want = (2 * (int64_t)mar_sm_count() + tiles - 1) / tiles;
    int64_t max_split = ceil_div(K, 512);
    int64_t split = want < max_split ? want : max_split;
    if (split > 1) {
      k_per_split = ceil_div(ceil_div(K, split), BK) * BK;
      split = ceil_div(K, k_per_split);
      grid.z = (unsigned)split;
      epi.atomic = 1;
      if (!epi.accumulate) {
        if (ldc == N) {
          cudaMemsetAsync(C, 0, (size_t)M * N * 4, st);
        } else {
          cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, (size_t)M, st);
        }
      }
    }
  }
  gemm_simt_kernel<TA, TB, TC><<<grid, 256, 0, st>>>(
      reinterpret_cast<const TA*>(A), sam, sak, reinterpret_cast<const TB*>(B), sbk, sbn,
      reinterpret_cast<TC*>(C), ldc, (int)M, (int)N, (int)K, (int)k_per_split, epi);
  MAR_LAUNCH_CHECK("gemm_simt");
  return MAR_OK;
}

}  // namespace

int gemm_simt(const void* A, int a_dtype, int64_t sam, int64_t sak, const void* B, int b_dtype, int64_t sbk,
              int64_t sbn, void* C, int c_dtype, int64_t ldc, int64_t M, int64_t N, int64_t K,
              const SimtEpilogue& epi, cudaStream_t st) {
  if (M <= 0 || N <= 0) return MAR_OK;
  MAR_CHECK_ARG(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "gemm_simt: dimension too large");
  mar_set_engine(MAR_ENGINE_SIMT);
  const int key = a_dtype * 4 + b_dtype * 2 + c_dtype;
  switch (key) {
    case 0: return launch<float, float, float>(A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, epi, st);
    case 1: return launch<float, float, bf16>(A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, epi, st);
    case 6: return launch<bf16, bf16, float>(A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, epi, st);
    case 7: return launch<bf16, bf16, bf16>(A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, epi, st);
    case 2: return launch<float, bf16, float>(A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, epi, st);
    case 4: return launch<bf16, float, float>(A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, epi, st);
    default: MAR_UNSUPPORTED("gemm_simt: dtype combination a=%d b=%d c=%d", a_dtype, b_dtype, c_dtype);
  }
}
