/*
 * mar.h — C ABI of libmar.so, the sm_100a kernel library behind the drop-in sequence-classifier
 * modules (multimodalaggressionrecognition_b200/models.py).
 *
 * The reference (cafe1930/MultimodalAggressionRecognition) has no native code and no FFI: every
 * device op on its hot path is an implicit torch.nn library call made from models.py
 * (SURVEY.md §2.2).  Each entry point below therefore cites the reference call site (file:line in
 * /root/reference) and the torch op it stands in for; INTEGRATION.md shows the ctypes binding a
 * reference maintainer would add.
 *
 * Conventions
 *  - Plain pointers and sizes only; every pointer is DEVICE memory owned by the caller
 *    (PyTorch's caching allocator on the Python side) and must outlive the call in stream order.
 *  - Every function enqueues work on `stream` (a cudaStream_t passed as void*) and returns at
 *    once; no hidden cudaDeviceSynchronize, safe under CUDA-graph capture.
 *  - Return value: 0 = OK, negative = error (MAR_ERR_*); the message is in mar_last_error()
 *    (thread-local).  Nothing throws across the ABI and nothing calls exit().
 *  - dtype: MAR_F32 or MAR_BF16 is the storage type of ACTIVATIONS and WEIGHTS handed in; all
 *    accumulation, statistics, losses and parameter gradients of reductions are fp32.
 *  - Row-major everywhere.  A "linear" weight is (N,K) = (out_features, in_features) exactly as
 *    nn.Linear stores it.
 *  - Dropout masks are a pure function of (seed, step) held in a 2×uint64 DEVICE buffer
 *    `rng_state`, a per-call `site` id and the element index, so backward regenerates the mask
 *    instead of storing it and a captured CUDA graph draws fresh masks on every replay
 *    (mar_rng_advance bumps `step` on the device).
 */
#ifndef MAR_H_
#define MAR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAR_VERSION 101

enum { MAR_F32 = 0, MAR_BF16 = 1 };

enum {
  MAR_OK = 0,
  MAR_ERR_INVALID = -1,     /* bad shape / null pointer / misaligned pointer */
  MAR_ERR_UNSUPPORTED = -2, /* configuration the kernels do not cover */
  MAR_ERR_CUDA = -3,        /* CUDA runtime / driver error (message has the cudaError string) */
  MAR_ERR_ARCH = -4         /* device is not compute capability 10.x */
};

/* GEMM engine selector.  AUTO picks tcgen05 for bf16 problems that meet its alignment rules and the
 * SIMT kernel otherwise (all fp32 problems: fp32 mode must hold 1e-4, which rules out TF32). */
enum { MAR_ENGINE_AUTO = 0, MAR_ENGINE_SIMT = 1, MAR_ENGINE_TCGEN05 = 2 };

/* Epilogue flags of mar_linear_fwd:  out = residual + relu_post(dropout(relu_pre(x·Wᵀ + bias))) */
enum {
  MAR_EPI_RELU_PRE = 1,   /* ReLU before dropout  (FFN linear1, EmbeddingLayer, MLP heads)     */
  MAR_EPI_DROPOUT = 2,    /* inverted dropout with probability p                               */
  MAR_EPI_RELU_POST = 4   /* ReLU after dropout   (adaptors: Linear→Dropout→ReLU, models.py:743-748) */
};

int mar_version(void);
const char* mar_last_error(void);
/* sm_count / cc_major / cc_minor of the current device. */
int mar_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* Number of kernels this library launched in this process since load / since the last reset
 * (bench.py's `gpu_launches`). */
int64_t mar_launch_count(void);
void mar_launch_count_reset(void);
/* Which engine the most recent mar_linear_* call on this thread used (MAR_ENGINE_*). */
int mar_last_engine(void);

/* ---- RNG state (dropout) -------------------------------------------------------------------- */
int mar_rng_init(uint64_t* rng_state, uint64_t seed, uint64_t step, void* stream);
int mar_rng_advance(uint64_t* rng_state, void* stream);

/* ---- Dense contractions --------------------------------------------------------------------- */
/* nn.Linear forward with fused epilogue.  Replaces F.linear at models.py:143 (EmbeddingLayer),
 * :383-386, :115-118 (MLP heads), :693/:744 (adaptors), :711-714 (aggression heads) and the
 * in_proj / out_proj / linear1 / linear2 GEMMs of nn.TransformerEncoderLayer built at
 * models.py:348, :398 (torch/nn/functional.py:5798, :6690; transformer.py:980-982).
 * x (M,K) row stride ldx; w (N,K) contiguous; bias (N) fp32 or NULL; residual (M,N) row stride
 * ldr or NULL; out (M,N) row stride ldo.  in_dtype is the type of x and w (and residual);
 * out_dtype the type of out. */
int mar_linear_fwd(const void* x, int64_t ldx, const void* w, const float* bias,
                   const void* residual, int64_t ldr, void* out, int64_t ldo,
                   int64_t M, int64_t N, int64_t K, int in_dtype, int out_dtype,
                   int flags, float p_drop, const uint64_t* rng_state, uint32_t site,
                   int engine, void* stream);

/* Backward through the epilogue: dz = dout ⊙ d(epilogue)/dz, and dbias += column-sum(dz) (fp32,
 * accumulated; pass NULL to skip).  `out` is the forward output (needed for the ReLU flags, may be
 * NULL otherwise).  dz may alias dout.  All of (M,N), contiguous.
 * pooled_rows = T > 0: the epilogue's output went through a mean over T consecutive rows (SequenceAverageFeatures
 * after an adaptor, models.py:693-699, :105): dout is the POOLED gradient (M/T, N) and row r of the result reads
 * dout row r / T scaled by 1/T — the broadcast is never materialised. */
int mar_linear_bwd_epilogue(const void* dout, const void* out, void* dz, float* dbias,
                            int64_t M, int64_t N, int dtype, int out_dtype, int flags, float p_drop,
                            const uint64_t* rng_state, uint32_t site, int64_t pooled_rows, void* stream);

/* dx = dz·W (+ add).  dz (M,N); w (N,K); wt (K,N) = Wᵀ in the same dtype, required by the tcgen05
 * engine (K-major B operand), may be NULL for SIMT; add (M,K) or NULL; dx (M,K) row stride lddx.
 * act (M,K) row stride lddx or NULL (not together with add): the forward VALUE of this linear's input when that input
 * came out of a ReLU(+dropout) epilogue — zero exactly where the activation's derivative is zero — so that
 * dx = (dz·W) ⊙ (act > 0 ? act_scale : 0) leaves the GEMM with the producer's activation backward already applied
 * (FFN linear2 → linear1 of transformer.py:980-982: the (M, d_ff) gradient is written once, masked).
 * dx_colsum (K) fp32 or NULL (needs lddx == K): += column sums of dx — when dx already IS the producer's dz (act given),
 * that is the producer's bias gradient, taken from the accumulator tiles in the GEMM's own epilogue instead of by a
 * pass over the (M, d_ff) tensor. */
int mar_linear_dgrad(const void* dz, const void* w, const void* wt, const void* add, const void* act, float act_scale,
                     void* dx, int64_t lddx, float* dx_colsum, int64_t M, int64_t N, int64_t K, int dtype, int engine,
                     void* stream);

/* dw (N,K) fp32 (+)= dzᵀ·x.  dz (M,N) contiguous, x (M,K) row stride ldx.  accumulate=0 overwrites. */
int mar_linear_wgrad(const void* dz, const void* x, int64_t ldx, float* dw, int64_t M, int64_t N,
                     int64_t K, int dtype, int accumulate, int engine, void* stream);

/* fp32 master weight (N,K) → compute-dtype copy w (N,K) and, if wt != NULL, its transpose (K,N). */
int mar_cast_weight(const float* src, void* w, void* wt, int64_t N, int64_t K, int dtype, void* stream);
/* elementwise cast between fp32 and bf16 buffers of n elements */
int mar_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, void* stream);

/* ---- Attention ------------------------------------------------------------------------------ */
/* softmax(QKᵀ/√dh + key mask)·V over the packed in-projection output qkv (B,T,3d) (rows of each
 * token: [Q(d) | K(d) | V(d)], head h at columns h*dh..).  Replaces F.scaled_dot_product_attention
 * (torch/nn/functional.py:6682) reached from models.py:365 (no mask) and :425 (key padding mask
 * from the zero-row rule).  key_mask (B,T) uint8, 1 = key ignored, or NULL.  out (B,T,d).
 * lse (B,H,T) fp32 log-sum-exp of the scaled scores (-inf for a fully masked row, whose output
 * row is 0 — torch 2.11 safe-softmax behaviour).  Dropout on P with p_drop when > 0. */
int mar_attention_fwd(const void* qkv, const uint8_t* key_mask, void* out, float* lse,
                      int64_t B, int64_t T, int64_t H, int64_t dh, int dtype, float p_drop,
                      const uint64_t* rng_state, uint32_t site, uint32_t* drop_bits, int engine, void* stream);
/* Dropout on P: the call draws ONE keep bit per (b, h, query, key) from (rng_state, site) into `drop_bits`
 * (mar_attention_dropbits_words() 32-bit words, 16 B aligned, caller-owned; may be NULL when p_drop == 0) before the
 * attention kernel runs: word (bh*T + q)*W + (k >> 5), bit k & 31, W = 4*ceil(T/128).  Every engine reads the same
 * bits and mar_attention_bwd takes the same buffer, so the backward mask IS the forward mask.  The keep probability
 * is quantised to m/256, m = round((1 - p_drop)*256) (the granularity of torch's fused SDPA kernels); kept scores are
 * scaled by 256/m. */
int64_t mar_attention_dropbits_words(int64_t B, int64_t T, int64_t H);
/* dqkv (B,T,3d) from dout (B,T,d).  work: mar_attention_bwd_work_floats() floats of 16 B-aligned scratch
 * (delta = rowsum(dO ⊙ O) (B,H,T), then the tcgen05 engine's fp32 dQ accumulator (B,T,d) when T > 128).
 * drop_bits: the buffer the forward call filled (NULL when p_drop == 0).
 * dqkv_colsum (3d) fp32 or NULL: += column sums of dqkv over all B*T tokens = the gradient of the in-projection bias
 * (functional.py:5798); the tcgen05 engine takes them from its dK / dV / dQ accumulators before they are written. */
int64_t mar_attention_bwd_work_floats(int64_t B, int64_t T, int64_t H, int64_t dh);
int mar_attention_bwd(const void* qkv, const uint8_t* key_mask, const void* out, const void* dout,
                      const float* lse, float* work, void* dqkv, float* dqkv_colsum, int64_t B, int64_t T, int64_t H,
                      int64_t dh, int dtype, float p_drop, const uint32_t* drop_bits, int engine, void* stream);

/* ---- LayerNorm ------------------------------------------------------------------------------ */
/* y = LN(x)·gamma + beta over the last dim D, eps inside the sqrt, biased variance
 * (nn.LayerNorm at models.py:352, :403; norm1/norm2 at transformer.py:953-956).  The residual add
 * is fused into the producing GEMM's epilogue.  mean/rstd (rows) fp32 are saved for backward
 * (may be NULL in inference).  If zero_rows != NULL (rows, uint8) rows flagged 1 are treated as
 * x = 0 (the eval-mode nested-tensor zero fill, transformer.py:547-548). */
int mar_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean,
                      float* rstd, const uint8_t* zero_rows, int64_t rows, int64_t D, float eps,
                      int dtype, void* stream);
/* dx, and dgamma/dbeta (fp32, ACCUMULATED into). */
int mar_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd,
                      const float* gamma, void* dx, float* dgamma, float* dbeta, int64_t rows,
                      int64_t D, int dtype, void* stream);

/* mar_layernorm_bwd for a LayerNorm whose input was  x = residual + dropout_p(z)  written by mar_linear_fwd with
 * MAR_EPI_DROPOUT (out_proj -> norm1 and linear2 -> norm2 of the post-norm encoder layer, transformer.py:953-956): besides
 * dx it writes dz = dx ⊙ keep/(1-p) (rows, D) — the gradient that linear's dgrad / wgrad consume, mask regenerated from
 * (rng_state, site) exactly as mar_linear_bwd_epilogue would — and accumulates dbias (D fp32, may be NULL) += column
 * sums of dz: the dropout-backward pass over the tensor is folded into the pass LayerNorm's backward makes anyway. */
int mar_layernorm_bwd_dropout(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                              void* dx, float* dgamma, float* dbeta, void* dz, float* dbias, int64_t rows, int64_t D,
                              int dtype, float p_drop, const uint64_t* rng_state, uint32_t site, void* stream);

/* The same two kernels with a ROW MAP on the output (forward: y) / incoming-gradient (backward: dy) side, so that
 * torch.cat along T (models.py:419) and the per-modality slices (models.py:430) cost no copy.  Rows are (b, t) =
 * (row / Tin, row % Tin); segment k covers t in [t0[k], t1[k]) — the segments tile [0, Tin) in order, 1..4 of them —
 * and places the row at  ptrs[k] + ((b * Tout[k]) + tout0[k] + (t - t0[k])) * D  (16 B aligned bases):
 *  - an extractor's final LayerNorm writes straight into its slice of the fused (B, T_a+T_v, d) sequence:
 *    one segment, Tout = T_a+T_v, tout0 = the modality's offset;
 *  - the fusion encoder's final LayerNorm writes one contiguous (B, T_k, d) tensor per modality: Tin = T_a+T_v,
 *    one segment per modality with Tout = T_k, tout0 = 0.
 * nseg = 0 is the plain call (y / dy used).  Pointer and index arrays are HOST arrays read during the call. */
int mar_layernorm_fwd_mapped(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                             const uint8_t* zero_rows, int64_t rows, int64_t D, float eps, int dtype, int nseg, int64_t Tin,
                             const int64_t* t0, const int64_t* t1, const int64_t* Tout, const int64_t* tout0,
                             void* const* y_ptrs, void* stream);
int mar_layernorm_bwd_mapped(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                             void* dx, float* dgamma, float* dbeta, int64_t rows, int64_t D, int dtype, int nseg, int64_t Tin,
                             const int64_t* t0, const int64_t* t1, const int64_t* Tout, const int64_t* tout0,
                             void* const* dy_ptrs, void* stream);

/* ---- Pooling / masks ------------------------------------------------------------------------ */
/* out (B,D) = mean over T of x (B,T,D) (SequenceAverageFeatures, models.py:105; :97). */
int mar_meanpool_fwd(const void* x, void* out, int64_t B, int64_t T, int64_t D, int dtype, void* stream);
int mar_meanpool_bwd(const void* dout, void* dx, int64_t B, int64_t T, int64_t D, int dtype, void* stream);
/* mask (rows) uint8 = (sum over D of x == 0) (models.py:421-422). */
int mar_rowzero_mask(const void* x, uint8_t* mask, int64_t rows, int64_t D, int dtype, void* stream);
/* Copy a (B,T,D) block into columns [t_off, t_off+T) of a (B,T_total,D) buffer and back
 * (torch.cat along T at models.py:419 and the slices at :430). */
int mar_concat_rows(const void* src, void* dst, int64_t B, int64_t T, int64_t T_total, int64_t t_off,
                    int64_t D, int dtype, int to_concat, void* stream);

/* ---- Classifier loss ------------------------------------------------------------------------ */
/* nn.CrossEntropyLoss (mean, optional class weights) on fp32 logits (B,C), int64 labels; rows
 * with label < 0 are ignored (EMPTY samples, models.py:247-253).  loss: 1 float.  dlogits (B,C)
 * = d loss / d logits (the caller scales by the upstream grad).  preds (B) int64 argmax or NULL
 * (trainer.py:170, :726). */
int mar_cross_entropy_fwd(const float* logits, const int64_t* labels, const float* class_weight,
                          float* loss, float* dlogits, int64_t* preds, int64_t B, int64_t C, void* stream);

/* Multi-class focal loss, the criterion train_multimodal.py:494-510 loads from torch.hub
 * (adeelh/pytorch-multi-class-focal-loss, unpinned and absent offline: restated from its published algorithm,
 * parity unpinned): per row  -alpha[y]·(1-p_y)^gamma·log p_y, 'mean' = plain mean over rows with label >= 0
 * (train_multimodal.py passes the class weights as alpha and gamma = 1.5..2).  alpha (C) fp32 or NULL. */
int mar_focal_loss_fwd(const float* logits, const int64_t* labels, const float* alpha, float gamma,
                       float* loss, float* dlogits, int64_t B, int64_t C, void* stream);

/* ---- GRU / LSTM recurrence ------------------------------------------------------------------ */
/* One-layer batch_first GRU, h0 = 0 (nn.GRU at models.py:110,122; gate order r,z,n).
 * gi (B,T,3H) = x·W_ihᵀ + b_ih is produced by mar_linear_fwd.  w_hh (3H,H) in `dtype`, b_hh fp32.
 * hseq (B,T,H): output sequence in `dtype`.  The hidden state is carried in fp32 between steps; only the
 * recurrent GEMM operand is rounded to `dtype`.  For training (NULL in inference):
 *   saved (B,T,5H) fp32: r, z, n, (W_hn h + b_hn), h_{t-1};   hprev (B,T,H) in `dtype`: h_{t-1} (wgrad operand).
 * work: mar_gru_work_floats() floats. */
int mar_gru_fwd(const void* gi, const void* w_hh, const float* b_hh, void* hseq, void* hprev, float* saved,
                float* work, int64_t B, int64_t T, int64_t H, int dtype, int engine, void* stream);
/* dhseq (B,T,H): incoming grad of every step's output (zeros except the last step for the reference's heads).
 * Produces dgi (B,T,3H) and dgh (B,T,3H) in `dtype`; dW_ih, dW_hh, the biases and dx then follow from
 * mar_linear_wgrad / mar_linear_dgrad / mar_linear_bwd_epilogue on those. */
int mar_gru_bwd(const void* dhseq, const float* saved, const void* w_hh, void* dgi, void* dgh, float* work,
                int64_t B, int64_t T, int64_t H, int dtype, int engine, void* stream);
int64_t mar_gru_work_floats(int64_t B, int64_t T, int64_t H);

/* LSTM (train_video_rnn.py:94-106; gate order i,f,g,o).  saved (B,T,5H) fp32: i,f,g,o,c; hprev as above. */
int mar_lstm_fwd(const void* gi, const void* w_hh, const float* b_hh, void* hseq, void* hprev, float* saved,
                 float* work, int64_t B, int64_t T, int64_t H, int dtype, int engine, void* stream);
int mar_lstm_bwd(const void* dhseq, const float* saved, const void* w_hh, void* dgates, float* work,
                 int64_t B, int64_t T, int64_t H, int dtype, int engine, void* stream);

/* ---- Optimizer ------------------------------------------------------------------------------ */
/* torch.optim.Adam defaults (train_multimodal.py:444) over one flat fp32 buffer of n elements;
 * `step_dev` (1 float on the device, incremented by the kernel's caller via mar_adam_tick) keeps
 * the bias-correction step so the update is graph-capturable. */
int mar_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, const float* step_dev,
                  int64_t n, float lr, float beta1, float beta2, float eps, void* stream);
int mar_adam_tick(float* step_dev, void* stream);
/* torch.optim.Adam as the reference drives it (train_multimodal.py:444 + trainer.py:140-150: zero_grad leaves the
 * gradients of an inactive head / branch None, and its AggrBatchSampler makes every batch homogeneous in aggression
 * type, datasets.py:630-645): a parameter WITHOUT a gradient this step is skipped — no moment decay, no update, no
 * step-count increment — and every parameter carries its own step count.  One segment per parameter over the flat
 * buffers; every segment starts on a multiple of `chunk` floats (a power of two); chunk_seg (n / chunk, int32) names
 * the segment of each chunk (< 0: padding); seg_active (nseg fp32) > 0 marks the parameters that received a gradient;
 * seg_steps (nseg fp32) are the per-parameter step counts (advanced here); seg_coef (2·nseg fp32, 8 B aligned) is
 * scratch.  bf16_mirror (n bf16, 8 B aligned, or NULL): the compute-dtype copy of the flat parameter buffer, rewritten
 * wherever a parameter is updated, so that no per-step weight cast is needed.  Two launches, graph-capturable. */
int mar_adam_step_segments(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                           const int32_t* chunk_seg, float* seg_steps, const float* seg_active, float* seg_coef,
                           int64_t n, int chunk, int nseg, float lr, float beta1, float beta2, float eps, void* bf16_mirror,
                           void* stream);
/* out[0] = sum over rows with 0 <= label < C of class_weight[label] (1 per row when class_weight is NULL): the
 * denominator of nn.CrossEntropyLoss(reduction='mean') (models.py:232-263).  Data-parallel ranks exchange it to weigh
 * their gradients so that the reduced gradient is the one of the loss over the GLOBAL batch. */
int mar_label_weight_sum(const int64_t* labels, const float* class_weight, float* out, int64_t B, int64_t C, void* stream);

/* ---- Data-parallel exchange fused with the optimizer (one process per GPU, one node, NVLink peer memory) --------
 * The reference has no distributed code; data parallelism over clips is this repo's addition
 * (§8e) and its only exchange is the parameter-gradient sum in front of torch.optim.Adam (train_multimodal.py:444).
 * mar_dp_allreduce_adam does that exchange AND the Adam step in ONE kernel: fp32 gradients -> bf16 wire copy ->
 * reduce-scatter by peer loads -> all-gather by peer stores -> per-parameter Adam (semantics of
 * mar_adam_step_segments; seg_active is the reduced header of the gradient buffer: a parameter is active if ANY rank
 * saw a gradient) -> bf16 mirror.  Every rank applies the same bf16-rounded sums: parameters stay bit-identical.
 *
 * Symmetric memory: every rank allocates one block of mar_dp_block_bytes(n) with mar_peer_alloc (cudaMalloc, zeroed),
 * exports it (64-byte CUDA IPC handle), the handles travel over the caller's own channel (torch.distributed here) and
 * each rank imports its peers' blocks.  `blocks` is a HOST array of `world` device pointers in rank order (own block
 * included).  `ctrl` is mar_dp_ctrl_bytes() of zeroed LOCAL device memory that must persist between calls (the
 * exchange counter lives there, so a captured CUDA graph replays).  seg_steps / chunk_seg / chunk / nseg as in
 * mar_adam_step_segments (chunk >= 8).  write_back != 0: the reduced gradient is also written back into `grad`.
 * All ranks must make the same sequence of calls.  A peer that never arrives makes the kernel give up after a
 * time-out instead of hanging the GPU; mar_dp_check (synchronises `stream`) reports it. */
int64_t mar_dp_block_bytes(int64_t n);
int64_t mar_dp_ctrl_bytes(void);
int mar_peer_alloc(void** ptr, int64_t bytes);
int mar_peer_free(void* ptr);
int mar_peer_export(void* ptr, void* handle64);
int mar_peer_import(const void* handle64, void** ptr);
int mar_peer_close(void* ptr);
int mar_dp_allreduce_adam(float* grad, int write_back, int64_t n, int world, int rank, void* const* blocks, void* ctrl,
                          float* param, float* exp_avg, float* exp_avg_sq, const int32_t* chunk_seg, float* seg_steps,
                          int chunk, int nseg, float lr, float beta1, float beta2, float eps, void* bf16_mirror,
                          void* stream);
int mar_dp_check(void* ctrl, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MAR_H_ */
