#pragma once
#include "common.cuh"

// step-per-launch engine (rnn.cu).  saved: fp32 (B,T,5H); hprev: (B,T,H) in the activation dtype (h_{t-1}, the wgrad operand)
int gru_fwd_steps(const void* gi, const void* w_hh, const float* b_hh, void* hseq, void* hprev, float* saved, float* work,
                  int64_t B, int64_t T, int64_t H, int dtype, cudaStream_t st);
int gru_bwd_steps(const void* dhseq, const float* saved, const void* w_hh, void* dgi, void* dgh, float* work, int64_t B,
                  int64_t T, int64_t H, int dtype, cudaStream_t st);
int lstm_fwd_steps(const void* gi, const void* w_hh, const float* b_hh, void* hseq, void* hprev, float* saved, float* work,
                   int64_t B, int64_t T, int64_t H, int dtype, cudaStream_t st);
int lstm_bwd_steps(const void* dhseq, const float* saved, const void* w_hh, void* dgates, float* work, int64_t B,
                   int64_t T, int64_t H, int dtype, cudaStream_t st);

// persistent cluster engine (rnn_persistent.cu), bf16 only
bool gru_persistent_supported(int64_t B, int64_t T, int64_t H, int dtype);
int gru_fwd_persistent(const void* gi, const void* w_hh, const float* b_hh, void* hseq, void* hprev, float* saved, int64_t B,
                       int64_t T, int64_t H, cudaStream_t st);
int gru_bwd_persistent(const void* dhseq, const float* saved, const void* w_hh, void* dgi, void* dgh, int64_t B, int64_t T,
                       int64_t H, cudaStream_t st);
// LSTM on the same persistent cluster engine (4 gate blocks; saved = i,f,g,o,c)
int lstm_fwd_persistent(const void* gi, const void* w_hh, const float* b_hh, void* hseq, void* hprev, float* saved, int64_t B,
                        int64_t T, int64_t H, cudaStream_t st);
int lstm_bwd_persistent(const void* dhseq, const float* saved, const void* w_hh, void* dgates, int64_t B, int64_t T, int64_t H,
                        cudaStream_t st);
