// Internal launcher declarations shared between translation units (not part of the C ABI).
#pragma once
#include "common.cuh"

int k_im2col_k3(const bf16* x, bf16* out, int B, int T, int C, int stride, cudaStream_t st);
int k_layernorm_bf16(const bf16* x, const float* w, const float* b, bf16* y, long long rows, int D, float eps, cudaStream_t st);
int k_rmsnorm_f32(const float* x, const float* w, bf16* y, const int* row_index, long long rows, int D, float eps, cudaStream_t st,
                  long long ldy = 0);
int k_rmsnorm_f32_bwd(const bf16* dy, const float* x, const float* w, float* dx, const int* row_index, long long rows, int D,
                      float eps, int accumulate, cudaStream_t st, bf16* dx_bf16 = nullptr, long long ld_b = 0);
int k_enc_rope(bf16* qkv, const float* cosT, const float* sinT, long long rows, int S, int H, int hd, int rd, cudaStream_t st);
int k_lm_qknorm_rope_fwd(const bf16* qkv, bf16* qk, const float* qw, const float* kw, const float* cosT, const float* sinT,
                         long long M, int S, int Hq, int Hkv, float eps, cudaStream_t st, const int* pos_ids = nullptr);
int k_lm_qknorm_rope_bwd(const bf16* qkv, const float* dq, const bf16* dk, const bf16* dv, bf16* dqkv, const float* qw,
                         const float* kw, const float* cosT, const float* sinT, long long M, int S, int Hq, int Hkv, float eps,
                         cudaStream_t st, long long ld_out = 0);
int k_cast_rows_f32_bf16(const float* in, bf16* out, long long rows, int D, long long ld_out, cudaStream_t st);
int k_swiglu_h(const bf16* gu, bf16* h, long long M, int F, long long ld_h, cudaStream_t st);
int k_proj_norm_fwd(const bf16* x, const float* w, void* y, long long rows, int D, float eps, int gelu, cudaStream_t st);
int k_proj_norm_bwd(const bf16* x, const float* w, const void* dy, int dy_is_f32, bf16* dx, float* dw, long long rows, int D,
                    float eps, int gelu, cudaStream_t st);
int k_audio_index(const long long* ids, const long long* counts, int* src_row, int B, int S, int n_a, long long audio_id,
                  cudaStream_t st);
int k_embed_scatter(const long long* ids, const int* src_row, const float* table, const float* audio, float* out,
                    long long n_tok, int D, long long vocab, cudaStream_t st);
int k_audio_grad_gather(const int* src_row, const float* d_emb, float* d_audio, long long n_tok, int D, cudaStream_t st);
int k_ce_fwd_bwd(bf16* logits, long long ld, const int* targets, long long rows, int V, int Vpad, float inv_items,
                 float* loss_sum, float* row_loss, int write_grad, cudaStream_t st);
int k_transpose_bf16(const bf16* in, bf16* out, int R, int C, long long ld_in, long long ld_out, cudaStream_t st);
int k_cast_f32_bf16(const float* in, bf16* out, long long n, cudaStream_t st);
int k_frame_stack(const bf16* x, bf16* out, int B, int S, int n, int k, int D, cudaStream_t st);
int k_sumsq(const float* g, long long n, float* out, cudaStream_t st);

// unfrozen-LM recipe (lm_wgrad.cu)
int k_rmsnorm_dw(const bf16* dy, const float* x, const int* row_index, long long rows, int D, float eps, float* dw, cudaStream_t st);
int k_qknorm_dw(const bf16* qkv, const float* dq, const bf16* dk, const float* cosT, const float* sinT, long long M, int S, int Hq, int Hkv,
                float eps, float* dqw, float* dkw, cudaStream_t st);
int k_embed_grad_scatter(const long long* ids, const float* d_emb, float* d_table, long long n_tok, int D, long long vocab,
                         long long audio_id, cudaStream_t st);

// KV-cache decode path (decode.cu)
int k_skinny_gemm(const bf16* X, long long ldx, const bf16* W, long long ldw, int M, int N, int K, int mode, void* out, long long ldo,
                  const float* resid, cudaStream_t st, int k_splits = 1);
int k_skinny_splits(int N, int K);
int k_kv_cache_store(const bf16* k_src, long long k_ld, const bf16* v_src, long long v_ld, bf16* k_cache, bf16* v_cache, int B, int S,
                     int KD, int max_seq, cudaStream_t st);
int k_decode_attn(const bf16* qkv, const bf16* q_ready, bf16* k_cache, bf16* v_cache, bf16* out, long long ld_out, const float* qw,
                  const float* kw, const float* cosT, const float* sinT, const int* pos, int B, int Hq, int Hkv, int max_seq, float eps,
                  float scale, cudaStream_t st, const int* kv_start = nullptr);
int k_decode_resid_rmsnorm(const float* x_in, const float* partial, int n_splits, int rows, float* x_out, const float* w, bf16* y, int D,
                           float eps, long long ldy, cudaStream_t st);
int k_embed_rows(const long long* ids, const float* table, float* out, int B, int D, long long vocab, cudaStream_t st);
int k_argmax_rows(const bf16* logits, long long ld, int rows, int V, long long* next_ids, int* pos_inc, cudaStream_t st);

// 2-D bf16 row-major tensor map, box = {64 cols, box_rows}, SWIZZLE_128B, zero fill out of bounds (cached)
int k_make_tensor_map_2d(CUtensorMap* out, const void* ptr, long long rows, long long cols, long long ld, int box_rows);

// tcgen05 attention forward (attn_tc.cu): *handled = 1 when the shape is supported and the kernel was launched
int k_attn_tc_fwd(const bf16* q, const bf16* k, const bf16* v, bf16* o, float* lse, int B, int S, int Hq, int Hkv, int head_dim,
                  long long q_rs, long long k_rs, long long v_rs, long long o_rs, int causal, float scale, cudaStream_t st,
                  int* handled, const int* kv_start = nullptr);
int k_attn_tc_enabled();
int k_attn_tc_bwd(const bf16* q, const bf16* k, const bf16* v, const bf16* d_o, const float* lse, const float* dsum, float* dq_acc,
                  bf16* dk, bf16* dv, int B, int S, int Hq, int Hkv, int head_dim, long long q_rs, long long k_rs, long long v_rs,
                  long long do_rs, long long dq_rs, long long dk_rs, long long dv_rs, int causal, float scale, cudaStream_t st,
                  int* handled);
// attention backward with an optional precomputed D = rowsum(dO o O) (attn_mma.cu)
int k_attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse, float* dsum_ws, float* dq_acc,
               void* dk, void* dv, int B, int S, int Hq, int Hkv, int head_dim, long long q_rs, long long k_rs, long long v_rs, long long o_rs,
               long long do_rs, long long dq_rs, long long dk_rs, long long dv_rs, int causal, float scale, void* stream, int dsum_ready);
// weight-gradient GEMM (gemm_sm100.cu); out_zeroed: skip the clear that precedes a split-K launch
int k_gemm_bf16_tn(const void* At, long long ldat, const void* Bt, long long ldbt, int M, int N, int K, float* out, long long ldo, float alpha,
                   void* stream, int out_zeroed);
