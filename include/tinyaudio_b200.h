/* tinyaudio_b200.h -- C ABI of libtinyaudio_b200.so (sm_100a).
 *
 * Drop-in boundary for tiny-audio's training hot path.  The reference is pure Python and has no FFI of its own;
 * its boundary is the Python plugin surface (ASRModel / ASRProcessor / PROJECTOR_CLASSES).
 * Each entry point below replaces the arithmetic behind one reference call site (cited as file:line;
 * "HF:" = the transformers package the reference calls into).  The Python host side in tiny_audio_b200/ binds
 * these with ctypes and keeps the reference's class / method names.
 *
 * Conventions
 *   - every function returns 0 on success, negative on error; ta_last_error_string() gives the message
 *   - all pointers are DEVICE pointers unless stated; the caller (PyTorch) owns every buffer
 *   - no allocation, no host synchronisation, no retained pointers; launches go to `stream` (cudaStream_t)
 *   - bf16 = __nv_bfloat16 storage; "ld*" = leading dimension in ELEMENTS
 */
#ifndef TINYAUDIO_B200_H
#define TINYAUDIO_B200_H

#ifdef __cplusplus
extern "C" {
#endif

const char* ta_last_error_string(void);
int ta_version(void);
unsigned long long ta_launch_count(void); /* kernels launched by the library so far (host-side counter) */
/* programmatic dependent launch (griddepcontrol) for the training towers' kernels: compiled in only by `make PDL=1`; returns 1 when
 * the switch took effect, 0 in a default build (where the call is a no-op).  Experimental: see DESIGN.md section 7. */
int ta_set_pdl(int on);

/* ------------------------------------------------------------------------------------------------
 * GEMM  C[M,N] = epilogue(A[M,K] . B[N,K]^T), bf16 in, fp32 accumulate (tcgen05 + TMEM + TMA)
 * replaces every nn.Linear on the path: HF:models/glmasr/modeling_glmasr.py:198-206,223,228-239;
 * HF:models/qwen3/modeling_qwen3.py:81-83,263-291,505; tiny_audio/projectors.py:66-71
 * ---------------------------------------------------------------------------------------------- */
enum {
    TA_EPI_BF16 = 0,       /* out bf16 = acc + bias                                              */
    TA_EPI_BF16_GELU = 1,  /* out bf16 = gelu_erf(bf16(acc + bias))                              */
    TA_EPI_BF16_RESID = 2, /* out bf16 = resid_bf16 + bf16(acc + bias)                           */
    TA_EPI_F32_RESID = 3,  /* out f32  = resid_f32 + bf16(acc + bias)                            */
    TA_EPI_F32 = 4,        /* out f32  = alpha * acc                                             */
    TA_EPI_SWIGLU = 5,     /* B rows interleaved [64 gate | 64 up]: out bf16 [M,N/2] = silu(g)*u; out2 = (g,u) stash */
    TA_EPI_SWIGLU_BWD = 6, /* acc = d(h) [M,N]; aux = (g,u) stash [M,2N]; out bf16 [M,2N] = (d gate | d up)          */
    TA_EPI_BF16_ROPE = 7,  /* out bf16 = rope(acc + bias): GLM-ASR partial rotary on columns < rope_cols (64-wide heads, 32 dims) */
    TA_EPI_BF16_ROWDOT = 8 /* out bf16 = acc, and out2 fp32 [B, N/128, S] (S = rope_seq, row = b S + s) = sum over each 128-wide head of out * aux:
                              the attention backward's D = rowsum(dO o O) computed where dO is produced (HF sdpa backward).  N % 256 == 0; optional zero_f32 */
};

typedef struct ta_gemm_epilogue {
    void* out;
    long long ldo;
    const float* bias;  /* [N] fp32 or NULL */
    const void* resid;  /* RESID modes */
    long long ldr;      /* 0 -> ldo */
    void* out2;         /* SWIGLU: optional (gate, up) stash [M,N] bf16 */
    long long ldo2;
    const void* aux;    /* SWIGLU_BWD: the stash */
    long long ldaux;
    float alpha;        /* F32 mode scale; 0 -> 1 */
    const float* rope_cos; /* BF16_ROPE: fp32 [rope_seq, 16] */
    const float* rope_sin;
    int rope_seq;       /* position = row % rope_seq */
    int rope_cols;      /* rotate output columns [0, rope_cols) (the q and k blocks of a fused qkv projection) */
    void* zero_f32;     /* BF16_ROWDOT: optional fp32 [M, N] buffer (ld_zero elements per row) that the epilogue fills with ZEROS next to its
                           output -- the attention backward's dQ accumulator, cleared by a tensor-bound kernel's idle store bandwidth */
    long long ld_zero;
} ta_gemm_epilogue;

int ta_gemm_bf16(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K, int epilogue_mode,
                 const ta_gemm_epilogue* epilogue, void* stream);
/* weight-gradient form: C fp32 [M,N] = alpha * At^T . Bt with At bf16 [K,M], Bt bf16 [K,N] row-major (both operands MN-major UMMA
 * tiles straight from the activations; autograd's dW = dY^T X of every nn.Linear, HF:trainer.py:1935) */
int ta_gemm_bf16_tn(const void* At, long long ldat, const void* Bt, long long ldbt, int M, int N, int K, float* out, long long ldo,
                    float alpha, void* stream);
int ta_gemm_set_tile_n(int bn); /* 0 = auto, 128, 256 (testing / tuning) */
int ta_gemm_set_tail_split(int on);
/* ta_gemm_bf16_tn: split the token contraction of few-tile / deep-K weight-gradient products over all CTA pairs, partial tiles
 * reduce-added into the (zeroed) fp32 output by TMA.  0 = off (default).  Experimental, not yet run on hardware. */
int ta_gemm_set_tn_splitk(int on); /* 1: a mostly empty last wave of 256 x 256 tiles is issued as a second launch of 256 x 128 tiles (default 0:
                                       no gain under the 1 kW power cap, see gemm_sm100.cu) */
int ta_gemm_set_resid_tma(int on); /* bf16-residual epilogue (encoder o-projection / fc2): 1 = residual sub-tiles in by TMA, sum out of the same shared-memory buffer; 0 = per-thread global loads */
int ta_gemm_set_swiglu_bwd_tma(int on); /* 1 (default): SwiGLU-backward epilogue moves the (gate, up) stash and the gradients by TMA, in place in shared memory; 0: per-thread global loads */
int ta_gemm_set_cta_pair(int on); /* 1: CTA-pair kernel (tcgen05 cta_group::2, 256 x N tiles); 0: 1-CTA kernel */


/* ------------------------------------------------------------------------------------------------
 * a1. log-mel front end  (replaces WhisperFeatureExtractor._torch_extract_fbank_features,
 *     HF:models/whisper/feature_extraction_whisper.py:135-164; mel bank :95-103)
 *   wave (B, L) fp32 zero-padded clips; T = L / 160 frames.  workspace: ta_logmel_workspace_floats().
 *   out_f32 (B,128,T) fp32 "input_features" (optional); out_conv1_im2col bf16 [B*T, 384] (optional):
 *   row (b,t) = [mel(t-1) | mel(t) | mel(t+1)], the A operand of conv1 (k=3, pad=1) as a GEMM.
 * ---------------------------------------------------------------------------------------------- */
int ta_logmel_workspace_floats(int B, int L, long long* n_floats);
int ta_logmel_fwd(const float* wave, long long ld_wave, int B, int L, float* workspace, float* out_f32,
                  void* out_conv1_im2col_bf16, void* stream);
/* features computed by the reference's CPU collator (scripts/train.py:327-333) -> conv1 operand */
int ta_mel_to_conv1_im2col(const float* mel /*(B,128,T)*/, int B, int T, void* out_conv1_im2col_bf16, void* stream);

/* ------------------------------------------------------------------------------------------------
 * attention (flash, fp32 softmax).  q/k/v/o are [B, S, heads, head_dim] views with arbitrary row strides
 * (elements between consecutive tokens), so they can point into fused qkv buffers.
 *   encoder: HF:models/glmasr/modeling_glmasr.py:208-221 (non-causal, no mask, 20 x 64)
 *   decoder: HF:models/qwen3/modeling_qwen3.py:273-291 + HF:integrations/sdpa_attention.py (causal GQA 16/8 x 128)
 * lse: [B, Hq, S] fp32 (natural log), optional in forward.
 * backward: dsum_ws [B,Hq,S] fp32 scratch; dq_acc [B,S,Hq*hd] fp32 (zeroed inside); dk/dv bf16.
 * ---------------------------------------------------------------------------------------------- */
int ta_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int S, int Hq, int Hkv,
                int head_dim, long long q_rs, long long k_rs, long long v_rs, long long o_rs, int causal, float scale,
                void* stream);
int ta_debug_set(int key, int value); /* profiling experiments only; never changes results when left at 0 */
int ta_attn_set_tc(int mode); /* forward-kernel selection.  0: mma.sync reference kernels (the ONLY way to reach them: with any other mode a shape
                                 the tcgen05 kernels do not cover is an error, not a fallback); 1: tcgen05, two threads per query row;
                                 14 (default): tcgen05, encoder shape (head_dim 64, non-causal) on 64-key tiles with three CTAs per SM, other
                                 shapes as mode 1; 2-13, 15, 16: earlier / experimental encoder-shape kernels kept as A/B references
                                 (csrc/attn_tc.cu lists them with their measured times) */
int ta_attn_set_tc_lm(int variant); /* decoder-shape forward (head_dim 128, causal): 1 (default) = 64-key tiles, two CTAs per SM, K/V ring of 4 slots (3 / 4 / 6 / 8 select the depth; 6 and 8 leave room for one CTA per SM only); 0 = 128-key tiles, one CTA per SM */
int ta_attn_tc_lm_ring_slots(void);  /* the ring depth in use */
/* diagnostic timeline of the persistent encoder-attention kernel (tools/attn_trace.py): buf int64 [3][steps][8] device memory, NULL = off */
int ta_attn_set_trace(void* buf, int steps);
int ta_attn_set_bwd_variant(int variant); /* decoder attention backward: 2 (default) = 64-query sub-tiles, software-pipelined, dQ^T with its own TMEM buffers; 1 = 128-query serial kernel */
int ta_attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse,
                float* dsum_ws, float* dq_acc, void* dk, void* dv, int B, int S, int Hq, int Hkv, int head_dim,
                long long q_rs, long long k_rs, long long v_rs, long long o_rs, long long do_rs, long long dq_rs,
                long long dk_rs, long long dv_rs, int causal, float scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * HBM-bound building blocks (exported for unit parity tests; the engines below call them internally)
 * ---------------------------------------------------------------------------------------------- */
int ta_im2col_k3(const void* x /*bf16 [B,T,C]*/, void* out /*bf16 [B*T2, 3C]*/, int B, int T, int C, int stride, void* stream);
int ta_layernorm_set_reverse(int on); /* 1 (default): rows are processed last-to-first (the tail of the producer's output is still in L2) */
int ta_layernorm_bf16(const void* x, const float* w, const float* b, void* y, long long rows, int D, float eps, void* stream);
int ta_rmsnorm_f32(const float* x, const float* w, void* y_bf16, const int* row_index, long long rows, int D, float eps, void* stream);
int ta_rmsnorm_f32_bwd(const void* dy_bf16, const float* x, const float* w, float* dx, const int* row_index, long long rows,
                       int D, float eps, int accumulate, void* stream);
int ta_enc_rope(void* qkv, const float* cos_t, const float* sin_t, long long rows, int S, int H, int hd, int rot_dim, void* stream);
int ta_lm_qknorm_rope_fwd(const void* qkv, void* qk, const float* q_norm_w, const float* k_norm_w, const float* cos_t,
                          const float* sin_t, long long M, int S, int Hq, int Hkv, float eps, void* stream);
int ta_lm_qknorm_rope_bwd(const void* qkv, const float* dq, const void* dk, const void* dv, void* dqkv, const float* q_norm_w,
                          const float* k_norm_w, const float* cos_t, const float* sin_t, long long M, int S, int Hq, int Hkv,
                          float eps, void* stream);
int ta_proj_norm_fwd(const void* x_bf16, const float* w, void* y, long long rows, int D, float eps, int gelu, void* stream);
int ta_proj_norm_bwd(const void* x_bf16, const float* w, const void* dy, int dy_is_f32, void* dx_bf16, float* dw, long long rows,
                     int D, float eps, int gelu, void* stream);
/* tiny_audio/asr_modeling.py:27-44 (_gather_audio_embeds) + :511-515 (masked_scatter): exact index semantics */
int ta_audio_index(const long long* input_ids, const long long* token_counts, int* src_row, int B, int S, int n_a,
                   long long audio_token_id, void* stream);
int ta_embed_scatter(const long long* input_ids, const int* src_row, const float* embed_table, const float* audio_embeds,
                     float* inputs_embeds, long long n_tok, int D, long long vocab, void* stream);
int ta_audio_grad_gather(const int* src_row, const float* d_inputs_embeds, float* d_audio_embeds, long long n_tok, int D, void* stream);
/* HF:loss/loss_utils.py:28-67 on the labelled rows only; loss_sum += sum_rows(CE) * inv_items; logits <- d(logits).
 * With a row_loss buffer [rows] the sum is taken afterwards in a fixed order in double precision (reproducible, independent of the
 * batch sharding); with row_loss = NULL every row adds itself with a float atomic. */
int ta_ce_fwd_bwd(void* logits_bf16, long long ld, const int* targets, long long rows, int V, int Vpad, float inv_items,
                  float* loss_sum, float* row_loss, int write_grad, void* stream);
int ta_transpose_bf16(const void* in, void* out, int R, int C, long long ld_in, long long ld_out, void* stream);
int ta_cast_f32_bf16(const float* in, void* out, long long n, void* stream);
/* a6. QFormer projector window attention (tiny_audio/projectors.py:431-475 -> HF:models/blip_2/modeling_blip_2.py:579-634):
 *     per (window, head): out = dropout(softmax(scale * q k^T)) v with nq <= 4 queries, nk <= 16 keys, head_dim <= 96.
 *     q, out, dq bf16 [n_win, nq, heads*head_dim]; k, v, dk, dv bf16 [n_win, nk, heads*head_dim];
 *     drop_mask NULL or f32 [n_win, heads, nq, nk] holding 0 or 1/(1-p) (what F.dropout multiplies by). */
int ta_window_attn_fwd(const void* q, const void* k, const void* v, const float* drop_mask, void* out, long long n_win, int nq,
                       int nk, int heads, int head_dim, float scale, void* stream);
int ta_window_attn_bwd(const void* q, const void* k, const void* v, const float* drop_mask, const void* d_out, void* dq, void* dk,
                       void* dv, long long n_win, int nq, int nk, int heads, int head_dim, float scale, void* stream);
/* A/B switch: 1 = lanes over head_dim (any head_dim), 2 = key per lane with 16-byte accesses (head_dim % 8 == 0, else falls back
 * to 1); returns the previous setting */
int ta_window_attn_set_variant(int variant);
/* a6 (QFormer projector glue; tiny_audio/projectors.py:359-475 = HF:models/blip_2/modeling_blip_2.py Blip2QFormerSelfOutput :637-648,
 * Blip2QFormerOutput :693-704, Blip2QFormerIntermediate :677-689, query LayerNorm + dropout :985-986):
 *   y = LayerNorm(bf16(o) * mask + resid[row % resid_rows]) * post_mask[row % post_rows]   (o / mask / resid / post_mask optional; masks are
 *   the 0 or 1/(1-p) dropout multipliers), fp32 out + optional bf16 copy for the next GEMM; stats [rows, 2] = (mean, rstd) for the backward.
 *   Backward: dy = g32 + g16 (either optional) -> d_o bf16 (masked), d_resid fp32 (atomically reduced when resid is row-broadcast), dw, db. */
int ta_add_layernorm_fwd(const void* o, const float* mask, const float* resid, long long resid_rows, const float* w, const float* b,
                         const float* post_mask, long long post_rows, float* y32, void* y16, float* stats, long long rows, int H, float eps,
                         void* stream);
int ta_add_layernorm_bwd(const float* g32, const void* g16, const void* o, const float* mask, const float* resid, long long resid_rows,
                         const float* w, const float* post_mask, long long post_rows, const float* stats, void* d_o, float* d_resid, float* dw,
                         float* db, float* partial /* ta_add_layernorm_bwd_partial_floats(H) floats of scratch: per-block dw / db partial
                         sums, reduced in a fixed order (deterministic) */, long long rows, int H, void* stream);
long long ta_add_layernorm_bwd_partial_floats(int H);
/* exact-erf GELU on bf16 storage and its backward dx = dy (Phi(x) + x phi(x)); n elements, multiple of 8 */
int ta_gelu_fwd_bf16(const void* x, void* y, long long n, void* stream);
int ta_gelu_bwd_bf16(const void* x, const void* dy, void* dx, long long n, void* stream);
/* out[j] = sum_r x[r, j] (bias gradients of the tcgen05 linears); out is zeroed inside */
int ta_colsum_bf16(const void* x, long long ld, float* out, long long rows, int cols, void* stream);
/* tiny_audio/projectors.py:79-87 (_frame_stack): row j <- frames k*j .. k*j+k-1, feature-major per frame */
int ta_frame_stack(const void* x /*bf16 [B,S,D]*/, void* out /*bf16 [B,n,k*D]*/, int B, int S, int n, int k, int D, void* stream);
/* a4. tiny_audio/asr_modeling.py:458-479 (_maybe_drop_audio_tokens): whole encoder frames zeroed by a {0,1} keep mask, no rescale.
 * The Bernoulli draw stays with torch's generator (same RNG contract as the reference); this applies it in place. */
int ta_frame_keep_mask(void* x /*bf16 [rows,D], in place*/, const float* keep /*[rows]*/, long long rows, int D, void* stream);
/* a9. HF:loss/loss_utils.py:56-59 (shift labels by one, ignore_index -100) as device-side bookkeeping: ascending flat positions
 * p = b*S + s with labels[b, s+1] != -100, their targets, and the count -- replaces a labels.cpu() + nonzero() on the host when
 * the Trainer hands device-resident labels; the caller reads back the 4-byte count to size the lm_head product. */
/* f2. GPU-side collation (scripts/train.py:324-348 assembles these on the CPU dataloader workers): row b = prefix | <audio> x counts[b] |
 * middle | response_b | suffix | pad; labels = -100 except the response and the first suffix token; attention_mask marks the real tokens. */
int ta_assemble_prompts(const long long* counts /*[B]*/, const long long* resp /*packed*/, const long long* resp_off /*[B+1]*/,
                        const long long* prefix, int n_prefix, const long long* middle, int n_middle, const long long* suffix, int n_suffix,
                        long long audio_id, long long pad_id, int B, int S, long long* ids /*[B,S]*/, long long* labels, long long* mask,
                        void* stream);
int ta_label_rows(const long long* labels /*[B,S]*/, int B, int S, int* rows /*[B*S]*/, int* targets /*[B*S]*/, int* count /*[1]*/,
                  void* stream);

/* ------------------------------------------------------------------------------------------------
 * a12. optimiser: clip_grad_norm_(1.0) + torch.optim.AdamW(fused) (configs/training/production.yaml:5-9)
 *   per step: zero *gnorm_sq, ta_grad_sumsq() per tensor (after the DDP all-reduce), ta_adamw_clip_step() per tensor
 * ---------------------------------------------------------------------------------------------- */
int ta_grad_sumsq(const float* g, long long n, float* gnorm_sq, void* stream);
int ta_adamw_clip_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                       float weight_decay, int step, float max_grad_norm, const float* gnorm_sq, void* stream);

/* ------------------------------------------------------------------------------------------------
 * a3. GLM-ASR encoder forward (frozen, no grad): GlmAsrEncoder.forward, HF:models/glmasr/modeling_glmasr.py:316-330
 *   weights are bf16 for GEMM operands, fp32 for biases / LayerNorm; `layers` is a HOST array of
 *   n_layers * TA_ENC_PTRS_PER_LAYER device pointers in the order of the enum below.
 * ---------------------------------------------------------------------------------------------- */
enum { TA_ENC_LN1_W = 0, TA_ENC_LN1_B, TA_ENC_WQKV, TA_ENC_BQKV, TA_ENC_WO, TA_ENC_BO, TA_ENC_LN2_W, TA_ENC_LN2_B,
       TA_ENC_W1, TA_ENC_B1, TA_ENC_W2, TA_ENC_B2, TA_ENC_PTRS_PER_LAYER };
typedef struct ta_encoder_weights {
    int n_layers, dim, ffn, heads, head_dim, rot_dim, n_mels, max_pos;
    float ln_eps;
    const void* conv1_w;   /* bf16 [dim, 3*n_mels], K index = tap*n_mels + c */
    const float* conv1_b;
    const void* conv2_w;   /* bf16 [dim, 3*dim],    K index = tap*dim + c    */
    const float* conv2_b;
    const float* lnf_w;
    const float* lnf_b;
    const float* rope_cos; /* [max_pos, rot_dim/2] (values already rounded to bf16, as the reference's cos.to(x.dtype)) */
    const float* rope_sin;
    const void* const* layers;
} ta_encoder_weights;
int ta_encoder_workspace_bytes(const ta_encoder_weights* w, int B, int T, long long* bytes);
int ta_encoder_forward(const ta_encoder_weights* w, const void* conv1_im2col_bf16 /*[B*T, 3*n_mels]*/, int B, int T,
                       void* workspace, long long workspace_bytes, void* out_bf16 /*[B*S_e, dim]*/, void* stream);

/* ------------------------------------------------------------------------------------------------
 * a5. MLP projector forward / backward: MLPAudioProjector.forward, tiny_audio/projectors.py:57-71 (+ autograd)
 *   x_stacked: bf16 [M, k*enc_dim] (frame-stacked encoder output).  Stash tensors are caller-owned.
 * ---------------------------------------------------------------------------------------------- */
typedef struct ta_mlp_projector_weights {
    int in_dim, hidden, out_dim;
    float eps;
    const void* w1;        /* bf16 [hidden, in_dim]  */
    const float* norm_w;   /* fp32 [hidden]          */
    const void* w2;        /* bf16 [out_dim, hidden] */
    const void* w2_t;      /* bf16 [hidden, out_dim] (backward only) */
    const float* norm2_w;  /* fp32 [out_dim]         */
} ta_mlp_projector_weights;
int ta_mlp_projector_forward(const ta_mlp_projector_weights* w, const void* x_stacked, long long M, void* y1 /*bf16 [M,hidden]*/,
                             void* a1 /*bf16 [M,hidden]*/, void* y2 /*bf16 [M,out]*/, float* out /*fp32 [M,out]*/, void* stream);
int ta_mlp_projector_backward_workspace_bytes(const ta_mlp_projector_weights* w, long long M, long long* bytes);
int ta_mlp_projector_backward(const ta_mlp_projector_weights* w, const void* x_stacked, long long M, const void* y1, const void* a1,
                              const void* y2, const float* d_out /*fp32 [M,out]*/, void* workspace, long long workspace_bytes,
                              float* d_w1 /*fp32 [hidden,in]*/, float* d_norm_w /*accumulated*/, float* d_w2, float* d_norm2_w,
                              void* stream);

/* ------------------------------------------------------------------------------------------------
 * a8-a10. Qwen3 forward + CE + backward-to-inputs (frozen LM, dgrad only):
 *   Qwen3ForCausalLM.forward(inputs_embeds, labels) HF:models/qwen3/modeling_qwen3.py:458-517, ForCausalLMLoss
 *   HF:loss/loss_utils.py:45-67, autograd backward HF:trainer.py:1935.  lm_head is evaluated on the labelled rows.
 * ---------------------------------------------------------------------------------------------- */
enum { TA_LM_LN1_W = 0, TA_LM_WQKV, TA_LM_WQKV_T, TA_LM_QNORM_W, TA_LM_KNORM_W, TA_LM_WO, TA_LM_WO_T, TA_LM_LN2_W,
       TA_LM_WGU, TA_LM_WGU_T, TA_LM_WD, TA_LM_WD_T,
       /* LoRA (a11; tiny_audio/asr_modeling.py:289-301, peft r=8 alpha=32 on q,k,v,o,gate,up,down), NULL when lora_pad == 0:
          A_*  bf16 [lora_pad, K_in]   stacked lora_A of the fused projection's adapters (q|k|v -> rows 0-7, 8-15, 16-23)
          BT_* bf16 [lora_pad, N_out]  (alpha/r * lora_B)^T with the same row blocks                                     */
       TA_LM_LORA_A_QKV, TA_LM_LORA_A_O, TA_LM_LORA_A_GU, TA_LM_LORA_A_D,
       TA_LM_LORA_BT_QKV, TA_LM_LORA_BT_O, TA_LM_LORA_BT_GU, TA_LM_LORA_BT_D, TA_LM_PTRS_PER_LAYER };
/* per-layer LoRA gradient outputs (fp32, caller-owned): dA [lora_pad, K_in] and dB' [N_out, lora_pad] = d loss / d (alpha/r * B) */
enum { TA_LM_LORA_DA_QKV = 0, TA_LM_LORA_DB_QKV, TA_LM_LORA_DA_O, TA_LM_LORA_DB_O, TA_LM_LORA_DA_GU, TA_LM_LORA_DB_GU,
       TA_LM_LORA_DA_D, TA_LM_LORA_DB_D, TA_LM_LORA_GRADS_PER_LAYER };
/* unfrozen LM (f3; configs/experiments/embedded.yaml:19-33 `freeze_language_model: false`): per-layer weight-gradient outputs, fp32,
   caller-owned.  W* are overwritten, in the layout of the packed operand (WQKV rows = [q | k | v], WGU rows interleaved in 64-blocks);
   the norm gradients are ACCUMULATED (atomics): the caller zeroes them before the step. */
enum { TA_LM_G_WQKV = 0, TA_LM_G_WO, TA_LM_G_WGU, TA_LM_G_WD, TA_LM_G_LN1, TA_LM_G_LN2, TA_LM_G_QNORM, TA_LM_G_KNORM, TA_LM_GRADS_PER_LAYER };
typedef struct ta_lm_weights {
    int n_layers, dim, ffn, n_q_heads, n_kv_heads, head_dim, max_pos;
    long long vocab, vocab_pad;
    float eps;
    int lora_pad;               /* 0, or 128: every W* / W*_T below carries lora_pad extra K columns ([W | aB] resp. [W^T | A^T]) */
    const float* embed_f32;     /* [vocab, dim]  (embedding lookup stays fp32 under autocast)            */
    const void* embed_bf16;     /* [vocab_pad, dim] bf16, rows >= vocab are zero (tied lm_head)          */
    const void* embed_bf16_t;   /* [dim, vocab_pad] bf16 (dgrad operand)                                  */
    const float* final_norm_w;
    const float* rope_cos;      /* [max_pos, head_dim/2] fp32 */
    const float* rope_sin;
    const void* const* layers;  /* HOST array, n_layers * TA_LM_PTRS_PER_LAYER device pointers:
                                   WQKV  bf16 [(Hq+2Hkv)*hd, dim]   WQKV_T bf16 [dim, (Hq+2Hkv)*hd]
                                   WO    bf16 [dim, Hq*hd]          WO_T   bf16 [Hq*hd, dim]
                                   WGU   bf16 [2*ffn, dim] rows interleaved in 64-blocks (gate, up)   WGU_T bf16 [dim, 2*ffn]
                                   WD    bf16 [dim, ffn]            WD_T   bf16 [ffn, dim] */
} ta_lm_weights;
typedef struct ta_lm_step_args {
    int B, S, n_labelled, with_backward;
    const float* inputs_embeds;   /* [B*S, dim] fp32 */
    const int* label_rows;        /* [n_labelled] flat token index whose hidden state predicts label_targets[i] */
    const int* label_targets;     /* [n_labelled] */
    float inv_num_items;          /* 1 / num_items_in_batch */
    float* loss;                  /* device scalar, accumulated into (caller zeroes) */
    float* row_loss;              /* optional [n_labelled] */
    float* d_inputs_embeds;       /* [B*S, dim] fp32 out (with_backward) */
    void* workspace;
    long long workspace_bytes;
    float* final_hidden;          /* optional [B*S, dim] fp32: last layer's output BEFORE the final norm (for ta_lm_hidden_to_logits) */
    float* const* lora_grads;     /* LoRA + backward: HOST array n_layers * TA_LM_LORA_GRADS_PER_LAYER device pointers */
    void* k_cache;                /* optional (prefill of generate): bf16 [n_layers, B, cache_max_seq, Hkv*hd]; rows [0, S) of every */
    void* v_cache;                /*   layer receive the prompt's roped keys / values (use_cache=True in HF generate)             */
    int cache_max_seq;
    float* const* lm_grads;       /* unfrozen LM + backward: HOST array n_layers * TA_LM_GRADS_PER_LAYER device pointers (else NULL) */
    float* d_embed;               /*   fp32 [vocab_pad, dim]: tied embed_tokens / lm_head gradient (overwritten)                    */
    float* d_final_norm;          /*   fp32 [dim], accumulated                                                                       */
    const long long* input_ids;   /*   [B*S]: which rows of the table the text positions read                                       */
    long long audio_token_id;
    /* forward-only (generate() with ragged prompts, tiny_audio/asr_modeling.py:587-640 -> HF generate with attention_mask): LEFT-padded
       prompts.  position_ids [B*S]: rotary position of every token (HF: cumsum(attention_mask) - 1); kv_start [B]: index of the first
       real token of each sequence -- keys before it are never attended to by real tokens.  NULL: arange(S) / no padding. */
    const int* position_ids;
    const int* kv_start;
    int lora_grads_zeroed;        /* != 0: the caller cleared every buffer in lora_grads before the call (one memset per stacked tensor
                                     instead of one in front of each of the 8 split-K gradient products per layer) */
} ta_lm_step_args;
/* with_backward: 0 forward only; 1 backward to inputs_embeds (frozen LM, LoRA); 2 additionally the weight gradients (lm_grads) */
int ta_lm_set_fused_attn_dsum(int on); /* 1 (default): the attention backward's D = rowsum(dO o O) comes out of the o-projection dgrad GEMM's epilogue (TA_EPI_BF16_ROWDOT); 0: separate preparation kernel */
int ta_lm_workspace_bytes(const ta_lm_weights* w, int B, int S, int n_labelled, int with_backward, long long* bytes);
/* building blocks of the unfrozen recipe (exported for unit parity tests and the host-side operand refresh) */
int ta_rmsnorm_dw(const void* dy_bf16, const float* x, const int* row_index, long long rows, int D, float eps, float* dw /*accumulated*/,
                  void* stream);
int ta_qknorm_dw(const void* qkv, const float* dq, const void* dk, const float* cos_t, const float* sin_t, long long M, int S, int Hq,
                 int Hkv, float eps, float* d_q_norm_w /*[128] accumulated*/, float* d_k_norm_w, void* stream);
int ta_embed_grad_scatter(const long long* input_ids, const float* d_inputs_embeds, float* d_table /*accumulated*/, long long n_tok,
                          int D, long long vocab, long long audio_token_id, void* stream);
/* fp32 master [R, C] -> bf16 operand rows (r / blk) * blk_stride + r % blk + row_off of dst (ld) and of the transposed copy dstT
   [C, ldT] (optional): refreshes the packed Qwen3 operands after an optimiser step */
int ta_pack_weight(const float* src, int R, int C, void* dst, long long ld, void* dstT, long long ldT, int blk, int blk_stride,
                   int row_off, void* stream);
int ta_lm_forward_backward(const ta_lm_weights* w, const ta_lm_step_args* a, void* stream);
/* full-vocabulary logits for given rows (eval / generate):  logits bf16 [n_rows, vocab_pad] */
int ta_lm_hidden_to_logits(const ta_lm_weights* w, const float* hidden_f32 /*[B*S, dim] pre-final-norm*/, const int* rows,
                           int n_rows, void* normed_ws /*bf16 [n_rows, dim]*/, void* logits_bf16, void* stream);

/* ------------------------------------------------------------------------------------------------
 * f1. greedy decode with a KV cache: ASRModel.generate -> language_model.generate(inputs_embeds, use_cache),
 *     tiny_audio/asr_modeling.py:562-646; greedy defaults asr_config.py:103-111.  Prefill = ta_lm_forward_backward with
 *     k_cache / v_cache set; then one ta_lm_decode_step per new token:
 *       ids [B] int64 (the token fed at position *pos) -> embed -> 28 x [RMSNorm, qkv, head-norm + RoPE + cache append,
 *       single-query attention over cache rows [kv_start[b], *pos], o + residual, RMSNorm, SwiGLU MLP + residual] -> final norm ->
 *       tied lm_head -> logits bf16 [B, vocab_pad] -> next_ids [B] = argmax;  *pos += 1 (device side, graph-replayable).
 *     B <= 32 per call.  All linears are HBM-bound skinny products (csrc/decode.cu).
 * ---------------------------------------------------------------------------------------------- */
enum { TA_SKINNY_BF16 = 0,      /* out bf16 [M,N] = X W^T                                                   */
       TA_SKINNY_F32_RESID = 1, /* out f32  [M,N] = resid + bf16(X W^T)                                     */
       TA_SKINNY_SWIGLU = 2,    /* W rows interleaved [64 gate | 64 up]: out bf16 [M,N/2] = bf16(silu(g)) * u */
       TA_SKINNY_PARTIAL = 3 }; /* out f32 [k_splits, M, ldo]: raw partial sums of a split-K product (small N)        */
int ta_skinny_gemm_bf16(const void* X, long long ldx, const void* W, long long ldw, int M /*<= 32*/, int N, int K, int mode,
                        void* out, long long ldo, const float* resid, int k_splits /*1 unless TA_SKINNY_PARTIAL*/, void* stream);
/* x_out = x_in + bf16(sum_s partial[s]) (partial may be NULL: x_out unused); y bf16 = RMSNorm(x_out) * w   (Qwen3RMSNorm) */
int ta_decode_resid_rmsnorm(const float* x_in, const float* partial, int n_splits, int rows, float* x_out, const float* w,
                            void* y_bf16, int D, float eps, long long ldy, void* stream);
int ta_decode_attn(const void* q /*bf16 [B, Hq*128]*/, const void* k_cache /*bf16 [B, max_seq, Hkv*128]*/, const void* v_cache,
                   void* out /*bf16 [B, ld_out]*/, long long ld_out, const int* pos /*device: attends rows [0, *pos]*/, int B, int Hq,
                   int Hkv, int max_seq, float scale, void* stream);
int ta_argmax_rows(const void* logits_bf16, long long ld, int rows, int V, long long* next_ids, void* stream);
int ta_lm_decode_workspace_bytes(const ta_lm_weights* w, int B, long long* bytes);
int ta_lm_decode_step(const ta_lm_weights* w, const long long* ids, int* pos /*device*/, int pos_host /*validation only*/,
                      void* k_cache, void* v_cache, int cache_max_seq, int B, void* workspace, long long workspace_bytes,
                      void* logits_bf16 /*[B, vocab_pad]*/, long long* next_ids /*[B]*/,
                      const int* kv_start /*[B] or NULL: first real cache row of each (left-padded) sequence; rotary position = *pos - kv_start[b]*/,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif
