// Host-side orchestration of the three towers of the path: all launches go to the caller's stream, no host
// synchronisation, no allocation (workspaces come from the caller).  This is the native "runtime" of the hot path:
// one C call per tower instead of ~2000 Python->ctypes launches per step.
#include "common.cuh"
#include "kernels.cuh"
#include "tinyaudio_b200.h"

extern int g_wgrad_transposed;   // attn_tc_bwd.cu (ta_debug_set key 2)
int g_fuse_attn_dsum = 1;         // ta_lm_set_fused_attn_dsum: D = rowsum(dO o O) in the o-projection dgrad epilogue (1, default) or its own kernel (0)

namespace {

struct Carver {
    uint8_t* base;
    long long off = 0, cap;
    Carver(void* p, long long c) : base(reinterpret_cast<uint8_t*>(p)), cap(c) {}
    template <typename T>
    T* take(long long n_elems) {
        off = (off + 255) & ~255LL;
        T* p = reinterpret_cast<T*>(base ? base + off : nullptr);
        off += n_elems * (long long)sizeof(T);
        return p;
    }
    bool ok() const { return off <= cap; }
};

inline int gemm(const void* A, long long lda, const void* B, long long ldb, long long M, int N, int K, int epi, void* out,
                long long ldo, const float* bias, const void* resid, void* out2, long long ldo2, const void* aux,
                long long ldaux, cudaStream_t st, float alpha = 1.0f, const float* rope_cos = nullptr,
                const float* rope_sin = nullptr, int rope_seq = 0, int rope_cols = 0) {
    if (M == 0) return 0;
    ta_gemm_epilogue e;
    memset(&e, 0, sizeof(e));
    e.rope_cos = rope_cos; e.rope_sin = rope_sin; e.rope_seq = rope_seq; e.rope_cols = rope_cols;
    e.out = out; e.ldo = ldo; e.bias = bias; e.resid = resid; e.ldr = ldo; e.out2 = out2; e.ldo2 = ldo2; e.aux = aux;
    e.ldaux = ldaux; e.alpha = alpha;
    return ta_gemm_bf16(A, lda, B, ldb, (int)M, N, K, epi, &e, st);
}
#define RUN(x)            \
    do {                  \
        int _rc = (x);    \
        if (_rc) return _rc; \
    } while (0)

inline int enc_out_len(int T) { return (T + 2 - 3) / 2 + 1; }

}  // namespace

// =============================================================================================================
// encoder
// =============================================================================================================
static long long enc_carve(const ta_encoder_weights* w, int B, int T, void* ws, long long cap, bf16** x, bf16** h, bf16** qkv,
                           bf16** f) {
    const long long S = enc_out_len(T), M = (long long)B * S, M1 = (long long)B * T;
    const long long D = w->dim, F = w->ffn;
    Carver c(ws, cap);
    *x = c.take<bf16>(M * D);
    *h = c.take<bf16>(M * D);
    // qkv doubles as the conv2 im2col buffer [M, 3D]; f doubles as the conv1 output [M1, D]
    *qkv = c.take<bf16>(M * 3 * D);
    const long long f_elems = (M * F > M1 * D) ? M * F : M1 * D;
    *f = c.take<bf16>(f_elems);
    return c.off;
}

TA_API int ta_encoder_workspace_bytes(const ta_encoder_weights* w, int B, int T, long long* bytes) {
    TA_REQUIRE(w && bytes, "null");
    bf16 *a, *b, *c, *d;
    *bytes = enc_carve(w, B, T, nullptr, 0, &a, &b, &c, &d) + 256;
    return 0;
}

TA_API int ta_encoder_forward(const ta_encoder_weights* w, const void* conv1_im2col, int B, int T, void* workspace,
                              long long workspace_bytes, void* out, void* stream) {
    TA_REQUIRE(w && conv1_im2col && workspace && out, "ta_encoder_forward: null pointer");
    TA_REQUIRE(w->head_dim == 64 && w->heads * w->head_dim == w->dim && w->rot_dim == 32,
               "encoder: head_dim must be 64 (rotary 32) and heads*64 == dim");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int S = enc_out_len(T);
    TA_REQUIRE(S <= w->max_pos, "encoder: %d frames exceed the rotary table (%d)", S, w->max_pos);
    const long long M = (long long)B * S, M1 = (long long)B * T;
    const int D = w->dim, F = w->ffn;
    bf16 *x, *h, *qkv, *f;
    const long long need = enc_carve(w, B, T, workspace, workspace_bytes, &x, &h, &qkv, &f);
    TA_REQUIRE(need <= workspace_bytes, "encoder workspace too small: need %lld, have %lld", need, workspace_bytes);
    if (M == 0) return 0;

    // conv1 (k3, s1, p1) + GELU as a GEMM over the im2col rows; conv2 (k3, s2, p1) + GELU likewise
    bf16* c1 = f;
    RUN(gemm(conv1_im2col, 3 * w->n_mels, w->conv1_w, 3 * w->n_mels, M1, D, 3 * w->n_mels, TA_EPI_BF16_GELU, c1, D, w->conv1_b,
             nullptr, nullptr, 0, nullptr, 0, st));
    bf16* im2 = qkv;
    RUN(k_im2col_k3(c1, im2, B, T, D, 2, st));
    RUN(gemm(im2, 3 * D, w->conv2_w, 3 * D, M, D, 3 * D, TA_EPI_BF16_GELU, x, D, w->conv2_b, nullptr, nullptr, 0, nullptr, 0, st));

    const float scale = 1.0f / sqrtf((float)w->head_dim);
    for (int l = 0; l < w->n_layers; ++l) {
        const void* const* L = w->layers + (long long)l * TA_ENC_PTRS_PER_LAYER;
        RUN(k_layernorm_bf16(x, (const float*)L[TA_ENC_LN1_W], (const float*)L[TA_ENC_LN1_B], h, M, D, w->ln_eps, st));
        // q|k|v projection with bias and the partial rotary embedding fused into the epilogue
        RUN(gemm(h, D, L[TA_ENC_WQKV], D, M, 3 * D, D, TA_EPI_BF16_ROPE, qkv, 3 * D, (const float*)L[TA_ENC_BQKV], nullptr, nullptr, 0,
                 nullptr, 0, st, 1.0f, w->rope_cos, w->rope_sin, S, 2 * D));
        RUN(ta_attn_fwd(qkv, qkv + D, qkv + 2 * D, h, nullptr, B, S, w->heads, w->heads, w->head_dim, 3 * D, 3 * D, 3 * D, D, 0,
                        scale, st));
        RUN(gemm(h, D, L[TA_ENC_WO], D, M, D, D, TA_EPI_BF16_RESID, x, D, (const float*)L[TA_ENC_BO], x, nullptr, 0, nullptr, 0, st));
        RUN(k_layernorm_bf16(x, (const float*)L[TA_ENC_LN2_W], (const float*)L[TA_ENC_LN2_B], h, M, D, w->ln_eps, st));
        RUN(gemm(h, D, L[TA_ENC_W1], D, M, F, D, TA_EPI_BF16_GELU, f, F, (const float*)L[TA_ENC_B1], nullptr, nullptr, 0, nullptr, 0, st));
        RUN(gemm(f, F, L[TA_ENC_W2], F, M, D, F, TA_EPI_BF16_RESID, x, D, (const float*)L[TA_ENC_B2], x, nullptr, 0, nullptr, 0, st));
    }
    RUN(k_layernorm_bf16(x, w->lnf_w, w->lnf_b, reinterpret_cast<bf16*>(out), M, D, w->ln_eps, st));
    return 0;
}

// =============================================================================================================
// MLP projector
// =============================================================================================================
TA_API int ta_mlp_projector_forward(const ta_mlp_projector_weights* w, const void* x_stacked, long long M, void* y1, void* a1,
                                    void* y2, float* out, void* stream) {
    TA_REQUIRE(w && x_stacked && y1 && a1 && y2 && out, "ta_mlp_projector_forward: null pointer");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (M == 0) return 0;
    RUN(gemm(x_stacked, w->in_dim, w->w1, w->in_dim, M, w->hidden, w->in_dim, TA_EPI_BF16, y1, w->hidden, nullptr, nullptr, nullptr,
             0, nullptr, 0, st));
    RUN(k_proj_norm_fwd((const bf16*)y1, w->norm_w, a1, M, w->hidden, w->eps, 1, st));
    RUN(gemm(a1, w->hidden, w->w2, w->hidden, M, w->out_dim, w->hidden, TA_EPI_BF16, y2, w->out_dim, nullptr, nullptr, nullptr, 0,
             nullptr, 0, st));
    RUN(k_proj_norm_fwd((const bf16*)y2, w->norm2_w, out, M, w->out_dim, w->eps, 0, st));
    return 0;
}

static long long proj_bwd_carve(const ta_mlp_projector_weights* w, long long M, void* ws, long long cap, bf16** dy2, bf16** da1,
                                bf16** tA, bf16** tB, long long* Mp) {
    *Mp = (M + 7) / 8 * 8;
    Carver c(ws, cap);
    *dy2 = c.take<bf16>(M * w->out_dim);
    *da1 = c.take<bf16>(M * w->hidden);
    const long long amax = (w->out_dim > w->hidden) ? w->out_dim : w->hidden;
    const long long bmax = (w->hidden > w->in_dim) ? w->hidden : w->in_dim;
    *tA = c.take<bf16>(amax * *Mp);
    *tB = c.take<bf16>(bmax * *Mp);
    return c.off;
}

TA_API int ta_mlp_projector_backward_workspace_bytes(const ta_mlp_projector_weights* w, long long M, long long* bytes) {
    TA_REQUIRE(w && bytes, "null");
    bf16 *a, *b, *c, *d;
    long long mp;
    *bytes = proj_bwd_carve(w, M, nullptr, 0, &a, &b, &c, &d, &mp) + 256;
    return 0;
}

TA_API int ta_mlp_projector_backward(const ta_mlp_projector_weights* w, const void* x_stacked, long long M, const void* y1,
                                     const void* a1, const void* y2, const float* d_out, void* workspace, long long workspace_bytes,
                                     float* d_w1, float* d_norm_w, float* d_w2, float* d_norm2_w, void* stream) {
    TA_REQUIRE(w && x_stacked && y1 && a1 && y2 && d_out && workspace && d_w1 && d_norm_w && d_w2 && d_norm2_w && w->w2_t,
               "ta_mlp_projector_backward: null pointer");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (M == 0) return 0;
    TA_REQUIRE(M < (1LL << 31), "projector backward: too many rows");
    bf16 *dy2, *da1, *tA, *tB;
    long long Mp;
    const long long need = proj_bwd_carve(w, M, workspace, workspace_bytes, &dy2, &da1, &tA, &tB, &Mp);
    TA_REQUIRE(need <= workspace_bytes, "projector backward workspace too small: need %lld, have %lld", need, workspace_bytes);
    const int H = w->hidden, O = w->out_dim, I = w->in_dim;
    // norm_2 backward -> d(y2) (bf16), d(norm_2.weight)
    RUN(k_proj_norm_bwd((const bf16*)y2, w->norm2_w, d_out, 1, dy2, d_norm2_w, M, O, w->eps, 0, st));
    // wgrad linear_2: dW2[O,H] = dy2^T . a1      (contraction over the M rows; TN GEMM: operands as they lie in memory)
    const bool tn = !g_wgrad_transposed;
    if (tn) {
        RUN(ta_gemm_bf16_tn(dy2, O, a1, H, O, H, (int)M, d_w2, H, 1.0f, st));
    } else {
        RUN(k_transpose_bf16(dy2, tA, (int)M, O, O, Mp, st));
        RUN(k_transpose_bf16((const bf16*)a1, tB, (int)M, H, H, Mp, st));
        RUN(gemm(tA, Mp, tB, Mp, O, H, (int)M, TA_EPI_F32, d_w2, H, nullptr, nullptr, nullptr, 0, nullptr, 0, st));
    }
    // dgrad linear_2: d(a1)[M,H] = dy2 . W2
    RUN(gemm(dy2, O, w->w2_t, O, M, H, O, TA_EPI_BF16, da1, H, nullptr, nullptr, nullptr, 0, nullptr, 0, st));
    // GELU + norm backward (in place: d(a1) -> d(y1)), d(norm.weight)
    RUN(k_proj_norm_bwd((const bf16*)y1, w->norm_w, da1, 0, da1, d_norm_w, M, H, w->eps, 1, st));
    // wgrad linear_1: dW1[H,I] = dy1^T . x_stacked
    if (tn) {
        RUN(ta_gemm_bf16_tn(da1, H, x_stacked, I, H, I, (int)M, d_w1, I, 1.0f, st));
    } else {
        RUN(k_transpose_bf16(da1, tA, (int)M, H, H, Mp, st));
        RUN(k_transpose_bf16((const bf16*)x_stacked, tB, (int)M, I, I, Mp, st));
        RUN(gemm(tA, Mp, tB, Mp, H, I, (int)M, TA_EPI_F32, d_w1, I, nullptr, nullptr, nullptr, 0, nullptr, 0, st));
    }
    return 0;
}

// =============================================================================================================
// Qwen3: forward + CE + backward to inputs_embeds  (+ optional LoRA adapters on all seven projections)
//
// LoRA (reference: tiny_audio/asr_modeling.py:289-301 -> peft LoraConfig(r, alpha, dropout 0) on q,k,v,o,gate,up,down):
//   y = x W^T + s (x A^T) B^T.  Every projection keeps running as ONE tcgen05 GEMM by augmenting the contraction:
//   the activation buffers carry P = lora_pad extra columns that hold t = x A^T (rank padded to P), and the packed
//   weights carry P extra columns that hold s*B, so  [x | t] . [W | sB]^T  is the adapted projection; the dgrad GEMMs
//   use [dy | u] . [W^T | A^T]^T with u = dy (sB).  Only the rank-P "down" products (t, u) and the A/B weight gradients
//   are extra launches.
// =============================================================================================================
namespace {
struct LmBufs {
    // per-layer stash (index by layer when with_backward, else all layers alias slot 0)
    float* resid;      // [L+1][M, D] : resid[l] = input of layer l (l >= 1), resid[L] = final hidden
    float* resid_mid;  // [L][M, D]
    bf16* qkv;         // [L][M, QKV]
    bf16* qk;          // [L][M, QK]
    bf16* att;         // [L][M, QD + P]   (extra columns: t_o)
    float* lse;        // [L][B*Hq*S]
    bf16* gu;          // [L][M, 2F]
    bf16* lt;          // [L][3][M, P]     LoRA only: t_qkv, t_gu, t_d
    // scratch
    bf16* xn;          // [M, D + P]
    bf16* h;           // [M, F + P]
    bf16* hl;          // [n_lab, D]
    bf16* logits;      // [n_lab, Vpad]
    bf16* dhl;         // [n_lab, D]
    float* row_loss;   // [n_lab] per-row CE losses (reduced in a fixed order)
    int h_layers;      // 1: h is kept per layer (LoRA / unfrozen decoder with backward)
    int lora_zeroed;   // the caller cleared every LoRA gradient buffer before the step (ta_lm_step_args.lora_grads_zeroed)
    bf16* dxb;         // [M, D + P]
    bf16* big;         // [M, max(2F, QKV) + P]
    bf16* dxn;         // [M, D]
    bf16* datt;        // [M, QD]
    float* dq;         // [M, QD]
    bf16* dk;          // [M, KD]
    bf16* dv;          // [M, KD]
    float* dsum;       // [B*Hq*S]
    bf16* tr_a;        // LoRA wgrad transposes: [max(2F, QKV), Mp], [P, Mp], [P, Mp], [max(F, QD), Mp]
    bf16* tr_t;
    bf16* tr_u;
    bf16* tr_x;
    long long stride_layers;   // 1 if with_backward else 0
};

long long lm_carve(const ta_lm_weights* w, int B, int S, int n_lab, int with_bwd, void* ws, long long cap, LmBufs* b,
                   int train_lm = 0) {
    const long long M = (long long)B * S, D = w->dim, F = w->ffn, P = w->lora_pad;
    const long long QD = (long long)w->n_q_heads * w->head_dim, KD = (long long)w->n_kv_heads * w->head_dim;
    const long long QKV = QD + 2 * KD, QK = QD + KD;
    const long long L = with_bwd ? w->n_layers : 1;
    const long long Mp = (M + 7) / 8 * 8;
    Carver c(ws, cap);
    b->stride_layers = with_bwd ? 1 : 0;
    b->resid = c.take<float>((with_bwd ? (w->n_layers + 1) : 2) * M * D);
    b->resid_mid = c.take<float>(L * M * D);
    b->qkv = c.take<bf16>(L * M * QKV);
    b->qk = c.take<bf16>(L * M * QK);
    b->att = c.take<bf16>(L * M * (QD + P));
    b->lse = c.take<float>(L * (long long)B * w->n_q_heads * S);
    b->gu = c.take<bf16>(L * M * 2 * F);
    b->lt = c.take<bf16>(P ? L * 3 * M * P : 0);
    b->xn = c.take<bf16>(M * (D + P));
    // h = silu(gate) * up is the input of down_proj: recipes that need its weight / adapter gradients keep it per layer (2.7 GB at the
    // headline shape) instead of recomputing it from the (gate, up) stash in the backward (1.5 ms per step)
    b->h_layers = (with_bwd && (P || train_lm)) ? 1 : 0;
    b->h = c.take<bf16>((b->h_layers ? (long long)w->n_layers : 1) * M * (F + P));
    b->hl = c.take<bf16>((long long)n_lab * D);
    b->logits = c.take<bf16>((long long)n_lab * w->vocab_pad);
    b->dhl = c.take<bf16>((long long)n_lab * D);
    b->row_loss = c.take<float>((long long)n_lab);
    if (with_bwd) {
        b->dxb = c.take<bf16>(M * (D + P));
        const long long bigw = (2 * F > QKV) ? 2 * F : QKV;
        b->big = c.take<bf16>(M * (bigw + P));
        b->dxn = c.take<bf16>(M * D);
        b->datt = c.take<bf16>(M * QD);
        b->dq = c.take<float>(M * QD);
        b->dk = c.take<bf16>(M * KD);
        b->dv = c.take<bf16>(M * KD);
        b->dsum = c.take<float>((long long)B * w->n_q_heads * S);
        if (P) {
            b->tr_a = c.take<bf16>(bigw * Mp);
            b->tr_t = c.take<bf16>(P * Mp);
            b->tr_u = c.take<bf16>(P * Mp);
            b->tr_x = c.take<bf16>(((F > QD) ? F : QD) * Mp);
        } else if (train_lm) {   // operand transposes of the weight-gradient GEMMs (dW = dY^T X, contraction over the tokens)
            const long long nlp = ((long long)n_lab + 7) / 8 * 8;
            const long long a1 = bigw * Mp, a2 = w->vocab_pad * nlp;
            b->tr_a = c.take<bf16>(a1 > a2 ? a1 : a2);
            const long long x1 = ((F > QD) ? F : QD) * Mp, x2 = D * nlp;
            b->tr_x = c.take<bf16>(x1 > x2 ? x1 : x2);
        }
    }
    return c.off;
}

inline int plain(const void* A, long long lda, const void* Bm, long long ldb, long long M, int N, int K, void* out, long long ldo,
                 cudaStream_t st) {
    return gemm(A, lda, Bm, ldb, M, N, K, TA_EPI_BF16, out, ldo, nullptr, nullptr, nullptr, 0, nullptr, 0, st);
}

// A / B gradients of one adapter group:  dBs[N_out, P] = dy^T t,  dA[P, K_in] = u^T x   (contractions over the M tokens)
int lora_wgrad(const LmBufs& b, long long M, int P, const bf16* dy, long long ld_dy, int n_out, const bf16* t, long long ld_t,
               const bf16* u, long long ld_u, const bf16* x, long long ld_x, int k_in, float* dA, float* dBs, cudaStream_t st) {
    if (!g_wgrad_transposed) {
        RUN(k_gemm_bf16_tn(dy, ld_dy, t, ld_t, n_out, P, (int)M, dBs, P, 1.0f, st, b.lora_zeroed));
        RUN(k_gemm_bf16_tn(u, ld_u, x, ld_x, P, k_in, (int)M, dA, k_in, 1.0f, st, b.lora_zeroed));
        return 0;
    }
    const long long Mp = (M + 7) / 8 * 8;
    RUN(k_transpose_bf16(dy, b.tr_a, (int)M, n_out, ld_dy, Mp, st));
    RUN(k_transpose_bf16(t, b.tr_t, (int)M, P, ld_t, Mp, st));
    RUN(k_transpose_bf16(u, b.tr_u, (int)M, P, ld_u, Mp, st));
    RUN(k_transpose_bf16(x, b.tr_x, (int)M, k_in, ld_x, Mp, st));
    RUN(gemm(b.tr_a, Mp, b.tr_t, Mp, n_out, P, (int)M, TA_EPI_F32, dBs, P, nullptr, nullptr, nullptr, 0, nullptr, 0, st));
    RUN(gemm(b.tr_u, Mp, b.tr_x, Mp, P, k_in, (int)M, TA_EPI_F32, dA, k_in, nullptr, nullptr, nullptr, 0, nullptr, 0, st));
    return 0;
}
// weight gradient of one linear:  dW[n_out, k_in] = dy^T x   (contraction over the M tokens; fp32 out, overwritten)
int full_wgrad(const LmBufs& b, long long M, const bf16* dy, long long ld_dy, int n_out, const bf16* x, long long ld_x, int k_in,
               float* dW, cudaStream_t st) {
    if (!g_wgrad_transposed)   // both operands MN-major straight from the activations (gemm_sm100.cu, TN variant)
        return ta_gemm_bf16_tn(dy, ld_dy, x, ld_x, n_out, k_in, (int)M, dW, k_in, 1.0f, st);
    const long long Mp = (M + 7) / 8 * 8;
    RUN(k_transpose_bf16(dy, b.tr_a, (int)M, n_out, ld_dy, Mp, st));
    RUN(k_transpose_bf16(x, b.tr_x, (int)M, k_in, ld_x, Mp, st));
    RUN(gemm(b.tr_a, Mp, b.tr_x, Mp, n_out, k_in, (int)M, TA_EPI_F32, dW, k_in, nullptr, nullptr, nullptr, 0, nullptr, 0, st));
    return 0;
}
}  // namespace

TA_API int ta_lm_set_fused_attn_dsum(int on) {
    g_fuse_attn_dsum = on ? 1 : 0;
    return 0;
}

TA_API int ta_lm_workspace_bytes(const ta_lm_weights* w, int B, int S, int n_labelled, int with_backward, long long* bytes) {
    TA_REQUIRE(w && bytes, "null");
    LmBufs b;
    // with_backward: 0 forward only, 1 backward to the inputs (frozen LM), 2 backward + weight gradients (unfrozen LM)
    *bytes = lm_carve(w, B, S, n_labelled, with_backward != 0, nullptr, 0, &b, with_backward == 2) + 256;
    return 0;
}

TA_API int ta_lm_forward_backward(const ta_lm_weights* w, const ta_lm_step_args* a, void* stream) {
    TA_REQUIRE(w && a && a->inputs_embeds && a->workspace && a->loss, "ta_lm_forward_backward: null pointer");
    TA_REQUIRE(w->head_dim == 128, "Qwen3 path: head_dim must be 128");
    TA_REQUIRE(a->S <= w->max_pos, "sequence length %d exceeds the rotary table (%d)", a->S, w->max_pos);
    TA_REQUIRE(!a->with_backward || a->d_inputs_embeds, "with_backward needs d_inputs_embeds");
    TA_REQUIRE(a->n_labelled == 0 || (a->label_rows && a->label_targets), "label rows / targets missing");
    TA_REQUIRE(w->lora_pad == 0 || w->lora_pad == 128, "lora_pad must be 0 or 128");
    const int P = w->lora_pad;
    TA_REQUIRE(!(P && a->with_backward) || a->lora_grads, "LoRA backward needs the gradient pointer table");
    TA_REQUIRE(!a->k_cache || (a->v_cache && a->S <= a->cache_max_seq), "prefill: v_cache missing or S %d > cache_max_seq %d", a->S,
               a->cache_max_seq);
    TA_REQUIRE(!(a->position_ids || a->kv_start) || !a->with_backward, "position_ids / kv_start (left-padded prompts) are forward-only");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int B = a->B, S = a->S, nl = a->n_labelled;
    const long long M = (long long)B * S;
    const int D = w->dim, F = w->ffn, Hq = w->n_q_heads, Hkv = w->n_kv_heads, hd = w->head_dim;
    const int QD = Hq * hd, KD = Hkv * hd, QKV = QD + 2 * KD, QK = QD + KD;
    const int Lyr = w->n_layers;
    const long long ldX = D + P, ldH = F + P, ldAtt = QD + P;
    LmBufs b;
    memset(&b, 0, sizeof(b));
    float* const* G_all = a->with_backward ? a->lm_grads : nullptr;      // unfrozen LM: per-layer weight-gradient outputs
    TA_REQUIRE(!(G_all && P), "LoRA adapters and an unfrozen LM are mutually exclusive (asr_modeling.py:251-301)");
    TA_REQUIRE(!G_all || (a->d_embed && a->d_final_norm && a->input_ids), "unfrozen LM: d_embed / d_final_norm / input_ids missing");
    const long long need = lm_carve(w, B, S, nl, a->with_backward != 0, a->workspace, a->workspace_bytes, &b, G_all != nullptr);
    b.lora_zeroed = a->lora_grads_zeroed;
    TA_REQUIRE(need <= a->workspace_bytes, "LM workspace too small: need %lld, have %lld", need, a->workspace_bytes);
    if (M == 0) return 0;
    const long long sl = b.stride_layers;
    const float scale = 1.0f / sqrtf((float)hd);
    const long long lse_n = (long long)B * Hq * S;

    // ------------------------------ forward ------------------------------
    const float* x_in = a->inputs_embeds;
    for (int l = 0; l < Lyr; ++l) {
        const void* const* Lw = w->layers + (long long)l * TA_LM_PTRS_PER_LAYER;
        bf16* qkv = b.qkv + sl * l * M * QKV;
        bf16* qk = b.qk + sl * l * M * QK;
        bf16* att = b.att + sl * l * M * ldAtt;
        float* lse = b.lse + sl * l * lse_n;
        float* x_mid = b.resid_mid + sl * l * M * D;
        bf16* gu = b.gu + sl * l * M * 2 * F;
        bf16* lt = P ? b.lt + sl * l * 3 * M * P : nullptr;
        float* x_out = b.resid + (a->with_backward ? (long long)(l + 1) : (long long)((l + 1) & 1)) * M * D;

        RUN(k_rmsnorm_f32(x_in, (const float*)Lw[TA_LM_LN1_W], b.xn, nullptr, M, D, w->eps, st, ldX));
        if (P) {   // t_qkv = xn A_qkv^T into the extra columns (+ a copy kept for the B gradients)
            RUN(plain(b.xn, ldX, Lw[TA_LM_LORA_A_QKV], D, M, P, D, b.xn + D, ldX, st));
            TA_CHECK_CUDA(cudaMemcpy2DAsync(lt, P * 2, b.xn + D, ldX * 2, P * 2, M, cudaMemcpyDeviceToDevice, st));
        }
        RUN(gemm(b.xn, ldX, Lw[TA_LM_WQKV], ldX, M, QKV, D + P, TA_EPI_BF16, qkv, QKV, nullptr, nullptr, nullptr, 0, nullptr, 0, st));
        RUN(k_lm_qknorm_rope_fwd(qkv, qk, (const float*)Lw[TA_LM_QNORM_W], (const float*)Lw[TA_LM_KNORM_W], w->rope_cos, w->rope_sin,
                                 M, S, Hq, Hkv, w->eps, st, a->position_ids));
        if (a->k_cache) {   // prefill of generate(): keep the prompt's roped keys and values for the decode steps
            const long long per_layer = (long long)B * a->cache_max_seq * KD;
            RUN(k_kv_cache_store(qk + QD, QK, qkv + QK, QKV, reinterpret_cast<bf16*>(a->k_cache) + l * per_layer,
                                 reinterpret_cast<bf16*>(a->v_cache) + l * per_layer, B, S, KD, a->cache_max_seq, st));
        }
        if (a->kv_start) {   // left-padded prompts: the tcgen05 kernel masks the padding keys per sequence
            int handled = 0;
            RUN(k_attn_tc_fwd(qk, qk + QD, qkv + QK, att, lse, B, S, Hq, Hkv, hd, QK, QK, QKV, ldAtt, 1, scale, st, &handled, a->kv_start));
            TA_REQUIRE(handled, "left-padded attention: shape not supported by the tcgen05 kernel");
        } else {
            RUN(ta_attn_fwd(qk, qk + QD, qkv + QK, att, lse, B, S, Hq, Hkv, hd, QK, QK, QKV, ldAtt, 1, scale, st));
        }
        if (P) RUN(plain(att, ldAtt, Lw[TA_LM_LORA_A_O], QD, M, P, QD, att + QD, ldAtt, st));
        RUN(gemm(att, ldAtt, Lw[TA_LM_WO], ldAtt, M, D, QD + P, TA_EPI_F32_RESID, x_mid, D, nullptr, x_in, nullptr, 0, nullptr, 0, st));
        RUN(k_rmsnorm_f32(x_mid, (const float*)Lw[TA_LM_LN2_W], b.xn, nullptr, M, D, w->eps, st, ldX));
        if (P) {
            RUN(plain(b.xn, ldX, Lw[TA_LM_LORA_A_GU], D, M, P, D, b.xn + D, ldX, st));
            TA_CHECK_CUDA(cudaMemcpy2DAsync(lt + M * P, P * 2, b.xn + D, ldX * 2, P * 2, M, cudaMemcpyDeviceToDevice, st));
        }
        bf16* h_l = b.h + (b.h_layers ? (long long)l * M * ldH : 0);
        RUN(gemm(b.xn, ldX, Lw[TA_LM_WGU], ldX, M, 2 * F, D + P, TA_EPI_SWIGLU, h_l, ldH, nullptr, nullptr, gu, 2 * F, nullptr, 0, st));
        if (P) {
            RUN(plain(h_l, ldH, Lw[TA_LM_LORA_A_D], F, M, P, F, h_l + F, ldH, st));
            TA_CHECK_CUDA(cudaMemcpy2DAsync(lt + 2 * M * P, P * 2, h_l + F, ldH * 2, P * 2, M, cudaMemcpyDeviceToDevice, st));
        }
        RUN(gemm(h_l, ldH, Lw[TA_LM_WD], ldH, M, D, F + P, TA_EPI_F32_RESID, x_out, D, nullptr, x_mid, nullptr, 0, nullptr, 0, st));
        x_in = x_out;
    }
    const float* x_final = x_in;
    if (a->final_hidden)   // pre-final-norm hidden states for callers that need logits of arbitrary rows (generate)
        TA_CHECK_CUDA(cudaMemcpyAsync(a->final_hidden, x_final, sizeof(float) * M * D, cudaMemcpyDeviceToDevice, st));

    // ------------------------------ head: final norm on the labelled rows, lm_head, CE ------------------------------
    if (nl > 0) {
        RUN(k_rmsnorm_f32(x_final, w->final_norm_w, b.hl, a->label_rows, nl, D, w->eps, st));
        RUN(gemm(b.hl, D, w->embed_bf16, D, nl, (int)w->vocab_pad, D, TA_EPI_BF16, b.logits, w->vocab_pad, nullptr, nullptr, nullptr,
                 0, nullptr, 0, st));
        // per-row losses always go through a buffer: the batch loss is then reduced in a fixed order in double precision
        RUN(k_ce_fwd_bwd(b.logits, w->vocab_pad, a->label_targets, nl, (int)w->vocab, (int)w->vocab_pad, a->inv_num_items, a->loss,
                         a->row_loss ? a->row_loss : b.row_loss, a->with_backward, st));
    }
    if (!a->with_backward) return 0;

    // ------------------------------ backward ------------------------------
    float* dx = a->d_inputs_embeds;
    TA_CHECK_CUDA(cudaMemsetAsync(dx, 0, sizeof(float) * M * D, st));
    if (nl > 0) {
        // d(normed hidden) = d(logits) . E     [nl, D]
        RUN(gemm(b.logits, w->vocab_pad, w->embed_bf16_t, w->vocab_pad, nl, D, (int)w->vocab_pad, TA_EPI_BF16, b.dhl, D, nullptr,
                 nullptr, nullptr, 0, nullptr, 0, st));
        RUN(k_rmsnorm_f32_bwd(b.dhl, x_final, w->final_norm_w, dx, a->label_rows, nl, D, w->eps, 0, st));
        if (G_all) {
            RUN(k_rmsnorm_dw(b.dhl, x_final, a->label_rows, nl, D, w->eps, a->d_final_norm, st));
            // tied lm_head: d(embed)[v, :] = sum_rows d(logits)[row, v] * normed_hidden[row, :]  (overwrites; the embed_tokens
            // contribution is scatter-added below)
            RUN(full_wgrad(b, nl, b.logits, w->vocab_pad, (int)w->vocab_pad, b.hl, D, D, a->d_embed, st));
        }
    } else if (G_all) {
        TA_CHECK_CUDA(cudaMemsetAsync(a->d_embed, 0, sizeof(float) * w->vocab_pad * D, st));
    }
    RUN(k_cast_rows_f32_bf16(dx, b.dxb, M, D, ldX, st));   // later bf16 copies of dx come out of the RMSNorm-backward kernels
    const long long ldBig = ((2 * F > QKV) ? 2 * F : QKV) + P;
    for (int l = Lyr - 1; l >= 0; --l) {
        const void* const* Lw = w->layers + (long long)l * TA_LM_PTRS_PER_LAYER;
        float* const* Lg = P ? a->lora_grads + (long long)l * TA_LM_LORA_GRADS_PER_LAYER : nullptr;
        const bf16* qkv = b.qkv + (long long)l * M * QKV;
        const bf16* qk = b.qk + (long long)l * M * QK;
        bf16* att = b.att + (long long)l * M * ldAtt;
        const float* lse = b.lse + (long long)l * lse_n;
        const float* x_mid = b.resid_mid + (long long)l * M * D;
        const bf16* gu = b.gu + (long long)l * M * 2 * F;
        const bf16* lt = P ? b.lt + (long long)l * 3 * M * P : nullptr;
        const float* x_l = (l == 0) ? a->inputs_embeds : (b.resid + (long long)l * M * D);

        // ---- MLP branch: b.dxb = bf16(d x_out) ----
        bf16* h_b = b.h + (b.h_layers ? (long long)l * M * ldH : 0);
        if (P) {
            RUN(plain(b.dxb, ldX, Lw[TA_LM_LORA_BT_D], D, M, P, D, b.dxb + D, ldX, st));          // u_d = dy (s B_d)
            if (!b.h_layers) RUN(k_swiglu_h(gu, h_b, M, F, ldH, st));                              // x of down_proj (kept per layer otherwise)
            RUN(lora_wgrad(b, M, P, b.dxb, ldX, D, lt + 2 * M * P, P, b.dxb + D, ldX, h_b, ldH, F, Lg[TA_LM_LORA_DA_D],
                           Lg[TA_LM_LORA_DB_D], st));
        }
        float* const* Gl = G_all ? G_all + (long long)l * TA_LM_GRADS_PER_LAYER : nullptr;
        if (Gl) {   // down_proj: x = h = silu(gate) * up, kept per layer by the forward
            if (!b.h_layers) RUN(k_swiglu_h(gu, h_b, M, F, ldH, st));
            RUN(full_wgrad(b, M, b.dxb, ldX, D, h_b, ldH, F, Gl[TA_LM_G_WD], st));
        }
        RUN(gemm(b.dxb, ldX, Lw[TA_LM_WD_T], ldX, M, F, D + P, TA_EPI_SWIGLU_BWD, b.big, ldBig, nullptr, nullptr, nullptr, 0, gu, 2 * F,
                 st));
        if (Gl) {   // gate / up (interleaved 64-row blocks, like the packed operand): x = RMSNorm(x_mid)
            RUN(k_rmsnorm_f32(x_mid, (const float*)Lw[TA_LM_LN2_W], b.xn, nullptr, M, D, w->eps, st, ldX));
            RUN(full_wgrad(b, M, b.big, ldBig, 2 * F, b.xn, ldX, D, Gl[TA_LM_G_WGU], st));
        }
        if (P) {
            RUN(plain(b.big, ldBig, Lw[TA_LM_LORA_BT_GU], 2 * F, M, P, 2 * F, b.big + 2 * F, ldBig, st));
            RUN(k_rmsnorm_f32(x_mid, (const float*)Lw[TA_LM_LN2_W], b.xn, nullptr, M, D, w->eps, st, ldX));   // x of gate/up
            RUN(lora_wgrad(b, M, P, b.big, ldBig, 2 * F, lt + M * P, P, b.big + 2 * F, ldBig, b.xn, ldX, D, Lg[TA_LM_LORA_DA_GU],
                           Lg[TA_LM_LORA_DB_GU], st));
        }
        RUN(gemm(b.big, ldBig, Lw[TA_LM_WGU_T], 2 * F + P, M, D, 2 * F + P, TA_EPI_BF16, b.dxn, D, nullptr, nullptr, nullptr, 0, nullptr,
                 0, st));
        if (Gl) RUN(k_rmsnorm_dw(b.dxn, x_mid, nullptr, M, D, w->eps, Gl[TA_LM_G_LN2], st));
        RUN(k_rmsnorm_f32_bwd(b.dxn, x_mid, (const float*)Lw[TA_LM_LN2_W], dx, nullptr, M, D, w->eps, 1, st, b.dxb, ldX));
        // ---- attention branch ----
        if (Gl) RUN(full_wgrad(b, M, b.dxb, ldX, D, att, ldAtt, QD, Gl[TA_LM_G_WO], st));
        if (P) {
            RUN(plain(b.dxb, ldX, Lw[TA_LM_LORA_BT_O], D, M, P, D, b.dxb + D, ldX, st));
            RUN(lora_wgrad(b, M, P, b.dxb, ldX, D, att + QD, ldAtt, b.dxb + D, ldX, att, ldAtt, QD, Lg[TA_LM_LORA_DA_O],
                           Lg[TA_LM_LORA_DB_O], st));
        }
        // d(att) = d(x) Wo, and in the same epilogue D = rowsum(d(att) o att) per head (one epilogue thread owns a row's 128-wide head):
        // the attention backward's preparation pass (a 122 MB read per layer) is gone
        const bool rowdot = g_fuse_attn_dsum && QD % 256 == 0;
        if (rowdot) {
            ta_gemm_epilogue e{};
            e.out = b.datt; e.ldo = QD; e.aux = att; e.ldaux = ldAtt; e.out2 = b.dsum; e.rope_seq = S;
            e.zero_f32 = b.dq; e.ld_zero = QD;       // the dQ accumulator is cleared here too: no separate 122 MB memset per layer
            RUN(ta_gemm_bf16(b.dxb, ldX, Lw[TA_LM_WO_T], ldX, (int)M, QD, D + P, TA_EPI_BF16_ROWDOT, &e, st));
        } else {
            RUN(gemm(b.dxb, ldX, Lw[TA_LM_WO_T], ldX, M, QD, D + P, TA_EPI_BF16, b.datt, QD, nullptr, nullptr, nullptr, 0, nullptr, 0, st));
        }
        RUN(k_attn_bwd(qk, qk + QD, qkv + QK, att, b.datt, lse, b.dsum, b.dq, b.dk, b.dv, B, S, Hq, Hkv, hd, QK, QK, QKV, ldAtt, QD, QD,
                       KD, KD, 1, scale, st, rowdot ? 3 : 0));
        RUN(k_lm_qknorm_rope_bwd(qkv, b.dq, b.dk, b.dv, b.big, (const float*)Lw[TA_LM_QNORM_W], (const float*)Lw[TA_LM_KNORM_W],
                                 w->rope_cos, w->rope_sin, M, S, Hq, Hkv, w->eps, st, ldBig));
        if (Gl) {
            RUN(k_qknorm_dw(qkv, b.dq, b.dk, w->rope_cos, w->rope_sin, M, S, Hq, Hkv, w->eps, Gl[TA_LM_G_QNORM], Gl[TA_LM_G_KNORM], st));
            RUN(k_rmsnorm_f32(x_l, (const float*)Lw[TA_LM_LN1_W], b.xn, nullptr, M, D, w->eps, st, ldX));     // x of q/k/v
            RUN(full_wgrad(b, M, b.big, ldBig, QKV, b.xn, ldX, D, Gl[TA_LM_G_WQKV], st));
        }
        if (P) {
            RUN(plain(b.big, ldBig, Lw[TA_LM_LORA_BT_QKV], QKV, M, P, QKV, b.big + QKV, ldBig, st));
            RUN(k_rmsnorm_f32(x_l, (const float*)Lw[TA_LM_LN1_W], b.xn, nullptr, M, D, w->eps, st, ldX));     // x of q/k/v
            RUN(lora_wgrad(b, M, P, b.big, ldBig, QKV, lt, P, b.big + QKV, ldBig, b.xn, ldX, D, Lg[TA_LM_LORA_DA_QKV],
                           Lg[TA_LM_LORA_DB_QKV], st));
        }
        RUN(gemm(b.big, ldBig, Lw[TA_LM_WQKV_T], QKV + P, M, D, QKV + P, TA_EPI_BF16, b.dxn, D, nullptr, nullptr, nullptr, 0, nullptr, 0,
                 st));
        if (Gl) RUN(k_rmsnorm_dw(b.dxn, x_l, nullptr, M, D, w->eps, Gl[TA_LM_G_LN1], st));
        RUN(k_rmsnorm_f32_bwd(b.dxn, x_l, (const float*)Lw[TA_LM_LN1_W], dx, nullptr, M, D, w->eps, 1, st, l > 0 ? b.dxb : nullptr, ldX));
    }
    if (G_all)   // embed_tokens: the text positions' input gradients flow into the (tied) table
        RUN(k_embed_grad_scatter(a->input_ids, dx, a->d_embed, M, D, w->vocab, a->audio_token_id, st));
    return 0;
}

TA_API int ta_lm_hidden_to_logits(const ta_lm_weights* w, const float* hidden, const int* rows, int n_rows, void* normed_ws,
                                  void* logits, void* stream) {
    TA_REQUIRE(w && hidden && normed_ws && logits, "ta_lm_hidden_to_logits: null pointer");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (n_rows == 0) return 0;
    RUN(k_rmsnorm_f32(hidden, w->final_norm_w, (bf16*)normed_ws, rows, n_rows, w->dim, w->eps, st));
    RUN(gemm(normed_ws, w->dim, w->embed_bf16, w->dim, n_rows, (int)w->vocab_pad, w->dim, TA_EPI_BF16, logits, w->vocab_pad, nullptr,
             nullptr, nullptr, 0, nullptr, 0, st));
    return 0;
}

// =============================================================================================================
// greedy decode with a KV cache: one token per sequence (csrc/decode.cu holds the kernels)
// =============================================================================================================
namespace {
struct DecBufs {
    float *x0, *x1, *x2;   // [B, D] fp32 residual stream (layer in / mid / out)
    bf16 *xn, *qkv, *q, *att, *h;
    float* part;           // [4, B, D] split-K partial sums of o_proj / down_proj
};
long long dec_carve(const ta_lm_weights* w, int B, void* ws, long long cap, DecBufs* b) {
    const long long D = w->dim, F = w->ffn, P = w->lora_pad;
    const long long QD = (long long)w->n_q_heads * w->head_dim, KD = (long long)w->n_kv_heads * w->head_dim;
    Carver c(ws, cap);
    b->x0 = c.take<float>(B * D);
    b->x1 = c.take<float>(B * D);
    b->x2 = c.take<float>(B * D);
    b->xn = c.take<bf16>(B * (D + P));
    b->qkv = c.take<bf16>(B * (QD + 2 * KD));
    b->q = c.take<bf16>(B * QD);
    b->att = c.take<bf16>(B * (QD + P));
    b->h = c.take<bf16>(B * (F + P));
    b->part = c.take<float>(4 * B * D);
    return c.off;
}
}  // namespace

TA_API int ta_lm_decode_workspace_bytes(const ta_lm_weights* w, int B, long long* bytes) {
    TA_REQUIRE(w && bytes, "null");
    DecBufs b;
    *bytes = dec_carve(w, B, nullptr, 0, &b) + 256;
    return 0;
}

TA_API int ta_lm_decode_step(const ta_lm_weights* w, const long long* ids, int* pos, int pos_host, void* k_cache, void* v_cache,
                             int cache_max_seq, int B, void* workspace, long long workspace_bytes, void* logits, long long* next_ids,
                             const int* kv_start, void* stream) {
    TA_REQUIRE(w && ids && pos && k_cache && v_cache && workspace && logits && next_ids, "ta_lm_decode_step: null pointer");
    TA_REQUIRE(w->head_dim == 128, "Qwen3 path: head_dim must be 128");
    TA_REQUIRE(B >= 1 && B <= 32, "decode step: batch %d not in [1, 32]", B);
    TA_REQUIRE(pos_host >= 0 && pos_host < cache_max_seq && pos_host < w->max_pos, "decode step: position %d outside the cache (%d) / rotary table (%d)",
               pos_host, cache_max_seq, w->max_pos);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int D = w->dim, F = w->ffn, Hq = w->n_q_heads, Hkv = w->n_kv_heads, hd = w->head_dim, P = w->lora_pad;
    const int QD = Hq * hd, KD = Hkv * hd, QKV = QD + 2 * KD;
    const long long ldX = D + P, ldH = F + P, ldAtt = QD + P;
    DecBufs b;
    const long long need = dec_carve(w, B, workspace, workspace_bytes, &b);
    TA_REQUIRE(need <= workspace_bytes, "decode workspace too small: need %lld, have %lld", need, workspace_bytes);
    const float scale = 1.0f / sqrtf((float)hd);
    const long long per_layer = (long long)B * cache_max_seq * KD;

    RUN(k_embed_rows(ids, w->embed_f32, b.x0, B, D, w->vocab, st));
    float* x_in = b.x0;
    float* x_mid = b.x1;
    float* x_out = b.x2;
    const int sp_o = k_skinny_splits(D, QD + P), sp_d = k_skinny_splits(D, F + P);
    const void* const* L0 = w->layers;
    RUN(k_decode_resid_rmsnorm(x_in, nullptr, 0, B, nullptr, (const float*)L0[TA_LM_LN1_W], b.xn, D, w->eps, ldX, st));
    for (int l = 0; l < w->n_layers; ++l) {
        const void* const* Lw = w->layers + (long long)l * TA_LM_PTRS_PER_LAYER;
        bf16* kc = reinterpret_cast<bf16*>(k_cache) + l * per_layer;
        bf16* vc = reinterpret_cast<bf16*>(v_cache) + l * per_layer;
        if (P) RUN(k_skinny_gemm(b.xn, ldX, (const bf16*)Lw[TA_LM_LORA_A_QKV], D, B, P, D, TA_SKINNY_BF16, b.xn + D, ldX, nullptr, st));
        RUN(k_skinny_gemm(b.xn, ldX, (const bf16*)Lw[TA_LM_WQKV], ldX, B, QKV, D + P, TA_SKINNY_BF16, b.qkv, QKV, nullptr, st));
        RUN(k_decode_attn(b.qkv, nullptr, kc, vc, b.att, ldAtt, (const float*)Lw[TA_LM_QNORM_W], (const float*)Lw[TA_LM_KNORM_W], w->rope_cos,
                          w->rope_sin, pos, B, Hq, Hkv, cache_max_seq, w->eps, scale, st, kv_start));
        if (P) RUN(k_skinny_gemm(b.att, ldAtt, (const bf16*)Lw[TA_LM_LORA_A_O], QD, B, P, QD, TA_SKINNY_BF16, b.att + QD, ldAtt, nullptr, st));
        // o_proj / down_proj have only dim / 16 = 64 row tiles: split K so that the whole GPU streams; the partial sums are
        // reduced (fixed order), rounded to bf16 and added to the residual by the RMSNorm kernel that follows anyway
        RUN(k_skinny_gemm(b.att, ldAtt, (const bf16*)Lw[TA_LM_WO], ldAtt, B, D, QD + P, TA_SKINNY_PARTIAL, b.part, D, nullptr, st, sp_o));
        RUN(k_decode_resid_rmsnorm(x_in, b.part, sp_o, B, x_mid, (const float*)Lw[TA_LM_LN2_W], b.xn, D, w->eps, ldX, st));
        if (P) RUN(k_skinny_gemm(b.xn, ldX, (const bf16*)Lw[TA_LM_LORA_A_GU], D, B, P, D, TA_SKINNY_BF16, b.xn + D, ldX, nullptr, st));
        RUN(k_skinny_gemm(b.xn, ldX, (const bf16*)Lw[TA_LM_WGU], ldX, B, 2 * F, D + P, TA_SKINNY_SWIGLU, b.h, ldH, nullptr, st));
        if (P) RUN(k_skinny_gemm(b.h, ldH, (const bf16*)Lw[TA_LM_LORA_A_D], F, B, P, F, TA_SKINNY_BF16, b.h + F, ldH, nullptr, st));
        RUN(k_skinny_gemm(b.h, ldH, (const bf16*)Lw[TA_LM_WD], ldH, B, D, F + P, TA_SKINNY_PARTIAL, b.part, D, nullptr, st, sp_d));
        const bool last = (l + 1 == w->n_layers);
        const float* next_norm = last ? w->final_norm_w : (const float*)(Lw + TA_LM_PTRS_PER_LAYER)[TA_LM_LN1_W];
        RUN(k_decode_resid_rmsnorm(x_mid, b.part, sp_d, B, x_out, next_norm, b.xn, D, w->eps, last ? (long long)D : ldX, st));
        float* t = x_in;
        x_in = x_out;
        x_out = t;
    }
    RUN(k_skinny_gemm(b.xn, D, (const bf16*)w->embed_bf16, D, B, (int)w->vocab_pad, D, TA_SKINNY_BF16, logits, w->vocab_pad, nullptr, st));
    RUN(k_argmax_rows((const bf16*)logits, w->vocab_pad, B, (int)w->vocab, next_ids, pos, st));
    return 0;
}
