// Window attention of the QFormer projector (reference: tiny_audio/projectors.py:431-475 -> HF:models/blip_2/modeling_blip_2.py:579-634,
// Blip2QFormerMultiHeadAttention): per 15-frame window, 3 learned queries attend to 3 keys (self-attention) or 15 keys
// (cross-attention over the window's encoder frames), 16 heads x 80.  The contraction sizes (3 x 15 x 80) are far below one
// tensor-core tile, and the whole problem is 270 MB of bf16 per pass at batch 32 x 30 s: this is HBM-bound SIMT work.
//
// One WARP per (window, head), lanes over the head dimension (coalesced 2-byte accesses, conflict-free), scores reduced with
// xor-shuffles, softmax / dropout mask / P.V in registers.  bf16 in and out (the producers and consumers are tcgen05 GEMMs that
// read and write bf16), fp32 arithmetic.  No atomics: every (window, head) owns its slices of dq, dk, dv.
#include "kernels.cuh"

namespace {

constexpr int MAXQ = 4;       // queries per window  (QFormer: window 15 / downsample 5 = 3)
constexpr int MAXK = 16;      // keys per window     (QFormer: 3 self, 15 cross)
constexpr int DPL = 3;        // head-dim elements per lane -> head_dim <= 96 (QFormer: 1280 / 16 = 80)
constexpr int WARPS = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// BWD = false: out = dropout(softmax(scale q k^T)) v.   BWD = true: dq, dk, dv from d_out (probabilities recomputed).
// Lane l owns head-dim elements l, l + 32, l + 64; the optional dropout mask holds 0 or 1 / (1 - p), as F.dropout produces.
template <bool BWD>
__global__ void __launch_bounds__(WARPS * 32)
window_attn_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v, const float* __restrict__ mask,
                   const bf16* __restrict__ d_out, bf16* __restrict__ out, bf16* __restrict__ dq, bf16* __restrict__ dk,
                   bf16* __restrict__ dv, long long n_pairs, int nq, int nk, int heads, int hd, float scale) {
    const long long pair = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5);      // (window, head)
    if (pair >= n_pairs) return;
    const int lane = threadIdx.x & 31;
    const long long win = pair / heads;
    const int head = (int)(pair % heads);
    const long long H = (long long)heads * hd;
    const bf16* qb = q + win * nq * H + (long long)head * hd;
    const bf16* kb = k + win * nk * H + (long long)head * hd;
    const bf16* vb = v + win * nk * H + (long long)head * hd;

    float qr[MAXQ][DPL], p[MAXQ][MAXK];
#pragma unroll
    for (int i = 0; i < MAXQ; ++i)
#pragma unroll
        for (int t = 0; t < DPL; ++t) {
            const int d = lane + 32 * t;
            qr[i][t] = (i < nq && d < hd) ? __bfloat162float(qb[i * H + d]) : 0.0f;
        }
    // S = scale * Q K^T
#pragma unroll
    for (int j = 0; j < MAXK; ++j) {
        float kr[DPL];
#pragma unroll
        for (int t = 0; t < DPL; ++t) {
            const int d = lane + 32 * t;
            kr[t] = (j < nk && d < hd) ? __bfloat162float(kb[j * H + d]) : 0.0f;
        }
#pragma unroll
        for (int i = 0; i < MAXQ; ++i) {
            float s = 0.0f;
#pragma unroll
            for (int t = 0; t < DPL; ++t) s = fmaf(qr[i][t], kr[t], s);
            p[i][j] = warp_sum(s) * scale;
        }
    }
    // softmax over keys (every lane holds the full, identical score matrix)
#pragma unroll
    for (int i = 0; i < MAXQ; ++i) {
        float m = -INFINITY;
#pragma unroll
        for (int j = 0; j < MAXK; ++j)
            if (j < nk) m = fmaxf(m, p[i][j]);
        float sum = 0.0f;
#pragma unroll
        for (int j = 0; j < MAXK; ++j) {
            p[i][j] = (j < nk) ? __expf(p[i][j] - m) : 0.0f;
            sum += p[i][j];
        }
        const float inv = 1.0f / sum;
#pragma unroll
        for (int j = 0; j < MAXK; ++j) p[i][j] *= inv;
    }
    const float* mk = mask ? mask + pair * nq * nk : nullptr;      // [window, head, nq, nk]

    if (!BWD) {
        // O = dropout(P) V
        float o[MAXQ][DPL];
#pragma unroll
        for (int i = 0; i < MAXQ; ++i)
#pragma unroll
            for (int t = 0; t < DPL; ++t) o[i][t] = 0.0f;
#pragma unroll
        for (int j = 0; j < MAXK; ++j) {
            if (j < nk) {
                float vr[DPL];
#pragma unroll
                for (int t = 0; t < DPL; ++t) {
                    const int d = lane + 32 * t;
                    vr[t] = d < hd ? __bfloat162float(vb[j * H + d]) : 0.0f;
                }
#pragma unroll
                for (int i = 0; i < MAXQ; ++i) {
                    const float w = (mk && i < nq) ? p[i][j] * mk[i * nk + j] : p[i][j];
#pragma unroll
                    for (int t = 0; t < DPL; ++t) o[i][t] = fmaf(w, vr[t], o[i][t]);
                }
            }
        }
        bf16* ob = out + win * nq * H + (long long)head * hd;
#pragma unroll
        for (int i = 0; i < MAXQ; ++i)
#pragma unroll
            for (int t = 0; t < DPL; ++t) {
                const int d = lane + 32 * t;
                if (i < nq && d < hd) ob[i * H + d] = __float2bfloat16(o[i][t]);
            }
        return;
    }

    // ---------------- backward ----------------
    // dO -> dV_j = sum_i Pd[i][j] dO_i ;  dPd[i][j] = dO_i . V_j ;  dP = dPd * mask ;  dS = P * (dP - sum_j P dP) ;
    // dQ_i = scale * sum_j dS[i][j] K_j ;  dK_j = scale * sum_i dS[i][j] Q_i          (Pd = dropout(P) = P * mask)
    const bf16* gb = d_out + win * nq * H + (long long)head * hd;
    float g[MAXQ][DPL];
#pragma unroll
    for (int i = 0; i < MAXQ; ++i)
#pragma unroll
        for (int t = 0; t < DPL; ++t) {
            const int d = lane + 32 * t;
            g[i][t] = (i < nq && d < hd) ? __bfloat162float(gb[i * H + d]) : 0.0f;
        }
    float dp[MAXQ][MAXK];
    bf16* dvb = dv + win * nk * H + (long long)head * hd;
#pragma unroll
    for (int j = 0; j < MAXK; ++j) {
        float vr[DPL], dvr[DPL];
#pragma unroll
        for (int t = 0; t < DPL; ++t) {
            const int d = lane + 32 * t;
            vr[t] = (j < nk && d < hd) ? __bfloat162float(vb[j * H + d]) : 0.0f;
            dvr[t] = 0.0f;
        }
#pragma unroll
        for (int i = 0; i < MAXQ; ++i) {
            const float mij = (mk && i < nq && j < nk) ? mk[i * nk + j] : 1.0f;
            float s = 0.0f;
#pragma unroll
            for (int t = 0; t < DPL; ++t) {
                s = fmaf(g[i][t], vr[t], s);
                dvr[t] = fmaf(p[i][j] * mij, g[i][t], dvr[t]);
            }
            dp[i][j] = warp_sum(s) * mij;
        }
        if (j < nk) {
#pragma unroll
            for (int t = 0; t < DPL; ++t) {
                const int d = lane + 32 * t;
                if (d < hd) dvb[j * H + d] = __float2bfloat16(dvr[t]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < MAXQ; ++i) {          // dS in place of dP (already multiplied by scale for dQ / dK)
        float dot = 0.0f;
#pragma unroll
        for (int j = 0; j < MAXK; ++j) dot = fmaf(p[i][j], dp[i][j], dot);
#pragma unroll
        for (int j = 0; j < MAXK; ++j) dp[i][j] = p[i][j] * (dp[i][j] - dot) * scale;
    }
    float dqr[MAXQ][DPL];
#pragma unroll
    for (int i = 0; i < MAXQ; ++i)
#pragma unroll
        for (int t = 0; t < DPL; ++t) dqr[i][t] = 0.0f;
    bf16* dkb = dk + win * nk * H + (long long)head * hd;
#pragma unroll
    for (int j = 0; j < MAXK; ++j) {
        if (j < nk) {
            float kr[DPL], dkr[DPL];
#pragma unroll
            for (int t = 0; t < DPL; ++t) {
                const int d = lane + 32 * t;
                kr[t] = d < hd ? __bfloat162float(kb[j * H + d]) : 0.0f;
                dkr[t] = 0.0f;
            }
#pragma unroll
            for (int i = 0; i < MAXQ; ++i)
#pragma unroll
                for (int t = 0; t < DPL; ++t) {
                    dqr[i][t] = fmaf(dp[i][j], kr[t], dqr[i][t]);
                    dkr[t] = fmaf(dp[i][j], qr[i][t], dkr[t]);
                }
#pragma unroll
            for (int t = 0; t < DPL; ++t) {
                const int d = lane + 32 * t;
                if (d < hd) dkb[j * H + d] = __float2bfloat16(dkr[t]);
            }
        }
    }
    bf16* dqb = dq + win * nq * H + (long long)head * hd;
#pragma unroll
    for (int i = 0; i < MAXQ; ++i)
#pragma unroll
        for (int t = 0; t < DPL; ++t) {
            const int d = lane + 32 * t;
            if (i < nq && d < hd) dqb[i * H + d] = __float2bfloat16(dqr[i][t]);
        }
}

int check_shape(const char* who, long long n_win, int nq, int nk, int heads, int hd) {
    TA_REQUIRE(n_win > 0 && heads > 0, "%s: empty problem (windows %lld, heads %d)", who, n_win, heads);
    TA_REQUIRE(nq >= 1 && nq <= MAXQ, "%s: %d queries per window (supported: 1..%d)", who, nq, MAXQ);
    TA_REQUIRE(nk >= 1 && nk <= MAXK, "%s: %d keys per window (supported: 1..%d)", who, nk, MAXK);
    TA_REQUIRE(hd >= 1 && hd <= 32 * DPL, "%s: head_dim %d (supported: 1..%d)", who, hd, 32 * DPL);
    TA_REQUIRE(n_win * heads <= 0x7fffffffLL * WARPS, "%s: too many (window, head) pairs", who);
    return 0;
}

}  // namespace

// out bf16 [n_win, nq, heads*hd] = dropout(softmax(scale * q k^T)) v   per (window, head);  q bf16 [n_win, nq, heads*hd],
// k, v bf16 [n_win, nk, heads*hd];  drop_mask NULL or f32 [n_win, heads, nq, nk] holding 0 or 1/(1-p)
TA_API int ta_window_attn_fwd(const void* q, const void* k, const void* v, const float* drop_mask, void* out, long long n_win,
                              int nq, int nk, int heads, int head_dim, float scale, void* stream) {
    TA_REQUIRE(q && k && v && out, "ta_window_attn_fwd: null pointer");
    if (int rc = check_shape("ta_window_attn_fwd", n_win, nq, nk, heads, head_dim)) return rc;
    const long long pairs = n_win * heads;
    const unsigned grid = (unsigned)((pairs + WARPS - 1) / WARPS);
    window_attn_kernel<false><<<grid, WARPS * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        (const bf16*)q, (const bf16*)k, (const bf16*)v, drop_mask, nullptr, (bf16*)out, nullptr, nullptr, nullptr, pairs, nq, nk, heads,
        head_dim, scale);
    TA_LAUNCH_CHECK();
    return 0;
}

// gradients of the above (probabilities are recomputed from q, k): dq [n_win, nq, H], dk, dv [n_win, nk, H], all bf16
TA_API int ta_window_attn_bwd(const void* q, const void* k, const void* v, const float* drop_mask, const void* d_out, void* dq, void* dk,
                              void* dv, long long n_win, int nq, int nk, int heads, int head_dim, float scale, void* stream) {
    TA_REQUIRE(q && k && v && d_out && dq && dk && dv, "ta_window_attn_bwd: null pointer");
    if (int rc = check_shape("ta_window_attn_bwd", n_win, nq, nk, heads, head_dim)) return rc;
    const long long pairs = n_win * heads;
    const unsigned grid = (unsigned)((pairs + WARPS - 1) / WARPS);
    window_attn_kernel<true><<<grid, WARPS * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        (const bf16*)q, (const bf16*)k, (const bf16*)v, drop_mask, (const bf16*)d_out, nullptr, (bf16*)dq, (bf16*)dk, (bf16*)dv, pairs, nq,
        nk, heads, head_dim, scale);
    TA_LAUNCH_CHECK();
    return 0;
}
