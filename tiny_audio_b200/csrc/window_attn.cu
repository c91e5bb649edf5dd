// Window attention of the QFormer projector (reference: tiny_audio/projectors.py:431-475 -> HF:models/blip_2/modeling_blip_2.py:579-634,
// Blip2QFormerMultiHeadAttention): per 15-frame window, 3 learned queries attend to 3 keys (self-attention) or 15 keys
// (cross-attention over the window's encoder frames), 16 heads x 80.  The contraction sizes (3 x 15 x 80) are far below one
// tensor-core tile, and the whole problem is 270 MB of bf16 per pass at batch 32 x 30 s: this is HBM-bound SIMT work.
//
// One WARP per (window, head); bf16 in and out (the producers and consumers are tcgen05 GEMMs that read and write bf16), fp32
// arithmetic.  No atomics: every (window, head) owns its slices of dq, dk, dv.  Two formulations:
//   variant 1 (`window_attn_kernel`): lanes over the head dimension, every score reduced with a 5-step xor-shuffle butterfly,
//     whole P / dP matrices in registers (149 / 231 registers -> 8 warps per SM; measured 0.96 / 1.2 ms per launch at
//     3200 windows x 16 heads: shuffle-latency bound).  Any head_dim <= 96.
//   variant 2 (`window_attn2_kernel`): KEY PER LANE for the score products (lane j streams key row j with 16-byte loads and keeps
//     its nq scores; no shuffles in the contraction), softmax as lane reductions, probabilities / dS through 0.5 KB of shared
//     memory per warp, then (row, 8-wide chunk) PER LANE for P.V, dV, dK, dQ with 16-byte loads and stores.  head_dim % 8 == 0.
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

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const float2 x = __bfloat1622float2(h[t]);
        f[2 * t] = x.x;
        f[2 * t + 1] = x.y;
    }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int t = 0; t < 4; ++t) h[t] = __floats2bfloat162_rn(f[2 * t], f[2 * t + 1]);
    return u;
}
// row pointer as 16-byte chunks (head_dim % 8 == 0 and H % 8 == 0 make every chunk aligned)
__device__ __forceinline__ const uint4* chunks_of(const bf16* base, long long row, long long H) {
    return reinterpret_cast<const uint4*>(base + row * H);
}

template <bool BWD>
__global__ void __launch_bounds__(WARPS * 32)
window_attn2_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v, const float* __restrict__ mask,
                    const bf16* __restrict__ d_out, bf16* __restrict__ out, bf16* __restrict__ dq, bf16* __restrict__ dk,
                    bf16* __restrict__ dv, long long n_pairs, int nq, int nk, int heads, int hd, float scale) {
    __shared__ float s_p[WARPS][MAXQ * MAXK];       // dropout(P)   [i][j]
    __shared__ float s_ds[WARPS][MAXQ * MAXK];      // scale * dS   [i][j]   (backward only)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long pair = (long long)blockIdx.x * WARPS + warp;
    if (pair >= n_pairs) return;
    const long long win = pair / heads;
    const int head = (int)(pair % heads);
    const long long H = (long long)heads * hd;
    const int chunks = hd >> 3;
    const bf16* qb = q + win * nq * H + (long long)head * hd;
    const bf16* kb = k + win * nk * H + (long long)head * hd;
    const bf16* vb = v + win * nk * H + (long long)head * hd;
    const float* mk = mask ? mask + pair * nq * nk : nullptr;      // [window, head, nq, nk]
    const bool has_key = lane < nk;

    // ---- scores: lane j owns key j ----
    float p[MAXQ];
#pragma unroll
    for (int i = 0; i < MAXQ; ++i) p[i] = 0.0f;
    if (has_key) {
        const uint4* kr = chunks_of(kb, lane, H);
        for (int c = 0; c < chunks; ++c) {
            float kf[8];
            unpack8(kr[c], kf);
#pragma unroll
            for (int i = 0; i < MAXQ; ++i) {
                if (i < nq) {
                    float qf[8];
                    unpack8(chunks_of(qb, i, H)[c], qf);
#pragma unroll
                    for (int t = 0; t < 8; ++t) p[i] = fmaf(qf[t], kf[t], p[i]);
                }
            }
        }
    }
    // ---- softmax over the keys = over lanes [0, nk) ----
#pragma unroll
    for (int i = 0; i < MAXQ; ++i) {
        const float x = has_key ? p[i] * scale : -INFINITY;
        const float m = warp_max(x);
        const float e = has_key ? __expf(x - m) : 0.0f;
        p[i] = e / warp_sum(e);
    }

    if (!BWD) {
        if (has_key) {
#pragma unroll
            for (int i = 0; i < MAXQ; ++i)
                if (i < nq) s_p[warp][i * MAXK + lane] = mk ? p[i] * mk[i * nk + lane] : p[i];
        }
        __syncwarp();
        // O[i][chunk] = sum_j Pd[i][j] V[j][chunk]: one (query row, 8-wide chunk) per lane
        for (int slot = lane; slot < nq * chunks; slot += 32) {
            const int i = slot / chunks, c = slot - i * chunks;
            float o[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) o[t] = 0.0f;
            for (int j = 0; j < nk; ++j) {
                float vf[8];
                unpack8(chunks_of(vb, j, H)[c], vf);
                const float w = s_p[warp][i * MAXK + j];
#pragma unroll
                for (int t = 0; t < 8; ++t) o[t] = fmaf(w, vf[t], o[t]);
            }
            reinterpret_cast<uint4*>(out + win * nq * H + (long long)head * hd + (long long)i * H)[c] = pack8(o);
        }
        return;
    }

    // ---------------- backward ----------------
    const bf16* gb = d_out + win * nq * H + (long long)head * hd;
    float dp[MAXQ];
#pragma unroll
    for (int i = 0; i < MAXQ; ++i) dp[i] = 0.0f;
    if (has_key) {                  // dPd[i][j] = dO_i . V_j, lane j owns key j
        const uint4* vr = chunks_of(vb, lane, H);
        for (int c = 0; c < chunks; ++c) {
            float vf[8];
            unpack8(vr[c], vf);
#pragma unroll
            for (int i = 0; i < MAXQ; ++i) {
                if (i < nq) {
                    float gf[8];
                    unpack8(chunks_of(gb, i, H)[c], gf);
#pragma unroll
                    for (int t = 0; t < 8; ++t) dp[i] = fmaf(gf[t], vf[t], dp[i]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < MAXQ; ++i) {
        const float mij = (mk && has_key && i < nq) ? mk[i * nk + lane] : 1.0f;
        const float dpi = has_key ? dp[i] * mij : 0.0f;               // dP = dPd * mask
        const float dot = warp_sum(has_key ? p[i] * dpi : 0.0f);
        if (has_key && i < nq) {
            s_p[warp][i * MAXK + lane] = p[i] * mij;                   // dropout(P)
            s_ds[warp][i * MAXK + lane] = p[i] * (dpi - dot) * scale;  // softmax backward, scale folded in
        }
    }
    __syncwarp();
    // dV[j][chunk] = sum_i Pd[i][j] dO_i[chunk] ;  dK[j][chunk] = sum_i dS[i][j] Q_i[chunk]: one (key row, chunk) per lane
    for (int slot = lane; slot < nk * chunks; slot += 32) {
        const int j = slot / chunks, c = slot - j * chunks;
        float av[8], ak[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) av[t] = ak[t] = 0.0f;
        for (int i = 0; i < nq; ++i) {
            float gf[8], qf[8];
            unpack8(chunks_of(gb, i, H)[c], gf);
            unpack8(chunks_of(qb, i, H)[c], qf);
            const float wv = s_p[warp][i * MAXK + j], wk = s_ds[warp][i * MAXK + j];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                av[t] = fmaf(wv, gf[t], av[t]);
                ak[t] = fmaf(wk, qf[t], ak[t]);
            }
        }
        const long long off = win * nk * H + (long long)head * hd + (long long)j * H;
        reinterpret_cast<uint4*>(dv + off)[c] = pack8(av);
        reinterpret_cast<uint4*>(dk + off)[c] = pack8(ak);
    }
    // dQ[i][chunk] = sum_j dS[i][j] K_j[chunk]
    for (int slot = lane; slot < nq * chunks; slot += 32) {
        const int i = slot / chunks, c = slot - i * chunks;
        float a[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) a[t] = 0.0f;
        for (int j = 0; j < nk; ++j) {
            float kf[8];
            unpack8(chunks_of(kb, j, H)[c], kf);
            const float w = s_ds[warp][i * MAXK + j];
#pragma unroll
            for (int t = 0; t < 8; ++t) a[t] = fmaf(w, kf[t], a[t]);
        }
        reinterpret_cast<uint4*>(dq + win * nq * H + (long long)head * hd + (long long)i * H)[c] = pack8(a);
    }
}

// 1: lanes over head_dim (any head_dim); 2 (default): key per lane (head_dim % 8 == 0, else variant 1 runs).  Measured on a B200,
// QFormer recipe at batch 32 x 30 s: 133.8 ms/step with variant 1, 128.9 ms with variant 2 (profiles/r02_c01_*)
int g_window_attn_variant = 2;

bool use_variant2(int hd, const void* a, const void* b, const void* c, const void* d) {
    auto aligned = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    return g_window_attn_variant == 2 && hd % 8 == 0 && aligned(a) && aligned(b) && aligned(c) && aligned(d);
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
    auto kern = use_variant2(head_dim, q, k, v, out) ? window_attn2_kernel<false> : window_attn_kernel<false>;
    kern<<<grid, WARPS * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        (const bf16*)q, (const bf16*)k, (const bf16*)v, drop_mask, nullptr, (bf16*)out, nullptr, nullptr, nullptr, pairs, nq, nk, heads,
        head_dim, scale);
    TA_LAUNCH_CHECK();
    return 0;
}

// A/B switch between the two formulations (both parity-tested); returns the previous value
TA_API int ta_window_attn_set_variant(int variant) {
    const int prev = g_window_attn_variant;
    if (variant == 1 || variant == 2) g_window_attn_variant = variant;
    return prev;
}

// gradients of the above (probabilities are recomputed from q, k): dq [n_win, nq, H], dk, dv [n_win, nk, H], all bf16
TA_API int ta_window_attn_bwd(const void* q, const void* k, const void* v, const float* drop_mask, const void* d_out, void* dq, void* dk,
                              void* dv, long long n_win, int nq, int nk, int heads, int head_dim, float scale, void* stream) {
    TA_REQUIRE(q && k && v && d_out && dq && dk && dv, "ta_window_attn_bwd: null pointer");
    if (int rc = check_shape("ta_window_attn_bwd", n_win, nq, nk, heads, head_dim)) return rc;
    const long long pairs = n_win * heads;
    const unsigned grid = (unsigned)((pairs + WARPS - 1) / WARPS);
    const bool v2 = use_variant2(head_dim, q, k, v, d_out) && use_variant2(head_dim, dq, dk, dv, nullptr);
    auto kern = v2 ? window_attn2_kernel<true> : window_attn_kernel<true>;
    kern<<<grid, WARPS * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        (const bf16*)q, (const bf16*)k, (const bf16*)v, drop_mask, (const bf16*)d_out, nullptr, (bf16*)dq, (bf16*)dk, (bf16*)dv, pairs, nq,
        nk, heads, head_dim, scale);
    TA_LAUNCH_CHECK();
    return 0;
}
