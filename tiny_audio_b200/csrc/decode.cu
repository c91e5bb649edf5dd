// KV-cache decode path of the Qwen3 decoder (reference: ASRModel.generate,
// tiny_audio/asr_modeling.py:562-646 -> language_model.generate with use_cache, greedy: asr_config.py:103-111).
//
// One new token per sequence and step => every linear is a "skinny" product  out[M <= 32, N] = X[M, K] . W[N, K]^T  whose
// cost is reading W once: HBM-bound, 2 bytes per weight.  tcgen05 tiles (128 x N x 16 per instruction) would idle 3/4 of the
// array and, worse, leave most SMs without a tile (N / 256 CTAs), so these kernels use warp-level mma.sync with the WEIGHTS
// as the 16-row A operand and the tokens as 8-wide B tiles, 16-byte coalesced weight loads (the k index inside each 32-wide
// chunk is permuted identically for both operands, which a dot product does not see), N / 16 CTAs so the whole GPU streams,
// and an in-CTA split-K over the 8 warps reduced in a fixed order through shared memory (bit-reproducible).
// Rounding points mirror the training path (gemm_sm100.cu epilogues): linear outputs are rounded to bf16 before residual
// adds / SwiGLU, h = bf16(silu(g)) * u.
#include "common.cuh"
#include "kernels.cuh"
#include "tinyaudio_b200.h"

namespace {

enum { SK_BF16 = TA_SKINNY_BF16, SK_F32_RESID = TA_SKINNY_F32_RESID, SK_SWIGLU = TA_SKINNY_SWIGLU };

__device__ __forceinline__ void ld8_bf16(const bf16* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}

__device__ __forceinline__ uint4 ld_stream16(const void* p) {   // weights: read once, keep them out of L1
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                               uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

constexpr int SK_WARPS = 8, SK_ROWS = 16, SK_UNROLL = 4;

// TT = number of 8-token tiles (M <= 8 * TT).  grid.x = N / 16 (SK_SWIGLU: N = 2F weight rows -> F / 8 CTAs).
template <int TT, int EPI>
__global__ void __launch_bounds__(SK_WARPS * 32)
skinny_gemm_kernel(const bf16* __restrict__ X, long long ldx, const bf16* __restrict__ W, long long ldw, int M, int K,
                   void* __restrict__ out, long long ldo, const float* __restrict__ resid) {
    __shared__ float red[SK_WARPS][SK_ROWS][TT * 8 + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    // weight rows of this CTA's A tile: tile row i in [0, 16)
    long long row_lo, row_hi;      // tile rows g and g + 8
    if (EPI == SK_SWIGLU) {
        const int f0 = blockIdx.x * 8;                                   // h features f0 .. f0 + 7
        row_lo = (long long)(f0 / 64) * 128 + (f0 % 64) + g;             // gate row of feature f0 + g
        row_hi = row_lo + 64;                                            // its up row
    } else {
        row_lo = (long long)blockIdx.x * SK_ROWS + g;
        row_hi = row_lo + 8;
    }
    const bf16* w_lo = W + row_lo * ldw + t * 8;
    const bf16* w_hi = W + row_hi * ldw + t * 8;
    const bf16* x_row[TT];
#pragma unroll
    for (int j = 0; j < TT; ++j) x_row[j] = X + (long long)min(8 * j + g, M - 1) * ldx + t * 8;

    float acc[TT][4];
#pragma unroll
    for (int j = 0; j < TT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;

    const int nchunks = K >> 5;
    for (int c0 = warp; c0 < nchunks; c0 += SK_WARPS * SK_UNROLL) {
        uint4 wa[SK_UNROLL], wb[SK_UNROLL];
#pragma unroll
        for (int u = 0; u < SK_UNROLL; ++u) {
            const int c = c0 + u * SK_WARPS;
            if (c < nchunks) {
                wa[u] = ld_stream16(w_lo + c * 32);
                wb[u] = ld_stream16(w_hi + c * 32);
            }
        }
#pragma unroll
        for (int u = 0; u < SK_UNROLL; ++u) {
            const int c = c0 + u * SK_WARPS;
            if (c < nchunks) {
#pragma unroll
                for (int j = 0; j < TT; ++j) {
                    const uint4 xb = __ldg(reinterpret_cast<const uint4*>(x_row[j] + c * 32));
                    mma_bf16_16816(acc[j], wa[u].x, wb[u].x, wa[u].y, wb[u].y, xb.x, xb.y);
                    mma_bf16_16816(acc[j], wa[u].z, wb[u].z, wa[u].w, wb[u].w, xb.z, xb.w);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < TT; ++j) {
        red[warp][g][8 * j + 2 * t] = acc[j][0];
        red[warp][g][8 * j + 2 * t + 1] = acc[j][1];
        red[warp][g + 8][8 * j + 2 * t] = acc[j][2];
        red[warp][g + 8][8 * j + 2 * t + 1] = acc[j][3];
    }
    __syncthreads();
    auto total = [&](int i, int m) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < SK_WARPS; ++w) s += red[w][i][m];     // fixed order: bit-reproducible
        return s;
    };
    if (EPI == SK_SWIGLU) {
        const int f0 = blockIdx.x * 8;
        for (int idx = threadIdx.x; idx < 8 * TT * 8; idx += blockDim.x) {
            const int m = idx >> 3, i = idx & 7;
            if (m < M) {
                const float gt = bf16_round(total(i, m)), up = bf16_round(total(i + 8, m));
                reinterpret_cast<bf16*>(out)[(long long)m * ldo + f0 + i] = __float2bfloat16_rn(bf16_round(gt * sigmoidf_(gt)) * up);
            }
        }
    } else {
        const long long n0 = (long long)blockIdx.x * SK_ROWS;
        for (int idx = threadIdx.x; idx < SK_ROWS * TT * 8; idx += blockDim.x) {
            const int m = idx >> 4, i = idx & 15;
            if (m < M) {
                const float s = total(i, m);
                if (EPI == SK_BF16) reinterpret_cast<bf16*>(out)[(long long)m * ldo + n0 + i] = __float2bfloat16_rn(s);
                else reinterpret_cast<float*>(out)[(long long)m * ldo + n0 + i] = resid[(long long)m * ldo + n0 + i] + bf16_round(s);
            }
        }
    }
}

template <int EPI>
int skinny_launch(const bf16* X, long long ldx, const bf16* W, long long ldw, int M, int N, int K, void* out, long long ldo,
                  const float* resid, cudaStream_t st) {
    const int grid = N / SK_ROWS;
    if (M <= 8) skinny_gemm_kernel<1, EPI><<<grid, SK_WARPS * 32, 0, st>>>(X, ldx, W, ldw, M, K, out, ldo, resid);
    else if (M <= 16) skinny_gemm_kernel<2, EPI><<<grid, SK_WARPS * 32, 0, st>>>(X, ldx, W, ldw, M, K, out, ldo, resid);
    else skinny_gemm_kernel<4, EPI><<<grid, SK_WARPS * 32, 0, st>>>(X, ldx, W, ldw, M, K, out, ldo, resid);
    TA_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// per-head RMSNorm + RoPE of the new token's q / k (same arithmetic as lm_qknorm_rope_fwd_kernel) and the cache append:
// one warp per (sequence, head) over the Hq + 2 Hkv heads of the fused qkv row.  pos = *pos_ptr (device: graph-replayable).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void decode_qknorm_rope_cache_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ q_out, bf16* __restrict__ k_cache,
                                                bf16* __restrict__ v_cache, const float* __restrict__ qw, const float* __restrict__ kw,
                                                const float* __restrict__ cosT, const float* __restrict__ sinT,
                                                const int* __restrict__ pos_ptr, int B, int Hq, int Hkv, int max_seq, float eps) {
    const int HD = 128;
    const int heads = Hq + 2 * Hkv;
    const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wid >= B * heads) return;
    const int lane = threadIdx.x & 31;
    const int h = wid % heads, b = wid / heads;
    const int pos = *pos_ptr;
    const bf16* src = qkv + (long long)b * heads * HD + (long long)h * HD;
    const uint32_t ra = *reinterpret_cast<const uint32_t*>(src + 2 * lane);
    const uint32_t rb = *reinterpret_cast<const uint32_t*>(src + 64 + 2 * lane);
    if (h >= Hq + Hkv) {      // value head: plain copy into the cache
        bf16* dst = v_cache + ((long long)b * max_seq + pos) * (Hkv * HD) + (long long)(h - Hq - Hkv) * HD;
        *reinterpret_cast<uint32_t*>(dst + 2 * lane) = ra;
        *reinterpret_cast<uint32_t*>(dst + 64 + 2 * lane) = rb;
        return;
    }
    const float* w = (h < Hq) ? qw : kw;
    const float2 a = unpack_bf16x2(ra), bb = unpack_bf16x2(rb);
    const float ss = warp_sum(a.x * a.x + a.y * a.y + bb.x * bb.x + bb.y * bb.y);
    const float rstd = rsqrtf(ss / HD + eps);
    const float n0 = bf16_round(a.x * rstd) * w[2 * lane], n1 = bf16_round(a.y * rstd) * w[2 * lane + 1];
    const float n2 = bf16_round(bb.x * rstd) * w[64 + 2 * lane], n3 = bf16_round(bb.y * rstd) * w[64 + 2 * lane + 1];
    const float c0 = cosT[pos * 64 + 2 * lane], c1 = cosT[pos * 64 + 2 * lane + 1];
    const float s0 = sinT[pos * 64 + 2 * lane], s1 = sinT[pos * 64 + 2 * lane + 1];
    bf16* dst = (h < Hq) ? q_out + (long long)b * (Hq * HD) + (long long)h * HD
                         : k_cache + ((long long)b * max_seq + pos) * (Hkv * HD) + (long long)(h - Hq) * HD;
    *reinterpret_cast<uint32_t*>(dst + 2 * lane) = pack_bf16x2(n0 * c0 - n2 * s0, n1 * c1 - n3 * s1);
    *reinterpret_cast<uint32_t*>(dst + 64 + 2 * lane) = pack_bf16x2(n2 * c0 + n0 * s0, n3 * c1 + n1 * s1);
}

// prefill: copy the prompt's roped keys (qk buffer) and values (raw qkv buffer) of one layer into the cache rows [0, S)
__global__ void kv_cache_store_kernel(const bf16* __restrict__ k_src, long long k_ld, const bf16* __restrict__ v_src, long long v_ld,
                                      bf16* __restrict__ k_cache, bf16* __restrict__ v_cache, int B, int S, int KD, int max_seq) {
    const int vecs = KD / 8;
    const long long total = (long long)B * S * vecs;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % vecs) * 8;
        const long long row = i / vecs;
        const int s = (int)(row % S);
        const long long b = row / S;
        const long long dst = (b * max_seq + s) * KD + c;
        *reinterpret_cast<uint4*>(k_cache + dst) = *reinterpret_cast<const uint4*>(k_src + row * k_ld + c);
        *reinterpret_cast<uint4*>(v_cache + dst) = *reinterpret_cast<const uint4*>(v_src + row * v_ld + c);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// single-query attention over the cache (HF:integrations/sdpa_attention.py with a 1-token query; GQA: G = Hq / Hkv query
// heads share one K/V head).  grid (Hkv, B); 8 warps x 4 key slots, 8 lanes per key (16 head dims each); online softmax per
// slot, slots merged through shared memory.  Reads each cached K/V row exactly once: HBM-bound, 512 B per cached token and
// kv head.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int DA_WARPS = 8, DA_SLOTS = DA_WARPS * 4;

template <int G>
__global__ void __launch_bounds__(DA_WARPS * 32)
decode_attn_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k_cache, const bf16* __restrict__ v_cache,
                   bf16* __restrict__ out, long long ld_out, const int* __restrict__ pos_ptr, int Hq, int Hkv, int max_seq,
                   float scale_log2) {
    const int HD = 128;
    __shared__ float s_acc[DA_SLOTS][G][HD];
    __shared__ float s_m[DA_SLOTS][G], s_l[DA_SLOTS][G];
    const int kvh = blockIdx.x, b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = warp * 4 + (lane >> 3), l8 = lane & 7;
    const int n_keys = *pos_ptr + 1;
    const int KD = Hkv * HD;

    float qf[G][16];
#pragma unroll
    for (int gq = 0; gq < G; ++gq) {
        const bf16* qp = q + (long long)b * (Hq * HD) + (long long)(kvh * G + gq) * HD + l8 * 16;
        const uint4 u0 = *reinterpret_cast<const uint4*>(qp), u1 = *reinterpret_cast<const uint4*>(qp + 8);
        const uint32_t w[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float2 f = unpack_bf16x2(w[i]);
            qf[gq][2 * i] = f.x * scale_log2;
            qf[gq][2 * i + 1] = f.y * scale_log2;
        }
    }
    float m[G], l[G], acc[G][16];
#pragma unroll
    for (int gq = 0; gq < G; ++gq) {
        m[gq] = -INFINITY;
        l[gq] = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[gq][i] = 0.f;
    }
    const bf16* kb = k_cache + (long long)b * max_seq * KD + (long long)kvh * HD + l8 * 16;
    const bf16* vb = v_cache + (long long)b * max_seq * KD + (long long)kvh * HD + l8 * 16;
    for (int key0 = 0; key0 < n_keys; key0 += DA_SLOTS) {      // uniform trip count: the shuffles below need the full warp
        const int key = key0 + slot;
        const bool live = key < n_keys;
        const int kk = live ? key : n_keys - 1;
        const uint4 k0 = *reinterpret_cast<const uint4*>(kb + (long long)kk * KD), k1 = *reinterpret_cast<const uint4*>(kb + (long long)kk * KD + 8);
        const uint4 v0 = *reinterpret_cast<const uint4*>(vb + (long long)kk * KD), v1 = *reinterpret_cast<const uint4*>(vb + (long long)kk * KD + 8);
        const uint32_t kw[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
        const uint32_t vw[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        float kf[16], vf[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float2 a = unpack_bf16x2(kw[i]), c = unpack_bf16x2(vw[i]);
            kf[2 * i] = a.x; kf[2 * i + 1] = a.y;
            vf[2 * i] = c.x; vf[2 * i + 1] = c.y;
        }
#pragma unroll
        for (int gq = 0; gq < G; ++gq) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) s = fmaf(qf[gq][i], kf[i], s);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            if (live) {
                const float m_new = fmaxf(m[gq], s);
                const float corr = exp2f(m[gq] - m_new), p = exp2f(s - m_new);
                m[gq] = m_new;
                l[gq] = l[gq] * corr + p;
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[gq][i] = fmaf(acc[gq][i], corr, p * vf[i]);
            }
        }
    }
#pragma unroll
    for (int gq = 0; gq < G; ++gq) {
#pragma unroll
        for (int i = 0; i < 16; ++i) s_acc[slot][gq][l8 * 16 + i] = acc[gq][i];
        if (l8 == 0) {
            s_m[slot][gq] = m[gq];
            s_l[slot][gq] = l[gq];
        }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < G * HD; idx += blockDim.x) {
        const int gq = idx / HD, d = idx % HD;
        float mx = -INFINITY;
#pragma unroll 1
        for (int s = 0; s < DA_SLOTS; ++s) mx = fmaxf(mx, s_m[s][gq]);
        float num = 0.f, den = 0.f;
#pragma unroll 1
        for (int s = 0; s < DA_SLOTS; ++s) {
            const float wgt = exp2f(s_m[s][gq] - mx);      // empty slots: exp2(-inf) = 0
            num = fmaf(s_acc[s][gq][d], wgt, num);
            den = fmaf(s_l[s][gq], wgt, den);
        }
        out[(long long)b * ld_out + (long long)(kvh * G + gq) * HD + d] = __float2bfloat16_rn(num / den);
    }
}

// fp32 embedding lookup of the fed token (HF:models/qwen3/modeling_qwen3.py:391 embed_tokens; fp32 under autocast)
__global__ void embed_rows_kernel(const long long* __restrict__ ids, const float* __restrict__ table, float* __restrict__ out, int D,
                                  long long vocab) {
    const int b = blockIdx.x;
    long long id = ids[b];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    for (int c = threadIdx.x * 4; c < D; c += blockDim.x * 4)
        *reinterpret_cast<float4*>(out + (long long)b * D + c) = *reinterpret_cast<const float4*>(table + id * D + c);
}

// greedy pick: argmax over the first V logits of each row (lowest index wins ties, like torch.argmax); block 0 then advances
// the device-side position counter so that the next step (or graph replay) sees pos + 1
__global__ void __launch_bounds__(1024)
argmax_rows_kernel(const bf16* __restrict__ logits, long long ld, int V, long long* __restrict__ next_ids, int* __restrict__ pos_inc) {
    __shared__ float s_v[32];
    __shared__ int s_i[32];
    const bf16* lr = logits + (long long)blockIdx.x * ld;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x * 8; i < V; i += blockDim.x * 8) {
        float v[8];
        ld8_bf16(lr + i, v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (i + j < V && (v[j] > best || (v[j] == best && i + j < bi))) {
                best = v[j];
                bi = i + j;
            }
    }
    auto better = [](float v, int i, float bv, int b2) { return v > bv || (v == bv && i < b2); };
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (better(ov, oi, best, bi)) { best = ov; bi = oi; }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s_v[warp] = best; s_i[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
        best = (lane < (int)(blockDim.x >> 5)) ? s_v[lane] : -INFINITY;
        bi = (lane < (int)(blockDim.x >> 5)) ? s_i[lane] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (better(ov, oi, best, bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) {
            next_ids[blockIdx.x] = bi;
            if (pos_inc && blockIdx.x == 0) *pos_inc += 1;
        }
    }
}

}  // namespace

// ---- internal launchers (engine.cu) ----
int k_skinny_gemm(const bf16* X, long long ldx, const bf16* W, long long ldw, int M, int N, int K, int mode, void* out, long long ldo,
                  const float* resid, cudaStream_t st) {
    TA_REQUIRE(M >= 1 && M <= 32, "skinny GEMM: M = %d not in [1, 32]", M);
    TA_REQUIRE(N % 16 == 0 && K % 32 == 0 && ldx % 8 == 0 && ldw % 8 == 0, "skinny GEMM: N %% 16, K %% 32, ld %% 8 must be 0 (N %d K %d)", N, K);
    TA_REQUIRE(mode != SK_SWIGLU || N % 128 == 0, "skinny SwiGLU: N must be a multiple of 128");
    TA_REQUIRE(mode != SK_F32_RESID || resid, "skinny GEMM: residual missing");
    switch (mode) {
        case SK_BF16: return skinny_launch<SK_BF16>(X, ldx, W, ldw, M, N, K, out, ldo, resid, st);
        case SK_F32_RESID: return skinny_launch<SK_F32_RESID>(X, ldx, W, ldw, M, N, K, out, ldo, resid, st);
        case SK_SWIGLU: return skinny_launch<SK_SWIGLU>(X, ldx, W, ldw, M, N, K, out, ldo, resid, st);
    }
    TA_REQUIRE(false, "skinny GEMM: unknown mode %d", mode);
}

int k_decode_qknorm_rope_cache(const bf16* qkv, bf16* q_out, bf16* k_cache, bf16* v_cache, const float* qw, const float* kw,
                               const float* cosT, const float* sinT, const int* pos, int B, int Hq, int Hkv, int max_seq, float eps,
                               cudaStream_t st) {
    const int warps = B * (Hq + 2 * Hkv);
    decode_qknorm_rope_cache_kernel<<<ceil_div(warps, 8), 256, 0, st>>>(qkv, q_out, k_cache, v_cache, qw, kw, cosT, sinT, pos, B, Hq,
                                                                        Hkv, max_seq, eps);
    TA_LAUNCH_CHECK();
    return 0;
}

int k_kv_cache_store(const bf16* k_src, long long k_ld, const bf16* v_src, long long v_ld, bf16* k_cache, bf16* v_cache, int B, int S,
                     int KD, int max_seq, cudaStream_t st) {
    const long long total = (long long)B * S * (KD / 8);
    const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    kv_cache_store_kernel<<<grid, 256, 0, st>>>(k_src, k_ld, v_src, v_ld, k_cache, v_cache, B, S, KD, max_seq);
    TA_LAUNCH_CHECK();
    return 0;
}

int k_decode_attn(const bf16* q, const bf16* k_cache, const bf16* v_cache, bf16* out, long long ld_out, const int* pos, int B, int Hq,
                  int Hkv, int max_seq, float scale, cudaStream_t st) {
    const float sl2 = scale * 1.4426950408889634f;
    dim3 grid(Hkv, B);
    const int G = Hq / Hkv;
    if (G == 1) decode_attn_kernel<1><<<grid, DA_WARPS * 32, 0, st>>>(q, k_cache, v_cache, out, ld_out, pos, Hq, Hkv, max_seq, sl2);
    else if (G == 2) decode_attn_kernel<2><<<grid, DA_WARPS * 32, 0, st>>>(q, k_cache, v_cache, out, ld_out, pos, Hq, Hkv, max_seq, sl2);
    else TA_REQUIRE(false, "decode attention: Hq / Hkv = %d not supported (1 or 2)", G);
    TA_LAUNCH_CHECK();
    return 0;
}

int k_embed_rows(const long long* ids, const float* table, float* out, int B, int D, long long vocab, cudaStream_t st) {
    embed_rows_kernel<<<B, 256, 0, st>>>(ids, table, out, D, vocab);
    TA_LAUNCH_CHECK();
    return 0;
}

int k_argmax_rows(const bf16* logits, long long ld, int rows, int V, long long* next_ids, int* pos_inc, cudaStream_t st) {
    argmax_rows_kernel<<<rows, 1024, 0, st>>>(logits, ld, V, next_ids, pos_inc);
    TA_LAUNCH_CHECK();
    return 0;
}

// ---- exported for unit parity tests ----
TA_API int ta_skinny_gemm_bf16(const void* X, long long ldx, const void* W, long long ldw, int M, int N, int K, int mode, void* out,
                               long long ldo, const float* resid, void* stream) {
    TA_REQUIRE(X && W && out, "ta_skinny_gemm_bf16: null pointer");
    return k_skinny_gemm((const bf16*)X, ldx, (const bf16*)W, ldw, M, N, K, mode, out, ldo, resid, reinterpret_cast<cudaStream_t>(stream));
}
TA_API int ta_decode_attn(const void* q, const void* k_cache, const void* v_cache, void* out, long long ld_out, const int* pos, int B,
                          int Hq, int Hkv, int max_seq, float scale, void* stream) {
    TA_REQUIRE(q && k_cache && v_cache && out && pos, "ta_decode_attn: null pointer");
    return k_decode_attn((const bf16*)q, (const bf16*)k_cache, (const bf16*)v_cache, (bf16*)out, ld_out, pos, B, Hq, Hkv, max_seq, scale,
                         reinterpret_cast<cudaStream_t>(stream));
}
TA_API int ta_argmax_rows(const void* logits, long long ld, int rows, int V, long long* next_ids, void* stream) {
    TA_REQUIRE(logits && next_ids, "ta_argmax_rows: null pointer");
    return k_argmax_rows((const bf16*)logits, ld, rows, V, next_ids, nullptr, reinterpret_cast<cudaStream_t>(stream));
}
