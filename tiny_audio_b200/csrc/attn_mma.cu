// Flash attention (forward + backward), bf16 in / fp32 softmax, for the two attention shapes of the path:
//   * GLM-ASR encoder: non-causal MHA, 20 heads x 64, S <= 1500, NO mask   (HF:models/glmasr/modeling_glmasr.py:208-221)
//   * Qwen3 decoder  : causal GQA, 16 q / 8 kv heads x 128, fwd + bwd       (HF:models/qwen3/modeling_qwen3.py:273-291)
//
// v1 implementation: warp-level mma.sync.m16n8k16 (bf16) + ldmatrix + cp.async double buffering.  The encoder
// forward has a tcgen05 successor (attn_tc.cu, when enabled); this file stays as its parity reference and as
// the decoder forward/backward.
#include "common.cuh"
#include "kernels.cuh"
#include "tinyaudio_b200.h"

namespace {

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int TILE = 64;      // q rows per CTA step and kv rows per step
constexpr int NTHREADS = 128; // 4 warps x 16 rows

// smem tile [64][W] bf16 with the 16-byte chunk index XOR-swizzled by (row & 7)
template <int W>
__device__ __forceinline__ bf16* tile_ptr(bf16* base, int row, int chunk) {
    return base + row * W + ((chunk ^ (row & 7)) << 3);
}

template <int W>
__device__ __forceinline__ void load_tile_async(bf16* sm, const bf16* g, long long row_stride, int row0, int S, int tid) {
    constexpr int CH = W / 8;
#pragma unroll
    for (int i = tid; i < TILE * CH; i += NTHREADS) {
        const int r = i / CH, c = i % CH;
        const bool ok = (row0 + r) < S;
        const bf16* src = ok ? (g + (long long)(row0 + r) * row_stride + c * 8) : g;
        cp_async16(tile_ptr<W>(sm, r, c), src, ok);
    }
}

// -------------------------------------------------------------------------------------------------------------
// forward
// -------------------------------------------------------------------------------------------------------------
template <int HD, bool CAUSAL>
__global__ void __launch_bounds__(NTHREADS)
attn_fwd_kernel(const bf16* __restrict__ Q, const bf16* __restrict__ K, const bf16* __restrict__ V, bf16* __restrict__ O,
                float* __restrict__ LSE, int S, int Hq, int Hkv, long long q_rs, long long k_rs, long long v_rs,
                long long o_rs, float scale_log2) {
    extern __shared__ __align__(128) uint8_t smem_attn[];
    bf16* sQ = reinterpret_cast<bf16*>(smem_attn);
    bf16* sK = sQ + TILE * HD;          // 2 stages
    bf16* sV = sK + 2 * TILE * HD;      // 2 stages

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int hk = h / (Hq / Hkv);
    const int q0 = qt * TILE;
    const bf16* Qb = Q + (long long)b * S * q_rs + (long long)h * HD;
    const bf16* Kb = K + (long long)b * S * k_rs + (long long)hk * HD;
    const bf16* Vb = V + (long long)b * S * v_rs + (long long)hk * HD;

    const int n_kv_all = (S + TILE - 1) / TILE;
    const int n_kv = CAUSAL ? min(n_kv_all, qt + 1) : n_kv_all;

    load_tile_async<HD>(sQ, Qb, q_rs, q0, S, tid);
    load_tile_async<HD>(sK, Kb, k_rs, 0, S, tid);
    load_tile_async<HD>(sV, Vb, v_rs, 0, S, tid);
    cp_async_commit();

    float o_acc[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) o_acc[i][j] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

    const int mi = lane >> 3, lr = lane & 7;
    for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        if (j + 1 < n_kv) {
            load_tile_async<HD>(sK + (st ^ 1) * TILE * HD, Kb, k_rs, (j + 1) * TILE, S, tid);
            load_tile_async<HD>(sV + (st ^ 1) * TILE * HD, Vb, v_rs, (j + 1) * TILE, S, tid);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const bf16* cK = sK + st * TILE * HD;
        const bf16* cV = sV + st * TILE * HD;

        // ---- S = Q K^T (16 x 64 per warp) ----
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) s[i][jj] = 0.f;
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
            uint32_t a[4];
            ldsm_x4(a, tile_ptr<HD>(sQ, warp * 16 + (mi & 1) * 8 + lr, ks * 2 + (mi >> 1)));
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                uint32_t bb[4];
                ldsm_x4(bb, tile_ptr<HD>(const_cast<bf16*>(cK), np * 16 + (mi >> 1) * 8 + lr, ks * 2 + (mi & 1)));
                mma_bf16(s[2 * np], a, bb[0], bb[1]);
                mma_bf16(s[2 * np + 1], a, bb[2], bb[3]);
            }
        }
        // ---- mask + online softmax (log2 domain) ----
        const int row_a = q0 + warp * 16 + g;
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int col = j * TILE + nt * 8 + 2 * t + (e & 1);
                const int row = row_a + (e >> 1) * 8;
                float v = s[nt][e] * scale_log2;
                if (col >= S || (CAUSAL && col > row)) v = -INFINITY;
                s[nt][e] = v;
                mx[e >> 1] = fmaxf(mx[e >> 1], v);
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        }
        float corr[2], m_new[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            m_new[r] = fmaxf(m_run[r], mx[r]);
            const float m_use = (m_new[r] == -INFINITY) ? 0.f : m_new[r];
            corr[r] = exp2f(m_run[r] - m_use);
            m_run[r] = m_new[r];
            m_new[r] = m_use;
        }
        float ls[2] = {0.f, 0.f};
        uint32_t p[8][2];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float p0 = exp2f(s[nt][0] - m_new[0]), p1 = exp2f(s[nt][1] - m_new[0]);
            const float p2 = exp2f(s[nt][2] - m_new[1]), p3 = exp2f(s[nt][3] - m_new[1]);
            ls[0] += p0 + p1;
            ls[1] += p2 + p3;
            p[nt][0] = pack_bf16x2(p0, p1);
            p[nt][1] = pack_bf16x2(p2, p3);
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + ls[r];
#pragma unroll
        for (int i = 0; i < HD / 8; ++i) {
            o_acc[i][0] *= corr[0]; o_acc[i][1] *= corr[0];
            o_acc[i][2] *= corr[1]; o_acc[i][3] *= corr[1];
        }
        // ---- O += P V ----
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {   // 16 kv rows per step
            uint32_t a[4] = {p[2 * kk][0], p[2 * kk][1], p[2 * kk + 1][0], p[2 * kk + 1][1]};
#pragma unroll
            for (int dp = 0; dp < HD / 16; ++dp) {
                uint32_t bb[4];
                ldsm_x4_t(bb, tile_ptr<HD>(const_cast<bf16*>(cV), kk * 16 + (mi & 1) * 8 + lr, dp * 2 + (mi >> 1)));
                mma_bf16(o_acc[2 * dp], a, bb[0], bb[1]);
                mma_bf16(o_acc[2 * dp + 1], a, bb[2], bb[3]);
            }
        }
        __syncthreads();
    }

    // ---- finalise ----
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int row = q0 + warp * 16 + g + r * 8;
        if (row < S) {
            const float inv = 1.0f / l_run[r];
            bf16* orow = O + ((long long)b * S + row) * o_rs + (long long)h * HD;
#pragma unroll
            for (int i = 0; i < HD / 8; ++i) {
                const uint32_t u = pack_bf16x2(o_acc[i][2 * r] * inv, o_acc[i][2 * r + 1] * inv);
                *reinterpret_cast<uint32_t*>(orow + i * 8 + 2 * t) = u;
            }
            if (LSE && t == 0) LSE[((long long)b * Hq + h) * S + row] = (m_run[r] + log2f(l_run[r])) * 0.69314718055994531f;
        }
    }
}

// -------------------------------------------------------------------------------------------------------------
// backward: D = rowsum(dO * O)
// -------------------------------------------------------------------------------------------------------------
template <int HD>
__global__ void attn_bwd_prep_kernel(const bf16* __restrict__ O, const bf16* __restrict__ dO, float* __restrict__ D,
                                     int B, int S, int Hq, long long o_rs, long long do_rs) {
    TA_PDL_ENTRY();
    // HD / 8 lanes per (token, head), 16-byte loads: a warp reads 512 contiguous bytes of O and of dO
    constexpr int LPH = HD / 8;
    const long long item = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LPH;
    const int sub = threadIdx.x % LPH;
    const long long total = (long long)B * S * Hq;
    const bool live = item < total;
    const long long it = live ? item : 0;
    const int h = (int)(it % Hq);
    const long long bs = it / Hq;
    const int s = (int)(bs % S), b = (int)(bs / S);
    const uint4 uo = *reinterpret_cast<const uint4*>(O + bs * o_rs + (long long)h * HD + 8 * sub);
    const uint4 ud = *reinterpret_cast<const uint4*>(dO + bs * do_rs + (long long)h * HD + 8 * sub);
    const uint32_t wo[4] = {uo.x, uo.y, uo.z, uo.w}, wd[4] = {ud.x, ud.y, ud.z, ud.w};
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 a = unpack_bf16x2(wo[i]), c = unpack_bf16x2(wd[i]);
        acc += a.x * c.x + a.y * c.y;
    }
#pragma unroll
    for (int o = LPH / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (live && sub == 0) D[((long long)b * Hq + h) * S + s] = acc;
}

// -------------------------------------------------------------------------------------------------------------
// backward main: one CTA per (kv tile, kv head, batch); loops over the q heads of the group and the q tiles
//   dQ accumulated with fp32 atomics into dQacc [B,S,Hq,HD] (pre-zeroed); dK, dV written as bf16
// -------------------------------------------------------------------------------------------------------------
template <int HD, bool CAUSAL>
__global__ void __launch_bounds__(NTHREADS, 1)
attn_bwd_kernel(const bf16* __restrict__ Q, const bf16* __restrict__ K, const bf16* __restrict__ V,
                const bf16* __restrict__ dO, const float* __restrict__ LSE, const float* __restrict__ Dsum,
                float* __restrict__ dQacc, bf16* __restrict__ dK, bf16* __restrict__ dV, int S, int Hq, int Hkv,
                long long q_rs, long long k_rs, long long v_rs, long long do_rs, long long dq_rs, long long dk_rs,
                long long dv_rs, float scale, float scale_log2) {
    extern __shared__ __align__(128) uint8_t smem_attn[];
    bf16* sK = reinterpret_cast<bf16*>(smem_attn);
    bf16* sV = sK + TILE * HD;
    bf16* sQ = sV + TILE * HD;
    bf16* sdO = sQ + TILE * HD;
    bf16* sdS = sdO + TILE * HD;                       // [64 kv][64 q]
    float* sLse = reinterpret_cast<float*>(sdS + TILE * TILE);
    float* sD = sLse + TILE;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3, mi = lane >> 3, lr = lane & 7;
    const int kt = blockIdx.x, hk = blockIdx.y, b = blockIdx.z;
    const int G = Hq / Hkv;
    const int kv0 = kt * TILE;
    const bf16* Kb = K + (long long)b * S * k_rs + (long long)hk * HD;
    const bf16* Vb = V + (long long)b * S * v_rs + (long long)hk * HD;

    load_tile_async<HD>(sK, Kb, k_rs, kv0, S, tid);
    load_tile_async<HD>(sV, Vb, v_rs, kv0, S, tid);
    cp_async_commit();

    float dk_acc[HD / 8][4], dv_acc[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { dk_acc[i][j] = 0.f; dv_acc[i][j] = 0.f; }

    const int n_q = (S + TILE - 1) / TILE;
    const int qt_begin = CAUSAL ? kt : 0;
    const float LOG2E = 1.4426950408889634f;

    for (int hg = 0; hg < G; ++hg) {
        const int h = hk * G + hg;
        const bf16* Qb = Q + (long long)b * S * q_rs + (long long)h * HD;
        const bf16* dOb = dO + (long long)b * S * do_rs + (long long)h * HD;
        const float* lse_b = LSE + ((long long)b * Hq + h) * S;
        const float* d_b = Dsum + ((long long)b * Hq + h) * S;
        for (int qt = qt_begin; qt < n_q; ++qt) {
            const int q0 = qt * TILE;
            __syncthreads();   // previous iteration finished reading sQ / sdO / sdS / sLse
            load_tile_async<HD>(sQ, Qb, q_rs, q0, S, tid);
            load_tile_async<HD>(sdO, dOb, do_rs, q0, S, tid);
            cp_async_commit();
            if (tid < TILE) {
                const int r = q0 + tid;
                sLse[tid] = (r < S) ? lse_b[r] * LOG2E : 0.f;
                sD[tid] = (r < S) ? d_b[r] : 0.f;
            }
            cp_async_wait<0>();
            __syncthreads();

            // ---- S^T = K Q^T  (16 kv x 64 q per warp) ----
            float st[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) st[i][j] = 0.f;
#pragma unroll
            for (int ks = 0; ks < HD / 16; ++ks) {
                uint32_t a[4];
                ldsm_x4(a, tile_ptr<HD>(sK, warp * 16 + (mi & 1) * 8 + lr, ks * 2 + (mi >> 1)));
#pragma unroll
                for (int np = 0; np < 4; ++np) {
                    uint32_t bb[4];
                    ldsm_x4(bb, tile_ptr<HD>(sQ, np * 16 + (mi >> 1) * 8 + lr, ks * 2 + (mi & 1)));
                    mma_bf16(st[2 * np], a, bb[0], bb[1]);
                    mma_bf16(st[2 * np + 1], a, bb[2], bb[3]);
                }
            }
            // ---- P^T = exp2(S^T * scale_log2 - lse) with masks ----
            const int kv_a = kv0 + warp * 16 + g;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int qc = nt * 8 + 2 * t + (e & 1);
                    const int qrow = q0 + qc;
                    const int kv = kv_a + (e >> 1) * 8;
                    const bool ok = (kv < S) && (qrow < S) && (!CAUSAL || kv <= qrow);
                    st[nt][e] = ok ? exp2f(st[nt][e] * scale_log2 - sLse[qc]) : 0.f;
                }
            }
            // ---- dV += P^T dO ----
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {   // 16 q rows (contraction) per step
                uint32_t a[4] = {pack_bf16x2(st[2 * kk][0], st[2 * kk][1]), pack_bf16x2(st[2 * kk][2], st[2 * kk][3]),
                                 pack_bf16x2(st[2 * kk + 1][0], st[2 * kk + 1][1]),
                                 pack_bf16x2(st[2 * kk + 1][2], st[2 * kk + 1][3])};
#pragma unroll
                for (int dp = 0; dp < HD / 16; ++dp) {
                    uint32_t bb[4];
                    ldsm_x4_t(bb, tile_ptr<HD>(sdO, kk * 16 + (mi & 1) * 8 + lr, dp * 2 + (mi >> 1)));
                    mma_bf16(dv_acc[2 * dp], a, bb[0], bb[1]);
                    mma_bf16(dv_acc[2 * dp + 1], a, bb[2], bb[3]);
                }
            }
            // ---- dP^T = V dO^T ----
            float dp_[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dp_[i][j] = 0.f;
#pragma unroll
            for (int ks = 0; ks < HD / 16; ++ks) {
                uint32_t a[4];
                ldsm_x4(a, tile_ptr<HD>(sV, warp * 16 + (mi & 1) * 8 + lr, ks * 2 + (mi >> 1)));
#pragma unroll
                for (int np = 0; np < 4; ++np) {
                    uint32_t bb[4];
                    ldsm_x4(bb, tile_ptr<HD>(sdO, np * 16 + (mi >> 1) * 8 + lr, ks * 2 + (mi & 1)));
                    mma_bf16(dp_[2 * np], a, bb[0], bb[1]);
                    mma_bf16(dp_[2 * np + 1], a, bb[2], bb[3]);
                }
            }
            // ---- dS^T = P^T * (dP^T - D) ; keep as bf16 A-fragments and park a copy in smem for dQ ----
            uint32_t ds[8][2];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const int qc = nt * 8 + 2 * t;
                const float d0 = sD[qc], d1 = sD[qc + 1];
                const float v0 = st[nt][0] * (dp_[nt][0] - d0), v1 = st[nt][1] * (dp_[nt][1] - d1);
                const float v2 = st[nt][2] * (dp_[nt][2] - d0), v3 = st[nt][3] * (dp_[nt][3] - d1);
                ds[nt][0] = pack_bf16x2(v0, v1);
                ds[nt][1] = pack_bf16x2(v2, v3);
                const int r0 = warp * 16 + g;
                *reinterpret_cast<uint32_t*>(tile_ptr<TILE>(sdS, r0, nt) + 2 * t) = ds[nt][0];
                *reinterpret_cast<uint32_t*>(tile_ptr<TILE>(sdS, r0 + 8, nt) + 2 * t) = ds[nt][1];
            }
            // ---- dK += dS^T Q ----
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                uint32_t a[4] = {ds[2 * kk][0], ds[2 * kk][1], ds[2 * kk + 1][0], ds[2 * kk + 1][1]};
#pragma unroll
                for (int dp = 0; dp < HD / 16; ++dp) {
                    uint32_t bb[4];
                    ldsm_x4_t(bb, tile_ptr<HD>(sQ, kk * 16 + (mi & 1) * 8 + lr, dp * 2 + (mi >> 1)));
                    mma_bf16(dk_acc[2 * dp], a, bb[0], bb[1]);
                    mma_bf16(dk_acc[2 * dp + 1], a, bb[2], bb[3]);
                }
            }
            __syncthreads();   // sdS complete
            // ---- dQ (16 q rows per warp) = dS K, in two halves of the head dim ----
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                float dq[HD / 16][4];
#pragma unroll
                for (int i = 0; i < HD / 16; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dq[i][j] = 0.f;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {   // 16 kv (contraction) per step
                    uint32_t a[4];
                    // A[q][kv] read transposed from sdS[kv][q]
                    ldsm_x4_t(a, tile_ptr<TILE>(sdS, kk * 16 + (mi >> 1) * 8 + lr, warp * 2 + (mi & 1)));
#pragma unroll
                    for (int dp = 0; dp < HD / 32; ++dp) {
                        uint32_t bb[4];
                        ldsm_x4_t(bb, tile_ptr<HD>(sK, kk * 16 + (mi & 1) * 8 + lr, half * (HD / 16) + dp * 2 + (mi >> 1)));
                        mma_bf16(dq[2 * dp], a, bb[0], bb[1]);
                        mma_bf16(dq[2 * dp + 1], a, bb[2], bb[3]);
                    }
                }
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int row = q0 + warp * 16 + g + r * 8;
                    if (row < S) {
                        float* dst = dQacc + ((long long)b * S + row) * dq_rs + (long long)h * HD + half * (HD / 2) + 2 * t;
#pragma unroll
                        for (int i = 0; i < HD / 16; ++i) {
                            atomicAdd(dst + i * 8, dq[i][2 * r] * scale);
                            atomicAdd(dst + i * 8 + 1, dq[i][2 * r + 1] * scale);
                        }
                    }
                }
            }
        }
    }
    // ---- write dK (scaled), dV ----
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int row = kv0 + warp * 16 + g + r * 8;
        if (row < S) {
            bf16* dk = dK + ((long long)b * S + row) * dk_rs + (long long)hk * HD + 2 * t;
            bf16* dv = dV + ((long long)b * S + row) * dv_rs + (long long)hk * HD + 2 * t;
#pragma unroll
            for (int i = 0; i < HD / 8; ++i) {
                *reinterpret_cast<uint32_t*>(dk + i * 8) = pack_bf16x2(dk_acc[i][2 * r] * scale, dk_acc[i][2 * r + 1] * scale);
                *reinterpret_cast<uint32_t*>(dv + i * 8) = pack_bf16x2(dv_acc[i][2 * r], dv_acc[i][2 * r + 1]);
            }
        }
    }
}

template <int HD, bool CAUSAL>
int launch_fwd(const bf16* q, const bf16* k, const bf16* v, bf16* o, float* lse, int B, int S, int Hq, int Hkv, long long q_rs,
               long long k_rs, long long v_rs, long long o_rs, float scale, cudaStream_t st) {
    const int smem = 5 * TILE * HD * 2;
    auto kern = attn_fwd_kernel<HD, CAUSAL>;
    static bool done = false;
    if (!done) {
        TA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        done = true;
    }
    dim3 grid((S + TILE - 1) / TILE, Hq, B);
    kern<<<grid, NTHREADS, smem, st>>>(q, k, v, o, lse, S, Hq, Hkv, q_rs, k_rs, v_rs, o_rs, scale * 1.4426950408889634f);
    TA_LAUNCH_CHECK();
    return 0;
}

}  // namespace

TA_API int ta_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int S, int Hq, int Hkv,
                       int head_dim, long long q_rs, long long k_rs, long long v_rs, long long o_rs, int causal, float scale,
                       void* stream) {
    TA_REQUIRE(q && k && v && o, "ta_attn_fwd: null pointer");
    TA_REQUIRE(Hq % Hkv == 0, "ta_attn_fwd: Hq must be a multiple of Hkv");
    TA_REQUIRE((q_rs | k_rs | v_rs | o_rs) % 8 == 0, "ta_attn_fwd: row strides must be multiples of 8 elements");
    if (B == 0 || S == 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bf16 *Q = (const bf16*)q, *K = (const bf16*)k, *V = (const bf16*)v;
    bf16* O = (bf16*)o;
    {   // tcgen05 kernel for the encoder shape (attn_tc.cu); other shapes stay on the mma.sync kernels below
        int handled = 0;
        const int rc = k_attn_tc_fwd(Q, K, V, O, lse, B, S, Hq, Hkv, head_dim, q_rs, k_rs, v_rs, o_rs, causal, scale, st, &handled);
        if (rc) return rc;
        if (handled) return 0;
        // no silent second backend: with the tcgen05 path selected (the default) a shape it does not cover is an error; the mma.sync
        // kernels below run only when they are asked for explicitly (ta_attn_set_tc(0): parity / A-B reference)
        TA_REQUIRE(!k_attn_tc_enabled(), "ta_attn_fwd: shape not covered by the tcgen05 kernels (head_dim %d, 16-byte aligned q/k/v/o "
                   "required); ta_attn_set_tc(0) selects the mma.sync reference kernels", head_dim);
    }
    if (head_dim == 64 && !causal) return launch_fwd<64, false>(Q, K, V, O, lse, B, S, Hq, Hkv, q_rs, k_rs, v_rs, o_rs, scale, st);
    if (head_dim == 64 && causal) return launch_fwd<64, true>(Q, K, V, O, lse, B, S, Hq, Hkv, q_rs, k_rs, v_rs, o_rs, scale, st);
    if (head_dim == 128 && causal) return launch_fwd<128, true>(Q, K, V, O, lse, B, S, Hq, Hkv, q_rs, k_rs, v_rs, o_rs, scale, st);
    if (head_dim == 128 && !causal) return launch_fwd<128, false>(Q, K, V, O, lse, B, S, Hq, Hkv, q_rs, k_rs, v_rs, o_rs, scale, st);
    ta_set_error("ta_attn_fwd: unsupported head_dim %d (64 or 128)", head_dim);
    return -1;
}

// dsum_ready bit 0: dsum_ws already holds D = rowsum(dO o O) (the o-projection dgrad GEMM's ROWDOT epilogue wrote it), `o` is not read;
// bit 1: dq_acc is already zero (same epilogue)
int k_attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse, float* dsum_ws, float* dq_acc,
               void* dk, void* dv, int B, int S, int Hq, int Hkv, int head_dim, long long q_rs, long long k_rs, long long v_rs, long long o_rs,
               long long do_rs, long long dq_rs, long long dk_rs, long long dv_rs, int causal, float scale, void* stream, int dsum_ready) {
    TA_REQUIRE(q && k && v && (o || (dsum_ready & 1)) && d_o && lse && dsum_ws && dq_acc && dk && dv, "ta_attn_bwd: null pointer");
    TA_REQUIRE(head_dim == 128, "ta_attn_bwd: only head_dim 128 (Qwen3) is on the path, got %d", head_dim);
    TA_REQUIRE(Hq % Hkv == 0, "ta_attn_bwd: Hq must be a multiple of Hkv");
    if (B == 0 || S == 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    constexpr int HD = 128;
    if (!(dsum_ready & 1)) {
        const long long threads = (long long)B * S * Hq * (HD / 8);
        TA_KERNEL_LAUNCH(attn_bwd_prep_kernel<HD>, (unsigned)((threads + 255) / 256), 256, 0, st, (const bf16*)o, (const bf16*)d_o, dsum_ws, B,
                         S, Hq, o_rs, do_rs);
    }
    if (!(dsum_ready & 2)) TA_CHECK_CUDA(cudaMemsetAsync(dq_acc, 0, sizeof(float) * (size_t)B * S * dq_rs, st));
    if (k_attn_tc_enabled()) {   // tcgen05 kernel (attn_tc_bwd.cu); the mma.sync kernel below stays as A/B reference
        int handled = 0;
        const int rc = k_attn_tc_bwd((const bf16*)q, (const bf16*)k, (const bf16*)v, (const bf16*)d_o, lse, dsum_ws, dq_acc, (bf16*)dk,
                                     (bf16*)dv, B, S, Hq, Hkv, head_dim, q_rs, k_rs, v_rs, do_rs, dq_rs, dk_rs, dv_rs, causal, scale, st,
                                     &handled);
        if (rc) return rc;
        if (handled) return 0;
        TA_REQUIRE(false, "ta_attn_bwd: shape not covered by the tcgen05 kernel (causal, head_dim 128); ta_attn_set_tc(0) selects the "
                   "mma.sync reference kernel");
    }
    const int smem = 4 * TILE * HD * 2 + TILE * TILE * 2 + 2 * TILE * 4;
    dim3 grid((S + TILE - 1) / TILE, Hkv, B);
    const float sl2 = scale * 1.4426950408889634f;
    if (causal) {
        auto kern = attn_bwd_kernel<HD, true>;
        static bool done = false;
        if (!done) { TA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); done = true; }
        kern<<<grid, NTHREADS, smem, st>>>((const bf16*)q, (const bf16*)k, (const bf16*)v, (const bf16*)d_o, lse, dsum_ws, dq_acc,
                                           (bf16*)dk, (bf16*)dv, S, Hq, Hkv, q_rs, k_rs, v_rs, do_rs, dq_rs, dk_rs, dv_rs, scale, sl2);
    } else {
        auto kern = attn_bwd_kernel<HD, false>;
        static bool done = false;
        if (!done) { TA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); done = true; }
        kern<<<grid, NTHREADS, smem, st>>>((const bf16*)q, (const bf16*)k, (const bf16*)v, (const bf16*)d_o, lse, dsum_ws, dq_acc,
                                           (bf16*)dk, (bf16*)dv, S, Hq, Hkv, q_rs, k_rs, v_rs, do_rs, dq_rs, dk_rs, dv_rs, scale, sl2);
    }
    TA_LAUNCH_CHECK();
    return 0;
}

TA_API int ta_attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse,
                       float* dsum_ws, float* dq_acc, void* dk, void* dv, int B, int S, int Hq, int Hkv, int head_dim,
                       long long q_rs, long long k_rs, long long v_rs, long long o_rs, long long do_rs, long long dq_rs,
                       long long dk_rs, long long dv_rs, int causal, float scale, void* stream) {
    return k_attn_bwd(q, k, v, o, d_o, lse, dsum_ws, dq_acc, dk, dv, B, S, Hq, Hkv, head_dim, q_rs, k_rs, v_rs, o_rs, do_rs, dq_rs, dk_rs,
                      dv_rs, causal, scale, stream, 0);
}
