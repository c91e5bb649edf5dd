// tcgen05 / TMEM / TMA GEMM for sm_100a:  C[M,N] = epilogue( A[M,K] . B[N,K]^T )
//
//   A, B : bf16, K-major (row-major with K contiguous), leading dims in elements (multiples of 8)
//   accumulate fp32 in TMEM, persistent CTAs (one per SM), warp-specialised:
//     warp 0 : TMA producer (one elected lane)      smem ring of STAGES x {A 128x64, B BNx64}, SWIZZLE_128B
//     warp 1 : MMA issuer  (one lane)               tcgen05.mma.cta_group::1.kind::f16, 128 x BN x 16
//     warp 2 : TMEM allocator (2 x BN columns: double-buffered accumulator)
//     warps 4-7 : epilogue (tcgen05.ld 32x32b, fused bias / GELU / residual / SwiGLU / SwiGLU-backward)
//
// Every linear layer of the path (GLM-ASR encoder, projector, Qwen3 fwd + dgrad, lm_head) goes through this file.
// Reference call sites replaced: HF:models/glmasr/modeling_glmasr.py:198-206,223,228-239 ; HF:models/qwen3/
// modeling_qwen3.py:81-83,263-291,505 ; tiny_audio/projectors.py:66-71.
#include "common.cuh"
#include "tinyaudio_b200.h"

#include <mutex>
#include <unordered_map>
#include <cstdarg>

// ----------------------------------------------------------------------------------------------
// error string (shared by all translation units)
// ----------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "ok";
void ta_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
TA_API const char* ta_last_error_string(void) { return g_err; }
TA_API int ta_version(void) { return 100; }
unsigned long long g_ta_launches = 0;
TA_API unsigned long long ta_launch_count(void) { return g_ta_launches; }
// programmatic dependent launch for the training towers: only in builds made with `make PDL=1` (common.cuh); returns 1 when the
// request took effect, 0 when this build has no PDL code (the default build)
#ifdef TA_PDL
int g_ta_pdl = 0;
TA_API int ta_set_pdl(int on) {
    g_ta_pdl = on ? 1 : 0;
    return 1;
}
#else
TA_API int ta_set_pdl(int) { return 0; }
#endif

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 384;   // warps 0-3: TMA / MMA / TMEM alloc / spare; warps 4-11: two epilogue groups (column halves)

template <int BN>
struct Cfg {
    static constexpr int STAGES = (BN == 128) ? 6 : 4;
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 2 * BN * 4 /*bias*/;
};

struct EpiArgs {
    void* out;
    long long ldo;
    const float* bias;
    const void* resid;
    long long ldr;
    void* out2;
    long long ldo2;
    const bf16* aux;
    long long ldaux;
    float alpha;
    const float* rope_cos;   // BF16_ROPE: [seq, 16] tables, rotation applied to columns < rope_cols, first 32 dims of each 64-wide head
    const float* rope_sin;
    int rope_seq;
    int rope_cols;
    int k_splits;            // SPLITK instantiations only: the contraction is cut into k_splits ranges whose partial tiles are reduce-added
    int aux_tma;             // SWIGLU_BWD (pair kernel): the (gate, up) stash comes in by TMA and the gradients leave from the same buffer
    int zero_fill;           // BF16_ROWDOT (pair kernel): tmC2 maps an fp32 [M, N] buffer that the epilogue fills with zeros next to its output
};

__device__ __forceinline__ void store_bf16x32(bf16* dst, const float (&v)[32]) {
    uint4* p = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint4 u;
        u.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
        u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
        u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
        u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
        p[i] = u;
    }
}
__device__ __forceinline__ void load_bf16x32(const bf16* src, float (&v)[32]) {
    const uint4* p = reinterpret_cast<const uint4*>(src);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint4 u = p[i];
        float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
        v[8 * i + 0] = a.x; v[8 * i + 1] = a.y; v[8 * i + 2] = b.x; v[8 * i + 3] = b.y;
        v[8 * i + 4] = c.x; v[8 * i + 5] = c.y; v[8 * i + 6] = d.x; v[8 * i + 7] = d.y;
    }
}
__device__ __forceinline__ void store_f32x32(float* dst, const float (&v)[32]) {
    float4* p = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
__device__ __forceinline__ void load_f32x32(const float* src, float (&v)[32]) {
    const float4* p = reinterpret_cast<const float4*>(src);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float4 f = p[i];
        v[4 * i] = f.x; v[4 * i + 1] = f.y; v[4 * i + 2] = f.z; v[4 * i + 3] = f.w;
    }
}

// One accumulator tile (this thread's TMEM lane = one output row, BN columns starting at tile_col0) -> global memory
template <int BN, int EPI>
__device__ __forceinline__ void epilogue_tile(uint32_t taddr, long long row, bool row_ok, long long tile_col0, const EpiArgs& ep,
                                              int grp /* 0 / 1: which half of the tile's columns this warp group owns */,
                                              const float* s_bias /* smem: bias of this tile's BN columns */) {
    if constexpr (EPI == TA_EPI_SWIGLU) {
        // tile columns come in 128-wide groups: [64 gate | 64 up] (weights interleaved by the host)
        const int sb_lo = (BN == 256) ? grp : 0, sb_hi = (BN == 256) ? grp + 1 : 1;
        const int c_lo = (BN == 256) ? 0 : grp, c_hi = (BN == 256) ? 2 : grp + 1;
#pragma unroll 1
        for (int sb = sb_lo; sb < sb_hi; ++sb) {
#pragma unroll 1
            for (int c = c_lo; c < c_hi; ++c) {
                uint32_t rg[32], ru[32];
                tmem_ld_32x32(taddr + sb * 128 + c * 32, rg);
                tmem_ld_32x32(taddr + sb * 128 + 64 + c * 32, ru);
                tmem_ld_wait();
                __syncwarp();
                if (row_ok) {
                    float g[32], u[32], h[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        g[i] = bf16_round(__uint_as_float(rg[i]));
                        u[i] = bf16_round(__uint_as_float(ru[i]));
                        const float s = bf16_round(g[i] * sigmoidf_(g[i]));   // act_fn output is bf16 under autocast
                        h[i] = s * u[i];
                    }
                    const long long col_gu = tile_col0 + sb * 128 + c * 32;
                    const long long col_h = (tile_col0 + sb * 128) / 2 + c * 32;
                    store_bf16x32(reinterpret_cast<bf16*>(ep.out) + row * ep.ldo + col_h, h);
                    if (ep.out2) {
                        bf16* gu = reinterpret_cast<bf16*>(ep.out2) + row * ep.ldo2 + col_gu;
                        store_bf16x32(gu, g);
                        store_bf16x32(gu + 64, u);
                    }
                }
            }
        }
    } else {
        // operands read from global memory (residual / SwiGLU stash) are prefetched one chunk ahead as raw 16-byte vectors
        constexpr int NRAW = (EPI == TA_EPI_F32_RESID || EPI == TA_EPI_SWIGLU_BWD) ? 8 : (EPI == TA_EPI_BF16_RESID) ? 4 : 1;
        uint4 cur[NRAW], nxt[NRAW];
        auto prefetch = [&](int c, uint4 (&dst)[NRAW]) {
            if constexpr (EPI == TA_EPI_BF16_RESID || EPI == TA_EPI_F32_RESID || EPI == TA_EPI_SWIGLU_BWD) {
                if (!row_ok) return;
                const long long col = tile_col0 + c * 32;
                if constexpr (EPI == TA_EPI_BF16_RESID) {
                    const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(ep.resid) + row * ep.ldr + col);
#pragma unroll
                    for (int i = 0; i < 4; ++i) dst[i] = p[i];
                } else if constexpr (EPI == TA_EPI_F32_RESID) {
                    const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(ep.resid) + row * ep.ldr + col);
#pragma unroll
                    for (int i = 0; i < 8; ++i) dst[i] = p[i];
                } else {
                    const uint4* p = reinterpret_cast<const uint4*>(ep.aux + row * ep.ldaux + (col / 64) * 128 + (col % 64));
#pragma unroll
                    for (int i = 0; i < 4; ++i) { dst[i] = p[i]; dst[4 + i] = p[8 + i]; }   // gate chunk, up chunk (+64 elements)
                }
            }
        };
        const int c_lo = grp * (BN / 64), c_hi = (grp + 1) * (BN / 64);
        prefetch(c_lo, cur);
#pragma unroll 1
        for (int c = c_lo; c < c_hi; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(taddr + c * 32, r);
            if (c + 1 < c_hi) prefetch(c + 1, nxt);
            tmem_ld_wait();
            __syncwarp();
            if (row_ok) {
            const long long col = tile_col0 + c * 32;
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
            if constexpr (EPI != TA_EPI_F32 && EPI != TA_EPI_SWIGLU_BWD) {
                if (ep.bias) {   // staged in shared memory once per tile by the epilogue warps
                    const float4* b4 = reinterpret_cast<const float4*>(s_bias + c * 32);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 b = b4[i];
                        v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
                    }
                }
            }
            if constexpr (EPI == TA_EPI_BF16_ROPE) {
                // partial rotary embedding fused into the q|k|v projection: head = 64 columns, rotary dims = its first 32
                // (= exactly this chunk when col % 64 == 0), pairs (i, i+16)   (HF:models/glmasr/modeling_glmasr.py:156-171)
                if (col < ep.rope_cols && (col & 63) == 0) {
                    const int pos = (int)(row % ep.rope_seq);
                    float cs[16], sn[16];
                    const float4* c4 = reinterpret_cast<const float4*>(ep.rope_cos + pos * 16);
                    const float4* s4 = reinterpret_cast<const float4*>(ep.rope_sin + pos * 16);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 a = c4[i], b = s4[i];
                        cs[4 * i] = a.x; cs[4 * i + 1] = a.y; cs[4 * i + 2] = a.z; cs[4 * i + 3] = a.w;
                        sn[4 * i] = b.x; sn[4 * i + 1] = b.y; sn[4 * i + 2] = b.z; sn[4 * i + 3] = b.w;
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float x1 = bf16_round(v[i]), x2 = bf16_round(v[i + 16]);
                        v[i] = x1 * cs[i] - x2 * sn[i];
                        v[i + 16] = x2 * cs[i] + x1 * sn[i];
                    }
                }
                store_bf16x32(reinterpret_cast<bf16*>(ep.out) + row * ep.ldo + col, v);
            } else if constexpr (EPI == TA_EPI_BF16) {
                store_bf16x32(reinterpret_cast<bf16*>(ep.out) + row * ep.ldo + col, v);
            } else if constexpr (EPI == TA_EPI_BF16_GELU) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = gelu_erf_fast(bf16_round(v[i]));
                store_bf16x32(reinterpret_cast<bf16*>(ep.out) + row * ep.ldo + col, v);
            } else if constexpr (EPI == TA_EPI_BF16_RESID) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 a = unpack_bf16x2(cur[i].x), b = unpack_bf16x2(cur[i].y), cc = unpack_bf16x2(cur[i].z),
                                 d = unpack_bf16x2(cur[i].w);
                    v[8 * i + 0] = a.x + bf16_round(v[8 * i + 0]); v[8 * i + 1] = a.y + bf16_round(v[8 * i + 1]);
                    v[8 * i + 2] = b.x + bf16_round(v[8 * i + 2]); v[8 * i + 3] = b.y + bf16_round(v[8 * i + 3]);
                    v[8 * i + 4] = cc.x + bf16_round(v[8 * i + 4]); v[8 * i + 5] = cc.y + bf16_round(v[8 * i + 5]);
                    v[8 * i + 6] = d.x + bf16_round(v[8 * i + 6]); v[8 * i + 7] = d.y + bf16_round(v[8 * i + 7]);
                }
                store_bf16x32(reinterpret_cast<bf16*>(ep.out) + row * ep.ldo + col, v);
            } else if constexpr (EPI == TA_EPI_F32_RESID) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    v[4 * i + 0] = __uint_as_float(cur[i].x) + bf16_round(v[4 * i + 0]);
                    v[4 * i + 1] = __uint_as_float(cur[i].y) + bf16_round(v[4 * i + 1]);
                    v[4 * i + 2] = __uint_as_float(cur[i].z) + bf16_round(v[4 * i + 2]);
                    v[4 * i + 3] = __uint_as_float(cur[i].w) + bf16_round(v[4 * i + 3]);
                }
                store_f32x32(reinterpret_cast<float*>(ep.out) + row * ep.ldo + col, v);
            } else if constexpr (EPI == TA_EPI_F32) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] *= ep.alpha;
                store_f32x32(reinterpret_cast<float*>(ep.out) + row * ep.ldo + col, v);
            } else if constexpr (EPI == TA_EPI_SWIGLU_BWD) {
                // v = d(h) for h columns [col, col+32); the stash holds (gate, up) interleaved in 64-blocks
                const long long jb = col / 64, jo = col % 64;
                float dg[32], du[32];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t gw[4] = {cur[i].x, cur[i].y, cur[i].z, cur[i].w};
                    const uint32_t uw[4] = {cur[4 + i].x, cur[4 + i].y, cur[4 + i].z, cur[4 + i].w};
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const float2 g2 = unpack_bf16x2(gw[t]), u2 = unpack_bf16x2(uw[t]);
                        const float gq[2] = {g2.x, g2.y}, uq[2] = {u2.x, u2.y};
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int k = 8 * i + 2 * t + e;
                            const float dh = bf16_round(v[k]);
                            const float sg = sigmoidf_(gq[e]);
                            const float silu = bf16_round(gq[e] * sg);
                            du[k] = dh * silu;
                            dg[k] = dh * uq[e] * (sg * (1.0f + gq[e] * (1.0f - sg)));
                        }
                    }
                }
                bf16* o = reinterpret_cast<bf16*>(ep.out) + row * ep.ldo + jb * 128 + jo;
                store_bf16x32(o, dg);
                store_bf16x32(o + 64, du);
            }
            }   // row_ok
            if (c + 1 < c_hi) {
#pragma unroll
                for (int i = 0; i < NRAW; ++i) cur[i] = nxt[i];
            }
            __syncwarp();
        }
    }
}

template <int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K,
            EpiArgs ep) {
    using C = Cfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
    uint8_t* smA = smem;
    uint8_t* smB = smem + C::STAGES * C::A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + C::STAGES;
    uint64_t* tfull = bars + 2 * C::STAGES;
    uint64_t* tempty = bars + 2 * C::STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);
    float* s_bias_all = reinterpret_cast<float*>(smem + C::STAGES * C::STAGE_BYTES + 256);   // [2][BN]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int tiles_n = N / BN;
    const int tiles_m = (M + BM - 1) / BM;
    const int num_tiles = tiles_m * tiles_n;
    const int num_kb = (K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull[s], 1);
            mbar_init(&tempty[s], 8);
        }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc<C::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int n_blk = tile % tiles_n, m_blk = tile / tiles_n;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full[stage], C::STAGE_BYTES);
                    tma_load_2d(smA + stage * C::A_BYTES, &tmA, &full[stage], kb * BK, m_blk * BM);
                    tma_load_2d(smB + stage * C::B_BYTES, &tmB, &full[stage], kb * BK, n_blk * BN);
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // whole warp, warp-uniform operands, one elected lane issues (see gemm2_kernel: an `if (lane == 0)` region costs ~100 clk of
        // ELECT / R2UR.BROADCAST plumbing per UTCHMMA)
        {
            constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
            const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
            const bool issuer = elect_one();
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&tempty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tb + (uint32_t)(as * BN);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint64_t ad0 = umma_desc_sw128_kmajor(smem_u32(smA + stage * C::A_BYTES));
                    const uint64_t bd0 = umma_desc_sw128_kmajor(smem_u32(smB + stage * C::B_BYTES));
                    if (issuer) {
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k)      // + 32 bytes (>> 4 = 2) per 16-wide k step
                            umma_f16(d_tmem, ad0 + 2 * k, bd0 + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        umma_commit(&empty[stage]);
                    }
                    __syncwarp();
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                if (issuer) umma_commit(&tfull[as]);
                __syncwarp();
                as ^= 1;
                if (as == 0) aphase ^= 1;
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int q = warp & 3;   // TMEM lane quadrant this warp may touch
        int tile_iter = 0;
        int as = 0;
        uint32_t aphase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int n_blk = tile % tiles_n, m_blk = tile / tiles_n;
            if (ep.bias) {   // stage this tile's bias (BN floats) in shared memory; double buffered across tiles
                float* sb = s_bias_all + (tile_iter & 1) * BN;
                const int e_tid = (warp - 4) * 32 + lane;
                if (e_tid < BN) sb[e_tid] = ep.bias[(long long)n_blk * BN + e_tid];
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            const float* s_bias = s_bias_all + (tile_iter & 1) * BN;
            ++tile_iter;
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
            const long long row = (long long)m_blk * BM + q * 32 + lane;
            const bool row_ok = row < M;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);

            epilogue_tile<BN, EPI>(taddr, row, row_ok, (long long)n_blk * BN, ep, (warp - 4) >> 2, s_bias);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[as]);
            as ^= 1;
            if (as == 0) aphase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}



// ----------------------------------------------------------------------------------------------
// Epilogue through shared memory + TMA store (pair kernel).  Each epilogue warp group (4 warps = the 128 rows of the
// CTA's accumulator) owns two 16 KB staging buffers; a "sub-tile" is [128 rows x 128 bytes] (64 bf16 or 32 fp32
// columns) written with the SWIZZLE_128B pattern the output tensor map expects, then stored by one elected thread.
// This replaces 32-row-strided 16-byte global stores (32 sectors per warp instruction) by one bulk store per sub-tile
// and lets TMA clip the M tail.
// ----------------------------------------------------------------------------------------------
constexpr int STG_BYTES = 128 * 128;

struct Stager {
    uint8_t* buf;       // this group's 2 x STG_BYTES
    uint32_t bar_id;    // named barrier of the group's 128 threads
    int r;              // my row (0..127)
    int n;              // sub-tiles issued so far
    bool leader;
    const uint8_t* zbuf = nullptr;   // ROWDOT: 16 KB of zeros in shared memory (source of the zero-fill stores), or nullptr

    __device__ __forceinline__ uint8_t* begin() {
        if (leader && n >= 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        return buf + (n & 1) * STG_BYTES + r * 128;
    }
    template <bool REDUCE = false>
    __device__ __forceinline__ void end(const CUtensorMap* map, int col0, int row0) {
        fence_proxy_async_smem();
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        if (leader) {
            if constexpr (REDUCE) tma_reduce_add_2d(map, buf + (n & 1) * STG_BYTES, col0, row0);
            else tma_store_2d(map, buf + (n & 1) * STG_BYTES, col0, row0);
            tma_store_commit();
        }
        ++n;
    }
    // bf16 sub-tile (64 columns) as end(), plus -- in the same bulk group -- zeros for the same rows x columns of an fp32 tensor
    __device__ __forceinline__ void end_with_zero(const CUtensorMap* map, const CUtensorMap* zmap, int col0, int row0) {
        fence_proxy_async_smem();
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        if (leader) {
            tma_store_2d(map, buf + (n & 1) * STG_BYTES, col0, row0);
            if (zbuf) {
                tma_store_2d(zmap, zbuf, col0, row0);
                tma_store_2d(zmap, zbuf, col0 + 32, row0);
            }
            tma_store_commit();
        }
        ++n;
    }
    __device__ __forceinline__ void drain() {
        if (leader) tma_store_wait_all();
    }
};

// 32 bf16 values = half (which = 0 / 1) of my 128-byte staging row
__device__ __forceinline__ void stage_bf16x32(uint8_t* row_ptr, int r, int which, const float (&v)[32]) {
#pragma unroll
    for (int qd = 0; qd < 4; ++qd) {
        uint4 u;
        u.x = pack_bf16x2(v[8 * qd + 0], v[8 * qd + 1]);
        u.y = pack_bf16x2(v[8 * qd + 2], v[8 * qd + 3]);
        u.z = pack_bf16x2(v[8 * qd + 4], v[8 * qd + 5]);
        u.w = pack_bf16x2(v[8 * qd + 6], v[8 * qd + 7]);
        *reinterpret_cast<uint4*>(row_ptr + (((which * 4 + qd) ^ (r & 7)) << 4)) = u;
    }
}
// 32 fp32 values = my whole 128-byte staging row
__device__ __forceinline__ void stage_f32x32(uint8_t* row_ptr, int r, const float (&v)[32]) {
#pragma unroll
    for (int qd = 0; qd < 8; ++qd)
        *reinterpret_cast<float4*>(row_ptr + ((qd ^ (r & 7)) << 4)) = make_float4(v[4 * qd], v[4 * qd + 1], v[4 * qd + 2], v[4 * qd + 3]);
}

template <int BN, int EPI, bool REDUCE = false>
__device__ __forceinline__ void epilogue_tile_tma(uint32_t taddr, long long row, bool row_ok, int tile_row0, long long tile_col0,
                                                  const EpiArgs& ep, int grp, const float* s_bias, Stager& sg,
                                                  const CUtensorMap* tmC, const CUtensorMap* tmC2) {
    const int r = sg.r;
    if constexpr (EPI == TA_EPI_SWIGLU) {
        // 128 accumulator columns = [64 gate | 64 up] -> one h sub-tile (64 cols) + optional (gate, up) stash sub-tiles
        const int sb_lo = (BN == 256) ? grp : 0, sb_hi = (BN == 256) ? grp + 1 : 1;
#pragma unroll 1
        for (int sb = sb_lo; sb < sb_hi; ++sb) {
            if (BN == 128 && grp == 1) break;   // 128-wide tiles: one group does the whole (small) tile
            uint8_t* hrow = sg.begin();
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                uint32_t rg[32], ru[32];
                tmem_ld_32x32(taddr + sb * 128 + c * 32, rg);
                tmem_ld_32x32(taddr + sb * 128 + 64 + c * 32, ru);
                tmem_ld_wait();
                float h[32], gg[32], uu[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    gg[i] = __uint_as_float(rg[i]);
                    uu[i] = __uint_as_float(ru[i]);
                }
                bf16_round_all(gg);
                bf16_round_all(uu);
#pragma unroll
                for (int i = 0; i < 32; ++i) h[i] = gg[i] * sigmoidf_(gg[i]);
                bf16_round_all(h);                      // act_fn output is bf16 under autocast
#pragma unroll
                for (int i = 0; i < 32; ++i) h[i] *= uu[i];
                stage_bf16x32(hrow, r, c, h);
            }
            sg.end(tmC, (int)((tile_col0 + sb * 128) / 2), tile_row0);
            if (ep.out2) {
#pragma unroll 1
                for (int part = 0; part < 2; ++part) {      // gate block, up block
                    uint8_t* prow = sg.begin();
#pragma unroll 1
                    for (int c = 0; c < 2; ++c) {
                        uint32_t rr[32];
                        tmem_ld_32x32(taddr + sb * 128 + part * 64 + c * 32, rr);
                        tmem_ld_wait();
                        float v[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(rr[i]);
                        stage_bf16x32(prow, r, c, v);
                    }
                    sg.end(tmC2, (int)(tile_col0 + sb * 128 + part * 64), tile_row0);
                }
            }
        }
    } else {
        constexpr bool F32OUT = (EPI == TA_EPI_F32 || EPI == TA_EPI_F32_RESID);
        constexpr int NRAW = (EPI == TA_EPI_F32_RESID || EPI == TA_EPI_SWIGLU_BWD) ? 8 : (EPI == TA_EPI_BF16_RESID || EPI == TA_EPI_BF16_ROWDOT) ? 4 : 1;
        uint4 cur[NRAW], nxt[NRAW];
        auto prefetch = [&](int c, uint4 (&dst)[NRAW]) {
            if constexpr (EPI == TA_EPI_BF16_RESID || EPI == TA_EPI_F32_RESID || EPI == TA_EPI_SWIGLU_BWD || EPI == TA_EPI_BF16_ROWDOT) {
                if (!row_ok) return;
                const long long col = tile_col0 + c * 32;
                if constexpr (EPI == TA_EPI_BF16_ROWDOT) {
                    const uint4* p = reinterpret_cast<const uint4*>(ep.aux + row * ep.ldaux + col);
#pragma unroll
                    for (int i = 0; i < 4; ++i) dst[i] = p[i];
                } else if constexpr (EPI == TA_EPI_BF16_RESID) {
                    const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(ep.resid) + row * ep.ldr + col);
#pragma unroll
                    for (int i = 0; i < 4; ++i) dst[i] = p[i];
                } else if constexpr (EPI == TA_EPI_F32_RESID) {
                    const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(ep.resid) + row * ep.ldr + col);
#pragma unroll
                    for (int i = 0; i < 8; ++i) dst[i] = p[i];
                } else {
                    const uint4* p = reinterpret_cast<const uint4*>(ep.aux + row * ep.ldaux + (col / 64) * 128 + (col % 64));
#pragma unroll
                    for (int i = 0; i < 4; ++i) { dst[i] = p[i]; dst[4 + i] = p[8 + i]; }
                }
            }
        };
        const int c_lo = grp * (BN / 64), c_hi = (grp + 1) * (BN / 64);
        prefetch(c_lo, cur);
        uint8_t* srow = nullptr;
        float rowdot = 0.f;       // ROWDOT: sum over this group's 128 columns (one head) of out * aux
#pragma unroll 1
        for (int c = c_lo; c < c_hi; ++c) {
            const long long col = tile_col0 + c * 32;
            // sub-tile boundaries: fp32 outputs and SwiGLU-backward flush every chunk, bf16 outputs every second chunk
            const bool first = (EPI != TA_EPI_SWIGLU_BWD) && (F32OUT || ((c - c_lo) & 1) == 0);
            if (first) srow = sg.begin();
            uint32_t rr[32];
            tmem_ld_32x32(taddr + c * 32, rr);
            if (c + 1 < c_hi) prefetch(c + 1, nxt);
            tmem_ld_wait();
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(rr[i]);
            if constexpr (EPI != TA_EPI_F32 && EPI != TA_EPI_SWIGLU_BWD) {
                if (ep.bias) {
                    const float4* b4 = reinterpret_cast<const float4*>(s_bias + c * 32);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 b = b4[i];
                        v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
                    }
                }
            }
            // the linear layer's output is bf16 under autocast: round it ONCE here, in pairs (bf16_round_pair), for every epilogue that
            // continues to compute with it
            if constexpr (EPI == TA_EPI_BF16_ROPE || EPI == TA_EPI_BF16_GELU || EPI == TA_EPI_BF16_RESID || EPI == TA_EPI_F32_RESID ||
                          EPI == TA_EPI_BF16_ROWDOT || EPI == TA_EPI_SWIGLU_BWD)
                bf16_round_all(v);
            if constexpr (EPI == TA_EPI_BF16_ROPE) {
                if (col < ep.rope_cols && (col & 63) == 0) {
                    const int pos = (int)(row % ep.rope_seq);
                    const float4* c4 = reinterpret_cast<const float4*>(ep.rope_cos + pos * 16);
                    const float4* s4 = reinterpret_cast<const float4*>(ep.rope_sin + pos * 16);
                    float cs[16], sn[16];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 a = c4[i], b = s4[i];
                        cs[4 * i] = a.x; cs[4 * i + 1] = a.y; cs[4 * i + 2] = a.z; cs[4 * i + 3] = a.w;
                        sn[4 * i] = b.x; sn[4 * i + 1] = b.y; sn[4 * i + 2] = b.z; sn[4 * i + 3] = b.w;
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float x1 = v[i], x2 = v[i + 16];
                        v[i] = x1 * cs[i] - x2 * sn[i];
                        v[i + 16] = x2 * cs[i] + x1 * sn[i];
                    }
                }
            } else if constexpr (EPI == TA_EPI_BF16_GELU) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = gelu_erf_fast(v[i]);
            } else if constexpr (EPI == TA_EPI_BF16_RESID) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 a = unpack_bf16x2(cur[i].x), b = unpack_bf16x2(cur[i].y), cc = unpack_bf16x2(cur[i].z),
                                 d = unpack_bf16x2(cur[i].w);
                    v[8 * i + 0] = a.x + v[8 * i + 0]; v[8 * i + 1] = a.y + v[8 * i + 1];
                    v[8 * i + 2] = b.x + v[8 * i + 2]; v[8 * i + 3] = b.y + v[8 * i + 3];
                    v[8 * i + 4] = cc.x + v[8 * i + 4]; v[8 * i + 5] = cc.y + v[8 * i + 5];
                    v[8 * i + 6] = d.x + v[8 * i + 6]; v[8 * i + 7] = d.y + v[8 * i + 7];
                }
            } else if constexpr (EPI == TA_EPI_F32_RESID) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    v[4 * i + 0] = __uint_as_float(cur[i].x) + v[4 * i + 0];
                    v[4 * i + 1] = __uint_as_float(cur[i].y) + v[4 * i + 1];
                    v[4 * i + 2] = __uint_as_float(cur[i].z) + v[4 * i + 2];
                    v[4 * i + 3] = __uint_as_float(cur[i].w) + v[4 * i + 3];
                }
            } else if constexpr (EPI == TA_EPI_F32) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] *= ep.alpha;
            } else if constexpr (EPI == TA_EPI_BF16_ROWDOT) {
                if (row_ok) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 a = unpack_bf16x2(cur[i].x), b = unpack_bf16x2(cur[i].y), cc = unpack_bf16x2(cur[i].z),
                                     d = unpack_bf16x2(cur[i].w);
                        rowdot += v[8 * i + 0] * a.x + v[8 * i + 1] * a.y + v[8 * i + 2] * b.x +
                                  v[8 * i + 3] * b.y + v[8 * i + 4] * cc.x + v[8 * i + 5] * cc.y +
                                  v[8 * i + 6] * d.x + v[8 * i + 7] * d.y;
                    }
                }
            }
            if constexpr (EPI == TA_EPI_SWIGLU_BWD) {
                // v = d(h) for h columns [col, col+32) -> (d gate, d up) chunks of the interleaved [M, 2F] gradient
                float dg[32], du[32];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t gw[4] = {cur[i].x, cur[i].y, cur[i].z, cur[i].w};
                    const uint32_t uw[4] = {cur[4 + i].x, cur[4 + i].y, cur[4 + i].z, cur[4 + i].w};
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const float2 g2 = unpack_bf16x2(gw[t]), u2 = unpack_bf16x2(uw[t]);
                        const float gq[2] = {g2.x, g2.y}, uq[2] = {u2.x, u2.y};
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int k = 8 * i + 2 * t + e;
                            const float dh = v[k];
                            const float sgm = sigmoidf_(gq[e]);
                            du[k] = dh * bf16_round(gq[e] * sgm);
                            dg[k] = dh * uq[e] * (sgm * (1.0f + gq[e] * (1.0f - sgm)));
                        }
                    }
                }
                // two consecutive chunks cover h columns [64 j, 64 j + 64): their d(gate) halves fill the sub-tile at output
                // columns [128 j, +64) and their d(up) halves the one at [128 j + 64, +64); both staging buffers are used at once
                const int which = (c - c_lo) & 1;
                uint8_t* rowA = sg.buf + r * 128;
                uint8_t* rowB = sg.buf + STG_BYTES + r * 128;
                if (which == 0) {
                    if (sg.leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    asm volatile("bar.sync %0, 128;" ::"r"(sg.bar_id) : "memory");
                }
                stage_bf16x32(rowA, r, which, dg);
                stage_bf16x32(rowB, r, which, du);
                if (which == 1) {
                    const int jb = (int)((col - 32) / 64);
                    fence_proxy_async_smem();
                    asm volatile("bar.sync %0, 128;" ::"r"(sg.bar_id) : "memory");
                    if (sg.leader) {
                        tma_store_2d(tmC, sg.buf, jb * 128, tile_row0);
                        tma_store_2d(tmC, sg.buf + STG_BYTES, jb * 128 + 64, tile_row0);
                        tma_store_commit();
                    }
                }
            } else if constexpr (F32OUT) {
                stage_f32x32(srow, r, v);
                sg.template end<REDUCE>(tmC, (int)col, tile_row0);
            } else {
                const int which = (c - c_lo) & 1;
                stage_bf16x32(srow, r, which, v);
                if (which == 1) {
                    if constexpr (EPI == TA_EPI_BF16_ROWDOT) sg.end_with_zero(tmC, tmC2, (int)(col - 32), tile_row0);
                    else sg.end(tmC, (int)(col - 32), tile_row0);
                }
            }
            if (c + 1 < c_hi) {
#pragma unroll
                for (int i = 0; i < NRAW; ++i) cur[i] = nxt[i];
            }
        }
        if constexpr (EPI == TA_EPI_BF16_ROWDOT) {
            static_assert(EPI != TA_EPI_BF16_ROWDOT || BN == 256, "ROWDOT: one epilogue group = one 128-wide head needs 256-wide tiles");
            if (row_ok) {
                const long long head = (tile_col0 >> 7) + grp, n_heads = ep.ldo2;            // ldo2 carries N / 128
                const long long bb = row / ep.rope_seq, ss = row % ep.rope_seq;
                reinterpret_cast<float*>(ep.out2)[(bb * n_heads + head) * ep.rope_seq + ss] = rowdot;
            }
        }
    }
}

// ----------------------------------------------------------------------------------------------
// bf16 residual epilogue, in place through shared memory (pair kernel, ep.aux_tma): the agent warp of each epilogue group (warp 2 + grp)
// TMA-loads the residual sub-tile [128 rows x 64 columns] into the group's staging buffer two sub-tiles ahead, every thread adds its
// accumulator row into it, the agent stores the buffer.  No per-thread global loads (whose latency the short K = 1280 o-projection of
// the encoder could not hide: 1 013 TFLOP/s against 1 260 for the other K = 1280 products).
// ----------------------------------------------------------------------------------------------
template <int BN>
__device__ __forceinline__ void epilogue_resid_bf16_inplace(uint32_t taddr, int grp, int r, int lane, uint8_t* gbuf, uint64_t* ldf,
                                                            uint64_t* str, uint32_t& n, const float* s_bias, bool has_bias) {
    constexpr int SPT = BN / 128;              // 64-column sub-tiles per tile and group
    const int sw = r & 7;
#pragma unroll 1
    for (int sub = 0; sub < SPT; ++sub, ++n) {
        const uint32_t set = n & 1u;
        uint8_t* row_ptr = gbuf + set * STG_BYTES + r * 128;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
            const int col_in_tile = (grp * SPT + sub) * 64 + c * 32;
            uint32_t rr[32];
            tmem_ld_32x32(taddr + (uint32_t)col_in_tile, rr);
            if (c == 0) mbar_wait(&ldf[set], (n >> 1) & 1u);
            uint4 x4[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) x4[k] = *reinterpret_cast<const uint4*>(row_ptr + (((c * 4 + k) ^ sw) << 4));
            tmem_ld_wait();
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(rr[i]);
            if (has_bias) {
                const float4* b4 = reinterpret_cast<const float4*>(s_bias + col_in_tile);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 b = b4[i];
                    v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
                }
            }
            bf16_round_all(v);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 a = unpack_bf16x2(x4[k].x), b = unpack_bf16x2(x4[k].y), cc = unpack_bf16x2(x4[k].z), d = unpack_bf16x2(x4[k].w);
                uint4 o;
                o.x = pack_bf16x2(a.x + v[8 * k + 0], a.y + v[8 * k + 1]);
                o.y = pack_bf16x2(b.x + v[8 * k + 2], b.y + v[8 * k + 3]);
                o.z = pack_bf16x2(cc.x + v[8 * k + 4], cc.y + v[8 * k + 5]);
                o.w = pack_bf16x2(d.x + v[8 * k + 6], d.y + v[8 * k + 7]);
                *reinterpret_cast<uint4*>(row_ptr + (((c * 4 + k) ^ sw) << 4)) = o;
            }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&str[set]);
    }
}

// ----------------------------------------------------------------------------------------------
// SwiGLU-backward epilogue, in place through shared memory (pair kernel, ep.aux_tma).
//   A "chunk" is 64 h columns of this CTA's 128 accumulator rows.  Its (gate, up) stash values -- two [128 x 64] bf16 boxes, 16 KB
//   each, SWIZZLE_128B, i.e. whole 128-byte lines of the interleaved [M, 2F] layout -- are brought in by TMA by the agent warp
//   (warp 2), two chunks ahead; BOTH epilogue groups work on the same chunk (group g owns columns [32 g, 32 g + 32) of it = 16-byte
//   pieces 4g..4g+3 of every 128-byte row): each thread reads ITS row of both boxes, overwrites it with (d gate, d up), and the agent
//   stores the two boxes to the gradient tensor, which has the stash's layout.  No global loads in the epilogue threads and no
//   store-drain on their path: they wait on ld_full[set] (TMA landed) and signal st_ready[set] (rows rewritten).
//   History: per-thread 128-bit global loads + staged stores ran at 600 TFLOP/s (tensor pipe 33 % active, profiles/
//   r02_c01_ncu_lm_gemm.txt); 32-column in-place boxes (64-byte rows) reached 840 but wrote half lines, which cost 145 MB of DRAM
//   fill reads per launch (profiles/r02_c19_ncu_swiglu_bwd_tma.txt).
// ----------------------------------------------------------------------------------------------
constexpr int SWB_SET_BYTES = 2 * 16384;    // gate box + up box

template <int BN, int NSET>
__device__ __forceinline__ void epilogue_swiglu_bwd_inplace(uint32_t taddr, int grp, int r, int lane, uint8_t* sbuf, uint64_t* ldf,
                                                            uint64_t* str, uint32_t& n) {
    constexpr int CPT = BN / 64;               // chunks per tile
    const int sw = r & 7;                      // SWIZZLE_128B: 16-byte piece index ^= row & 7
#pragma unroll 1
    for (int i = 0; i < CPT; ++i, ++n) {
        const uint32_t set = n % NSET;
        uint32_t rr[32];
        tmem_ld_32x32(taddr + (uint32_t)(i * 64 + grp * 32), rr);
        mbar_wait(&ldf[set], (n / NSET) & 1u);
        uint8_t* gp = sbuf + set * SWB_SET_BYTES + r * 128;
        uint8_t* up = gp + 16384;
        uint4 g4[4], u4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            g4[k] = *reinterpret_cast<const uint4*>(gp + (((4 * grp + k) ^ sw) << 4));
            u4[k] = *reinterpret_cast<const uint4*>(up + (((4 * grp + k) ^ sw) << 4));
        }
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t gw[4] = {g4[k].x, g4[k].y, g4[k].z, g4[k].w};
            const uint32_t uw[4] = {u4[k].x, u4[k].y, u4[k].z, u4[k].w};
            uint32_t og[4], ou[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float2 g2 = unpack_bf16x2(gw[t]), u2 = unpack_bf16x2(uw[t]);
                const float gq[2] = {g2.x, g2.y}, uq[2] = {u2.x, u2.y};
                float dg[2], du[2];
                float dh[2] = {__uint_as_float(rr[8 * k + 2 * t]), __uint_as_float(rr[8 * k + 2 * t + 1])};
                bf16_round_pair(dh[0], dh[1]);
                const float sgm[2] = {sigmoidf_(gq[0]), sigmoidf_(gq[1])};
                float silu[2] = {gq[0] * sgm[0], gq[1] * sgm[1]};
                bf16_round_pair(silu[0], silu[1]);
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    du[e] = dh[e] * silu[e];
                    dg[e] = dh[e] * uq[e] * (sgm[e] * (1.0f + gq[e] * (1.0f - sgm[e])));
                }
                og[t] = pack_bf16x2(dg[0], dg[1]);
                ou[t] = pack_bf16x2(du[0], du[1]);
            }
            *reinterpret_cast<uint4*>(gp + (((4 * grp + k) ^ sw) << 4)) = make_uint4(og[0], og[1], og[2], og[3]);
            *reinterpret_cast<uint4*>(up + (((4 * grp + k) ^ sw) << 4)) = make_uint4(ou[0], ou[1], ou[2], ou[3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&str[set]);
    }
}

// =============================================================================================================
// 2-CTA variant (cta_group::2): a CTA pair (cluster 2x1x1, two SMs of one TPC) owns a 256 x BN tile.
//   CTA r holds A rows [m0 + 128 r, +128) and the B rows [n0 + r BN/2, + BN/2) of every stage; the leader (rank 0)
//   issues tcgen05.mma.cta_group::2 with M = 256, which reads both CTAs' shared memory and writes each CTA's
//   128 x BN accumulator half into its own TMEM.  Per CTA and k-block this loads 16 KB (A) + BN/2*128 B (B half)
//   instead of 16 KB + BN*128 B: a third less L2->SMEM traffic and shared-memory read bandwidth than the 1-CTA kernel.
//   Barriers:  full[s]   (leader only, count 1)  <- both CTAs' TMA complete_tx + the leader's expect_tx of both halves
//              empty[s]  (each CTA,   count 1)   <- tcgen05.commit ... multicast::cluster (mask 0b11)
//              tfull[a]  (each CTA,   count 1)   <- tcgen05.commit multicast
//              tempty[a] (leader only, count 8)  <- 4 epilogue warps of each CTA (remote arrive from rank 1)
// =============================================================================================================
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address -> rank 0 of the pair

__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const void* tmap, uint64_t* leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(tmap), "r"(smem_u32(leader_bar) & kPeerBitMask), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
// arrive on rank 0's copy of `bar`.  Used for "this warp has read its accumulator quadrant out of TMEM" (tempty): the reads were
// completed by tcgen05.wait::ld and ordered by tcgen05.fence::before_thread_sync, no memory written by this thread has to become
// visible to the MMA issuer, so the arrive is RELAXED -- .release.cluster compiles to MEMBAR.ALL.CTA + ERRBAR, which was 14 % of the
// epilogue warps' stall samples on the SwiGLU-backward GEMM (profiles/r02_c19_ncu_swiglu_bwd_tma.txt)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
#ifdef TA_TEMPTY_RELEASE
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
#else
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
#endif
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}

template <int BN>
struct Cfg2 {
    static constexpr int A_BYTES = BM * BK * 2;            // 128 rows of A per CTA
    static constexpr int B_BYTES = (BN / 2) * BK * 2;      // half of the B tile per CTA
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (BN == 256) ? 4 : 6;
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr int STG_OFF = STAGES * STAGE_BYTES;                   // 2 groups x 2 staging buffers x 16 KB (1024-aligned)
    static constexpr int BAR_OFF = STG_OFF + 4 * STG_BYTES;
    static constexpr int SMEM_BYTES = BAR_OFF + 256 + 2 * BN * 4 + 1024;
    static constexpr int ZERO_OFF = (BAR_OFF + 256 + 2 * BN * 4 + 1023) / 1024 * 1024;   // ROWDOT: 16 KB of zeros behind everything else
    static constexpr int SMEM_BYTES_ZERO = ZERO_OFF + STG_BYTES + 1024;
};

// MN-major SWIZZLE_128B operand tile (TN variant below): rows = K index (128 B = 64 bf16 of the M / N index per row), 8-row groups
// 1024 B apart, further 64-wide M / N atoms `lbo_bytes` apart -- the layout a plain 2-D TMA box {64 cols, K rows} of a row-major
// [K, M] matrix produces (same descriptor form as attn_tc.cu uses for V)
__device__ __forceinline__ uint64_t umma_desc_sw128_mnmajor_g(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// TN = true: C[M,N] = At^T . Bt with At [K, M] and Bt [K, N] row-major (both operands MN-major): the weight-gradient GEMM
// dW = dY^T X reads the activations as they lie in memory (contraction over the token rows), no transposed copies.
// SPLITK (fp32 output only): work item w = (k range w / num_tiles, tile w % num_tiles); every item reduce-adds its partial tile into
// a ZEROED output with TMA (cp.reduce.async.bulk.tensor .add), so few-tile / deep-K products (rank-8 LoRA gradients, d(lm_head),
// projector weight gradients) still fill all CTA pairs.
template <int BN, int EPI, bool TN = false, bool SPLITK = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC2, int M, int N, int K, EpiArgs ep) {
    using C = Cfg2<BN>;
    // SwiGLU-backward in place: NSET box sets next to an NSTAGE-deep operand ring.  Tried 3 sets + a 3-stage ring (same 192 KB): the
    // third set does not pay for the shallower ring -- 117.6 us vs 107.6 us with 2 sets + 4 stages at 14848 x 3072 x 1024 (call 45)
    constexpr int NSTAGE = C::STAGES;
    constexpr int NSET = 2;
    static_assert(EPI != TA_EPI_SWIGLU_BWD || NSTAGE * C::STAGE_BYTES + NSET * SWB_SET_BYTES <= C::BAR_OFF, "in-place sets do not fit");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
    uint8_t* smA = smem;
    uint8_t* smB = smem + NSTAGE * C::A_BYTES;
    uint8_t* smStage = smem + NSTAGE * C::STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
    uint64_t* full = bars;
    uint64_t* empty = bars + C::STAGES;
    uint64_t* tfull = bars + 2 * C::STAGES;
    uint64_t* tempty = bars + 2 * C::STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);
    uint64_t* swb_ld_full = bars + 20;       // [set]: in-place SwiGLU-backward epilogue (ep.aux_tma)
    uint64_t* swb_st_ready = bars + 24;
    static_assert(2 * C::STAGES + 5 <= 20, "barrier area layout");
    float* s_bias_all = reinterpret_cast<float*>(smem + C::BAR_OFF + 256);   // [2][BN]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int tiles_n = N / BN;
    const int tiles_m = (M + 2 * BM - 1) / (2 * BM);
    const int num_tiles = tiles_m * tiles_n;
    const int num_kb = (K + BK - 1) / BK;
    static_assert(!SPLITK || EPI == TA_EPI_F32, "split-K reduces fp32 partial tiles");
    const int kb_per = SPLITK ? (num_kb + ep.k_splits - 1) / ep.k_splits : num_kb;      // k-blocks per work item
    const int num_work = SPLITK ? num_tiles * ep.k_splits : num_tiles;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmC);
        tma_prefetch_desc(&tmC2);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull[s], 1);
            mbar_init(&tempty[s], 16);
        }
        if constexpr (EPI == TA_EPI_BF16_RESID) {      // [group][set]
            for (int s = 0; s < 4; ++s) {
                mbar_init(&swb_ld_full[s], 1);
                mbar_init(&swb_st_ready[s], 4);      // one arrive per warp of the group
            }
        }
        if constexpr (EPI == TA_EPI_SWIGLU_BWD) {
            for (int s = 0; s < NSET; ++s) {
                mbar_init(&swb_ld_full[s], 1);
                mbar_init(&swb_st_ready[s], 8);      // one arrive per epilogue warp of both groups
            }
        }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc_2sm<C::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // both CTAs' barriers are initialised before any remote arrive / multicast commit
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    TA_PDL_ENTRY();              // PDL builds: the prologue above overlaps the predecessor's tail; every global access is below

    if (warp == 0) {
        // ===================== TMA producer (both CTAs load their own halves) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int work = pair; work < num_work; work += n_pairs) {
                const int tile = SPLITK ? work % num_tiles : work;
                const int kb0 = SPLITK ? (work / num_tiles) * kb_per : 0;
                const int kb1 = SPLITK ? min(num_kb, kb0 + kb_per) : num_kb;
                const int n_blk = tile % tiles_n, m_blk = tile / tiles_n;
                const int row0 = m_blk * 2 * BM + (int)rank * BM;
                const int nrow0 = n_blk * BN + (int)rank * (BN / 2);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    if (leader) mbar_arrive_expect_tx(&full[stage], 2 * C::STAGE_BYTES);
                    if constexpr (TN) {      // boxes {64 m (or n), 64 k}: one 8 KB atom per 64 output rows / columns
#pragma unroll
                        for (int i = 0; i < BM / 64; ++i)
                            tma_load_2d_2sm(smA + stage * C::A_BYTES + i * 8192, &tmA, &full[stage], row0 + 64 * i, kb * BK);
#pragma unroll
                        for (int i = 0; i < BN / 128; ++i)
                            tma_load_2d_2sm(smB + stage * C::B_BYTES + i * 8192, &tmB, &full[stage], nrow0 + 64 * i, kb * BK);
                    } else {
                        tma_load_2d_2sm(smA + stage * C::A_BYTES, &tmA, &full[stage], kb * BK, row0);
                        tma_load_2d_2sm(smB + stage * C::B_BYTES, &tmB, &full[stage], kb * BK, nrow0);
                    }
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        // The WHOLE warp walks the loop with warp-uniform operands and one elected lane executes the tcgen05 instructions: inside an
        // `if (lane == 0)` region ptxas cannot keep the descriptors in uniform registers and wraps every UTCHMMA into an
        // ELECT / 5 x R2UR.BROADCAST / BRA.U.ANY loop (~20 instructions, ~100 clk of one thread's latency per MMA, measured on the
        // attention kernels with tools/attn_trace.py) -- next to 128 clk of tensor-pipe time per 256 x 256 x 16 MMA that leaves the
        // issuing thread, not the tensor pipe, as the pacing resource (ncu: tensor pipe 74.5 % active on the largest GEMM).
        if (leader) {
            constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN) | (TN ? ((1u << 15) | (1u << 16)) : 0u);
            const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
            const bool issuer = elect_one();
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            for (int work = pair; work < num_work; work += n_pairs) {
                const int kb0 = SPLITK ? (work / num_tiles) * kb_per : 0;
                const int kb1 = SPLITK ? min(num_kb, kb0 + kb_per) : num_kb;
                mbar_wait(&tempty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tb + (uint32_t)(as * BN);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(smA + stage * C::A_BYTES);
                    const uint32_t b0 = smem_u32(smB + stage * C::B_BYTES);
                    // descriptor of the first 16-wide k step; later steps add their byte offset >> 4 to the start-address field
                    const uint64_t ad0 = TN ? umma_desc_sw128_mnmajor_g(a0, 8192) : umma_desc_sw128_kmajor(a0);
                    const uint64_t bd0 = TN ? umma_desc_sw128_mnmajor_g(b0, 8192) : umma_desc_sw128_kmajor(b0);
                    constexpr uint32_t kstep = (TN ? UMMA_K * 128 : UMMA_K * 2) >> 4;
                    if (issuer) {
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k)
                            umma_f16_2sm(d_tmem, ad0 + (uint64_t)(k * kstep), bd0 + (uint64_t)(k * kstep), idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
                        umma_commit_2sm(&empty[stage]);
                    }
                    __syncwarp();
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
                if (issuer) umma_commit_2sm(&tfull[as]);
                __syncwarp();
                as ^= 1;
                if (as == 0) aphase ^= 1;
            }
        }
    } else if (EPI == TA_EPI_BF16_RESID && (warp == 2 || warp == 3)) {
        // ===================== residual agent of epilogue group (warp - 2) =====================
        if constexpr (EPI == TA_EPI_BF16_RESID) {
            if (ep.aux_tma) {
                constexpr int SPT = BN / 128;
                const int grp = warp - 2;
                uint8_t* gbuf = smStage + grp * 2 * STG_BYTES;
                uint64_t* ldf = swb_ld_full + 2 * grp;
                uint64_t* str = swb_st_ready + 2 * grp;
                const bool issuer = elect_one();
                const int my_tiles = pair < num_work ? (num_work - pair + n_pairs - 1) / n_pairs : 0;
                const uint32_t total = (uint32_t)(my_tiles * SPT);
                auto coords = [&](uint32_t n, int& x, int& y) {
                    const int tile = pair + (int)(n / SPT) * n_pairs;
                    const int n_blk = tile % tiles_n, m_blk = tile / tiles_n;
                    x = n_blk * BN + (grp * SPT + (int)(n % SPT)) * 64;
                    y = m_blk * 2 * BM + (int)rank * BM;
                };
                auto load = [&](uint32_t n) {
                    int x, y;
                    coords(n, x, y);
                    if (issuer) {
                        mbar_arrive_expect_tx(&ldf[n & 1u], STG_BYTES);
                        tma_load_2d(gbuf + (n & 1u) * STG_BYTES, &tmC2, &ldf[n & 1u], x, y);
                    }
                    __syncwarp();
                };
                if (total > 0) load(0);
                if (total > 1) load(1);
                for (uint32_t n = 0; n < total; ++n) {
                    mbar_wait(&str[n & 1u], (n >> 1) & 1u);
                    int x, y;
                    coords(n, x, y);
                    if (issuer) {
                        tma_store_2d(&tmC, gbuf + (n & 1u) * STG_BYTES, x, y);
                        tma_store_commit();
                        if (n + 2 < total) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    }
                    __syncwarp();
                    if (n + 2 < total) load(n + 2);
                }
                if (issuer) tma_store_wait_all();
                __syncwarp();
            }
        }
    } else if (EPI == TA_EPI_SWIGLU_BWD && warp == 2) {
        // ===================== SwiGLU-backward stash agent =====================
        if constexpr (EPI == TA_EPI_SWIGLU_BWD) {
            if (ep.aux_tma) {
                constexpr int CPT = BN / 64;
                uint64_t* ldf = swb_ld_full;
                uint64_t* str = swb_st_ready;
                const bool issuer = elect_one();
                const int my_tiles = pair < num_work ? (num_work - pair + n_pairs - 1) / n_pairs : 0;
                const uint32_t total = (uint32_t)(my_tiles * CPT);
                auto coords = [&](uint32_t n, int& x, int& y) {
                    const int tile = pair + (int)(n / CPT) * n_pairs;
                    const int n_blk = tile % tiles_n, m_blk = tile / tiles_n;
                    const int col = n_blk * BN + (int)(n % CPT) * 64;      // first h column of the chunk
                    x = 2 * col;                                            // its gate columns in the interleaved [M, 2F] layout
                    y = m_blk * 2 * BM + (int)rank * BM;
                };
                auto load = [&](uint32_t n) {
                    int x, y;
                    coords(n, x, y);
                    uint8_t* dst = smStage + (n % NSET) * SWB_SET_BYTES;
                    if (issuer) {
                        mbar_arrive_expect_tx(&ldf[n % NSET], SWB_SET_BYTES);
                        tma_load_2d(dst, &tmC2, &ldf[n % NSET], x, y);
                        tma_load_2d(dst + 16384, &tmC2, &ldf[n % NSET], x + 64, y);
                    }
                    __syncwarp();
                };
                // the stash streams from HBM (it was written a whole forward pass ago): warm L2 two tiles ahead so that the loads below,
                // of which only two can be in flight, see L2 latency
                constexpr uint32_t PF = 2 * CPT;
                auto prefetch = [&](uint32_t n) {
                    if (n >= total) return;
                    int x, y;
                    coords(n, x, y);
                    if (issuer) {
                        tma_prefetch_l2_2d(&tmC2, x, y);
                        tma_prefetch_l2_2d(&tmC2, x + 64, y);
                    }
                };
                for (uint32_t n = 0; n < (uint32_t)NSET && n < total; ++n) load(n);
                for (uint32_t n = NSET; n < PF; ++n) prefetch(n);
                for (uint32_t n = 0; n < total; ++n) {
                    prefetch(n + PF);
                    __syncwarp();
                    mbar_wait(&str[n % NSET], (n / NSET) & 1u);
                    int x, y;
                    coords(n, x, y);
                    const uint8_t* src = smStage + (n % NSET) * SWB_SET_BYTES;
                    if (issuer) {
                        tma_store_2d(&tmC, src, x, y);
                        tma_store_2d(&tmC, src + 16384, x + 64, y);
                        tma_store_commit();
                        if (n + NSET < total) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    }
                    __syncwarp();
                    if (n + NSET < total) load(n + NSET);
                }
                if (issuer) tma_store_wait_all();
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue (both CTAs, their own 128 rows) =====================
        const int q = warp & 3;
        const int grp = (warp - 4) >> 2;
        Stager sg;
        sg.buf = smStage + grp * 2 * STG_BYTES;
        sg.bar_id = 2 + grp;
        sg.r = q * 32 + lane;
        sg.n = 0;
        sg.leader = (q == 0 && lane == 0);
        if constexpr (EPI == TA_EPI_BF16_ROWDOT) {
            if (ep.zero_fill) {          // the launch reserved SMEM_BYTES_ZERO
                uint8_t* z = smem + C::ZERO_OFF;
                const int e_tid = (warp - 4) * 32 + lane;
#pragma unroll
                for (int i = 0; i < 4; ++i) reinterpret_cast<uint4*>(z)[e_tid * 4 + i] = make_uint4(0u, 0u, 0u, 0u);
                fence_proxy_async_smem();
                asm volatile("bar.sync 1, 256;" ::: "memory");
                sg.zbuf = z;
            }
        }
        int tile_iter = 0;
        int as = 0;
        uint32_t aphase = 0;
        uint32_t swb_n = 0;      // chunks this group has processed (in-place SwiGLU-backward)
        for (int work = pair; work < num_work; work += n_pairs) {
            const int tile = SPLITK ? work % num_tiles : work;
            const int n_blk = tile % tiles_n, m_blk = tile / tiles_n;
            if (ep.bias) {   // stage this tile's bias (BN floats) in shared memory; double buffered across tiles
                float* sb = s_bias_all + (tile_iter & 1) * BN;
                const int e_tid = (warp - 4) * 32 + lane;
                if (e_tid < BN) sb[e_tid] = ep.bias[(long long)n_blk * BN + e_tid];
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            const float* s_bias = s_bias_all + (tile_iter & 1) * BN;
            ++tile_iter;
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
            const long long row = (long long)m_blk * 2 * BM + (long long)rank * BM + q * 32 + lane;
            const bool row_ok = row < M;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
            bool done = false;
            if constexpr (EPI == TA_EPI_BF16_RESID) {
                if (ep.aux_tma) {
                    epilogue_resid_bf16_inplace<BN>(taddr, grp, sg.r, lane, sg.buf, swb_ld_full + 2 * grp, swb_st_ready + 2 * grp, swb_n, s_bias,
                                                    ep.bias != nullptr);
                    done = true;
                }
            }
            if constexpr (EPI == TA_EPI_SWIGLU_BWD) {
                if (ep.aux_tma) {
                    epilogue_swiglu_bwd_inplace<BN, NSET>(taddr, grp, sg.r, lane, smStage, swb_ld_full, swb_st_ready, swb_n);
                    done = true;
                }
            }
            if (!done)
                epilogue_tile_tma<BN, EPI, SPLITK>(taddr, row, row_ok, m_blk * 2 * BM + (int)rank * BM, (long long)n_blk * BN, ep, grp,
                                                   s_bias, sg, &tmC, &tmC2);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tempty[as]);
            as ^= 1;
            if (as == 0) aphase ^= 1;
        }
        sg.drain();              // all bulk stores of this group have left shared memory and are complete
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // nobody exits (or frees TMEM) while the peer may still signal / read
    if (warp == 2) tmem_dealloc_2sm<C::TMEM_COLS>(tmem_base);
}

// ----------------------------------------------------------------------------------------------
// host: tensor maps
// ----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

struct MapKey {
    const void* ptr;
    long long rows, cols, ld;
    int box_rows;   // negative: fp32 elements (box = 32 x |box_rows|)
    int box_cols;   // 0: the default 128-byte inner extent
    bool operator==(const MapKey& o) const {
        return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows && box_cols == o.box_cols;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = std::hash<const void*>()(k.ptr);
        h ^= std::hash<long long>()(k.rows * 1315423911LL + k.cols) + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2);
        h ^= std::hash<long long>()(k.ld * 31 + k.box_rows + 7919LL * k.box_cols) + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2);
        return h;
    }
};
std::mutex g_map_mu;
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;

// 2-D bf16 row-major [rows, cols] (ld elements), box = {64 cols, box_rows}, SWIZZLE_128B, zero fill out of bounds.
// box_cols = 32 (bf16): 64-byte inner extent with SWIZZLE_64B (the in-place SwiGLU-backward epilogue's (gate, up) chunks)
int make_map(CUtensorMap* out, const void* ptr, long long rows, long long cols, long long ld, int box_rows, bool f32 = false,
             int box_cols = 0) {
    MapKey key{ptr, rows, cols, ld, f32 ? -box_rows : box_rows, box_cols};
    {
        std::lock_guard<std::mutex> g(g_map_mu);
        auto it = g_maps.find(key);
        if (it != g_maps.end()) {
            *out = it->second;
            return 0;
        }
    }
    EncodeTiledFn fn = get_encode_fn();
    TA_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    TA_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "GEMM operand pointer must be 16-byte aligned");
    const int esz = f32 ? 4 : 2;
    TA_REQUIRE((ld * esz) % 16 == 0, "GEMM operand leading dimension must be a multiple of 16 bytes (got %lld elements)", ld);
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * esz};
    TA_REQUIRE(box_cols == 0 || (!f32 && box_cols == 32), "tensor map: unsupported box width %d", box_cols);
    cuuint32_t box[2] = {(cuuint32_t)(box_cols ? box_cols : (f32 ? 32 : BK)), (cuuint32_t)box_rows};   // 128-byte inner extent by default
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TA_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld ld=%lld)", (int)r, rows,
               cols, ld);
    std::lock_guard<std::mutex> g(g_map_mu);
    if (g_maps.size() > 8192) g_maps.clear();
    g_maps.emplace(key, *out);
    return 0;
}

}  // namespace

// shared with attn_tc.cu (kernels.cuh)
int k_make_tensor_map_2d(CUtensorMap* out, const void* ptr, long long rows, long long cols, long long ld, int box_rows) {
    return make_map(out, ptr, rows, cols, ld, box_rows);
}

namespace {

int g_num_sms = 0;
int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

template <int BN, int EPI>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, const EpiArgs& ep, cudaStream_t st) {
    using C = Cfg<BN>;
    auto kern = gemm_kernel<BN, EPI>;
    static bool attr_done = false;
    if (!attr_done) {
        TA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr_done = true;
    }
    const int tiles = ((M + BM - 1) / BM) * (N / BN);
    const int grid = tiles < num_sms() ? tiles : num_sms();
    kern<<<grid, GEMM_THREADS, C::SMEM_BYTES, st>>>(ta, tb, M, N, K, ep);
    TA_LAUNCH_CHECK();
    return 0;
}

template <int BN>
int dispatch_epi(int epi, const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, const EpiArgs& ep,
                 cudaStream_t st) {
    switch (epi) {
        case TA_EPI_BF16: return launch<BN, TA_EPI_BF16>(ta, tb, M, N, K, ep, st);
        case TA_EPI_BF16_GELU: return launch<BN, TA_EPI_BF16_GELU>(ta, tb, M, N, K, ep, st);
        case TA_EPI_BF16_RESID: return launch<BN, TA_EPI_BF16_RESID>(ta, tb, M, N, K, ep, st);
        case TA_EPI_F32_RESID: return launch<BN, TA_EPI_F32_RESID>(ta, tb, M, N, K, ep, st);
        case TA_EPI_F32: return launch<BN, TA_EPI_F32>(ta, tb, M, N, K, ep, st);
        case TA_EPI_SWIGLU: return launch<BN, TA_EPI_SWIGLU>(ta, tb, M, N, K, ep, st);
        case TA_EPI_SWIGLU_BWD: return launch<BN, TA_EPI_SWIGLU_BWD>(ta, tb, M, N, K, ep, st);
        case TA_EPI_BF16_ROPE: return launch<BN, TA_EPI_BF16_ROPE>(ta, tb, M, N, K, ep, st);
        case TA_EPI_BF16_ROWDOT: ta_set_error("the ROWDOT epilogue exists in the CTA-pair kernel only"); return -1;
        default: ta_set_error("unknown epilogue mode %d", epi); return -1;
    }
}

template <int BN, int EPI, bool TN = false, bool SPLITK = false>
int launch2(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tc2, int M, int N, int K,
            const EpiArgs& ep, cudaStream_t st) {
    using C = Cfg2<BN>;
    auto kern = gemm2_kernel<BN, EPI, TN, SPLITK>;
    constexpr int smem_bytes = (EPI == TA_EPI_BF16_ROWDOT) ? C::SMEM_BYTES_ZERO : C::SMEM_BYTES;
    static bool attr_done = false;
    if (!attr_done) {
        TA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        attr_done = true;
    }
    const int tiles = ((M + 2 * BM - 1) / (2 * BM)) * (N / BN) * (SPLITK ? ep.k_splits : 1);
    int pairs = num_sms() / 2;
    if (tiles < pairs) pairs = tiles;
    TA_KERNEL_LAUNCH(kern, 2 * pairs, GEMM_THREADS, smem_bytes, st, ta, tb, tc, tc2, M, N, K, ep);
    return 0;
}

template <int BN>
int dispatch_epi2(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tc2, int M, int N,
                  int K, const EpiArgs& ep, cudaStream_t st) {
    switch (epi) {
        case TA_EPI_BF16: return launch2<BN, TA_EPI_BF16>(ta, tb, tc, tc2, M, N, K, ep, st);
        case TA_EPI_BF16_GELU: return launch2<BN, TA_EPI_BF16_GELU>(ta, tb, tc, tc2, M, N, K, ep, st);
        case TA_EPI_BF16_RESID: return launch2<BN, TA_EPI_BF16_RESID>(ta, tb, tc, tc2, M, N, K, ep, st);
        case TA_EPI_F32_RESID: return launch2<BN, TA_EPI_F32_RESID>(ta, tb, tc, tc2, M, N, K, ep, st);
        case TA_EPI_F32: return launch2<BN, TA_EPI_F32>(ta, tb, tc, tc2, M, N, K, ep, st);
        case TA_EPI_SWIGLU: return launch2<BN, TA_EPI_SWIGLU>(ta, tb, tc, tc2, M, N, K, ep, st);
        case TA_EPI_SWIGLU_BWD: return launch2<BN, TA_EPI_SWIGLU_BWD>(ta, tb, tc, tc2, M, N, K, ep, st);
        case TA_EPI_BF16_ROPE: return launch2<BN, TA_EPI_BF16_ROPE>(ta, tb, tc, tc2, M, N, K, ep, st);
        case TA_EPI_BF16_ROWDOT:
            if constexpr (BN == 256) return launch2<256, TA_EPI_BF16_ROWDOT>(ta, tb, tc, tc2, M, N, K, ep, st);
            ta_set_error("the ROWDOT epilogue needs 256-wide tiles");
            return -1;
        default: ta_set_error("unknown epilogue mode %d", epi); return -1;
    }
}

int g_force_bn = 0;
int g_tail_split = 0; // 1: split off a poorly filled last wave into 256 x 128 tiles (ta_gemm_set_tail_split).  Off by default: on the
                      // power-capped B200 the step did not get faster (130.4 vs 129.3 ms) -- SMs idling in a short last wave hand their
                      // power budget to the busy ones, which then clock higher, so the quantisation loss is mostly virtual here.
int g_resid_tma = 0;        // 1: in-place TMA epilogue for the bf16 residual GEMMs (ta_gemm_set_resid_tma); 0: per-thread residual loads
int g_swiglu_bwd_tma = 1;   // 1 (default): in-place TMA epilogue for SwiGLU-backward (ta_gemm_set_swiglu_bwd_tma); 0: per-thread loads
int g_cta_pair = 1;   // 1 (default): CTA-pair kernel (cta_group::2, 256 x N tiles); 0: 1-CTA kernel (cta_group::1)

}  // namespace

TA_API int ta_gemm_set_tile_n(int bn) {
    if (bn != 0 && bn != 128 && bn != 256) {
        ta_set_error("tile N must be 0 (auto), 128 or 256");
        return -1;
    }
    g_force_bn = bn;
    return 0;
}

TA_API int ta_gemm_set_tail_split(int on) {
    g_tail_split = on ? 1 : 0;
    return 0;
}

// split-K for the weight-gradient (TN) form: 1 (default) = on for few-tile / deep-K problems, 0 = off (A/B reference).
// Measured on a B200, LoRA recipe at batch 32 x 30 s: 154.5 ms/step unsplit, 142.7 ms split (profiles/r02_c01_*).
int g_tn_splitk = 1;
TA_API int ta_gemm_set_tn_splitk(int on) {
    g_tn_splitk = on ? 1 : 0;
    return 0;
}

TA_API int ta_gemm_set_resid_tma(int on) {
    g_resid_tma = on ? 1 : 0;
    return 0;
}

TA_API int ta_gemm_set_swiglu_bwd_tma(int on) {
    g_swiglu_bwd_tma = on ? 1 : 0;
    return 0;
}

TA_API int ta_gemm_set_cta_pair(int on) {
    g_cta_pair = on ? 1 : 0;
    return 0;
}

// C fp32 [M, N] = alpha * At^T . Bt,  At bf16 [K, M] (ld ldat), Bt bf16 [K, N] (ld ldbt): weight gradients dW = dY^T X
int k_gemm_bf16_tn(const void* At, long long ldat, const void* Bt, long long ldbt, int M, int N, int K, float* out, long long ldo, float alpha,
                   void* stream, int out_zeroed);
TA_API int ta_gemm_bf16_tn(const void* At, long long ldat, const void* Bt, long long ldbt, int M, int N, int K, float* out, long long ldo,
                           float alpha, void* stream) {
    return k_gemm_bf16_tn(At, ldat, Bt, ldbt, M, N, K, out, ldo, alpha, stream, 0);
}

// out_zeroed != 0: the caller has already cleared `out` (the LoRA engine clears all adapters' gradient buffers with a handful of
// memsets per step instead of one per split-K product: 224 of them)
int k_gemm_bf16_tn(const void* At, long long ldat, const void* Bt, long long ldbt, int M, int N, int K, float* out, long long ldo, float alpha,
                   void* stream, int out_zeroed) {
    TA_REQUIRE(At && Bt && out, "ta_gemm_bf16_tn: null pointer");
    TA_REQUIRE(M > 0 && N > 0 && K > 0, "ta_gemm_bf16_tn: empty problem M=%d N=%d K=%d", M, N, K);
    TA_REQUIRE(N % 128 == 0, "ta_gemm_bf16_tn: N=%d must be a multiple of 128", N);
    const int bn = (N % 256 == 0) ? 256 : 128;
    EpiArgs ep;
    memset(&ep, 0, sizeof(ep));
    ep.out = out; ep.ldo = ldo; ep.ldr = ldo; ep.alpha = alpha == 0.0f ? 1.0f : alpha;
    CUtensorMap ta, tb, tc;
    int rc = make_map(&ta, At, K, M, ldat, 64);        // rows = contraction index, box {64 m, 64 k}
    if (rc) return rc;
    rc = make_map(&tb, Bt, K, N, ldbt, 64);
    if (rc) return rc;
    rc = make_map(&tc, out, M, N, ldo, BM, true);
    if (rc) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    ep.k_splits = 1;
    ep.aux_tma = 0;
    ep.zero_fill = 0;
    if (g_tn_splitk) {
        // few output tiles, deep contraction (rank-8 LoRA gradients: 4 ... 24 tiles x 230 k-blocks): cut K so that every CTA pair
        // gets a work item, each at least 8 k-blocks long; partial tiles are reduce-added into the zeroed output
        const int pairs = num_sms() / 2;
        const int tiles = ((M + 2 * BM - 1) / (2 * BM)) * (N / bn);
        const int num_kb = (K + BK - 1) / BK;
        int splits = tiles * 2 <= pairs ? pairs / tiles : 1;
        if (splits > num_kb / 8) splits = num_kb / 8;
        if (splits > 1) {
            const int kb_per = (num_kb + splits - 1) / splits;
            splits = (num_kb + kb_per - 1) / kb_per;          // no empty k range
            ep.k_splits = splits;
            if (!out_zeroed)
                TA_CHECK_CUDA(cudaMemset2DAsync(out, (size_t)ldo * sizeof(float), 0, (size_t)N * sizeof(float), (size_t)M, st));
            if (bn == 256) return launch2<256, TA_EPI_F32, true, true>(ta, tb, tc, tc, M, N, K, ep, st);
            return launch2<128, TA_EPI_F32, true, true>(ta, tb, tc, tc, M, N, K, ep, st);
        }
    }
    if (bn == 256) return launch2<256, TA_EPI_F32, true>(ta, tb, tc, tc, M, N, K, ep, st);
    return launch2<128, TA_EPI_F32, true>(ta, tb, tc, tc, M, N, K, ep, st);
}

TA_API int ta_gemm_bf16(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K, int epi,
                        const ta_gemm_epilogue* e, void* stream) {
    TA_REQUIRE(A && B && e && e->out, "ta_gemm_bf16: null pointer");
    TA_REQUIRE(M > 0 && N > 0 && K > 0, "ta_gemm_bf16: empty problem M=%d N=%d K=%d", M, N, K);
    TA_REQUIRE(N % 128 == 0, "ta_gemm_bf16: N=%d must be a multiple of 128", N);
    if (epi == TA_EPI_BF16_RESID || epi == TA_EPI_F32_RESID) TA_REQUIRE(e->resid, "residual epilogue needs resid");
    if (epi == TA_EPI_SWIGLU_BWD) TA_REQUIRE(e->aux && N % 64 == 0, "swiglu-bwd epilogue needs aux stash");
    if (epi == TA_EPI_BF16_ROWDOT)
        TA_REQUIRE(e->aux && e->out2 && e->rope_seq > 0 && N % 256 == 0 && M % e->rope_seq == 0,
                   "rowdot epilogue needs aux, out2, the sequence length in rope_seq (dividing M) and N %% 256 == 0");
    int bn = g_force_bn ? g_force_bn : ((N % 256 == 0) ? 256 : 128);
    if (N % bn != 0) bn = 128;
    if (epi == TA_EPI_BF16_ROWDOT) bn = 256;
    EpiArgs ep;
    ep.k_splits = 1;
    ep.aux_tma = 0;
    ep.zero_fill = 0;
    ep.out = e->out; ep.ldo = e->ldo; ep.bias = e->bias; ep.resid = e->resid; ep.ldr = e->ldr ? e->ldr : e->ldo;
    ep.out2 = e->out2; ep.ldo2 = e->ldo2; ep.aux = reinterpret_cast<const bf16*>(e->aux); ep.ldaux = e->ldaux;
    ep.alpha = e->alpha == 0.0f ? 1.0f : e->alpha;
    ep.rope_cos = e->rope_cos; ep.rope_sin = e->rope_sin; ep.rope_seq = e->rope_seq; ep.rope_cols = e->rope_cols;
    if (epi == TA_EPI_BF16_ROPE) TA_REQUIRE(e->rope_cos && e->rope_sin && e->rope_seq > 0, "rope epilogue needs cos/sin tables and the sequence length");
    CUtensorMap ta, tb;
    int rc = make_map(&ta, A, M, K, lda, BM);
    if (rc) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    bool use_pair = g_cta_pair != 0 || epi == TA_EPI_BF16_ROWDOT;
    if (use_pair && !g_force_bn && epi != TA_EPI_BF16_ROWDOT) {
        // few, deep tiles (the d(lm_head) product: M = labelled rows, N = dim, K = vocabulary): 256 x 256 pair tiles would leave
        // most SMs idle; 128 x 128 single-CTA tiles give 4 x as many work items (measured 0.60 vs 0.85 ms on 2080 x 1024 x 151936)
        const long long pair_tiles = (long long)((M + 2 * BM - 1) / (2 * BM)) * (N / bn);
        const long long small_tiles = (long long)((M + BM - 1) / BM) * (N / 128);
        if (pair_tiles * 2 <= num_sms() / 2 && small_tiles <= 2LL * num_sms() && epi != TA_EPI_SWIGLU_BWD) {
            use_pair = false;
            bn = 128;
        }
    }
    if (use_pair) {
        const bool f32out = (epi == TA_EPI_F32 || epi == TA_EPI_F32_RESID);
        const long long out_cols = (epi == TA_EPI_SWIGLU) ? N / 2 : (epi == TA_EPI_SWIGLU_BWD) ? 2LL * N : N;
        // rows [row0, row0 + rows) of the problem with tile width bnp (all row-indexed epilogue operands are offset accordingly)
        auto run_rows = [&](long long row0, int rows, int bnp) -> int {
            EpiArgs e2 = ep;
            e2.out = reinterpret_cast<uint8_t*>(ep.out) + row0 * ep.ldo * (f32out ? 4 : 2);
            if (ep.resid) e2.resid = reinterpret_cast<const uint8_t*>(ep.resid) + row0 * ep.ldr * (epi == TA_EPI_F32_RESID ? 4 : 2);
            if (ep.out2 && epi != TA_EPI_BF16_ROWDOT) e2.out2 = reinterpret_cast<uint8_t*>(ep.out2) + row0 * ep.ldo2 * 2;
            if (epi == TA_EPI_BF16_ROWDOT) e2.ldo2 = N / 128;      // heads per row of the D output
            if (ep.aux) e2.aux = ep.aux + row0 * ep.ldaux;
            CUtensorMap ma, mb, tc, tc2;
            int r2 = make_map(&ma, reinterpret_cast<const bf16*>(A) + row0 * lda, rows, K, lda, BM);
            if (r2) return r2;
            r2 = make_map(&mb, B, N, K, ldb, bnp / 2);     // each CTA of the pair loads half of the B tile
            if (r2) return r2;
            // output maps for the TMA-store epilogue: [rows, width] with 128-byte wide sub-tiles
            const bool swb_tma = (epi == TA_EPI_SWIGLU_BWD) && g_swiglu_bwd_tma;
            e2.aux_tma = swb_tma ? 1 : 0;
            r2 = make_map(&tc, e2.out, rows, out_cols, e->ldo, BM, f32out);
            if (r2) return r2;
            tc2 = tc;
            if (swb_tma) {
                r2 = make_map(&tc2, e2.aux, rows, 2LL * N, e->ldaux, BM, false);
                if (r2) return r2;
            }
            if (epi == TA_EPI_BF16_RESID && g_resid_tma && bnp == 256) {      // residual sub-tiles come in by TMA, the sum leaves from the same buffer
                r2 = make_map(&tc2, e2.resid, rows, N, ep.ldr, BM, false);
                if (r2) return r2;
                e2.aux_tma = 1;
            }
            if (epi == TA_EPI_BF16_ROWDOT && e->zero_f32) {      // fp32 [rows, N] buffer to be zero-filled alongside the output
                r2 = make_map(&tc2, reinterpret_cast<uint8_t*>(e->zero_f32) + row0 * e->ld_zero * 4, rows, N, e->ld_zero, BM, true);
                if (r2) return r2;
                e2.zero_fill = 1;
            }
            if (epi == TA_EPI_SWIGLU && e->out2) {
                r2 = make_map(&tc2, e2.out2, rows, N, e->ldo2, BM, false);
                if (r2) return r2;
            }
            if (bnp == 256) return dispatch_epi2<256>(epi, ma, mb, tc, tc2, rows, N, K, e2, st);
            return dispatch_epi2<128>(epi, ma, mb, tc, tc2, rows, N, K, e2, st);
        };
        // Wave quantisation: the persistent kernel runs ceil(tiles / 74 CTA pairs) waves of 256 x 256 tiles.  When the last wave
        // would be mostly empty (the decoder's N = 1024 products: 232 tiles = 3.14 waves -> 4), the trailing row blocks are
        // issued as a second launch with 256 x 128 tiles, which fills the SMs with half-cost tiles (3 + ~0.55 instead of 4).
        if (bn == 256 && !g_force_bn && g_tail_split && epi != TA_EPI_BF16_ROPE && epi != TA_EPI_BF16_ROWDOT) {
            const int pairs = num_sms() / 2;
            const int tm = (M + 2 * BM - 1) / (2 * BM), tn = N / 256;
            const long long tiles = (long long)tm * tn;
            const long long full = tiles / pairs, rem = tiles % pairs;
            if (full >= 1 && rem > 0 && rem * 100 < (long long)pairs * 60) {
                const int m_main = (int)((full * pairs) / tn);          // row blocks that fill `full` waves
                const int rem_m = tm - m_main;
                const long long tail_tiles = (long long)rem_m * (N / 128);
                const double cost_split = (double)((m_main * (long long)tn + pairs - 1) / pairs) + 0.58 * (double)((tail_tiles + pairs - 1) / pairs);
                if (m_main > 0 && rem_m > 0 && cost_split + 0.15 < (double)(full + 1)) {
                    const long long rows_main = (long long)m_main * 2 * BM;
                    rc = run_rows(0, (int)rows_main, 256);
                    if (rc) return rc;
                    return run_rows(rows_main, (int)(M - rows_main), 128);
                }
            }
        }
        return run_rows(0, M, bn);
    }
    rc = make_map(&tb, B, N, K, ldb, bn);
    if (rc) return rc;
    if (bn == 256) return dispatch_epi<256>(epi, ta, tb, M, N, K, ep, st);
    return dispatch_epi<128>(epi, ta, tb, M, N, K, ep, st);
}
