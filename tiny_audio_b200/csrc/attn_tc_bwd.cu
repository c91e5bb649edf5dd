// tcgen05 flash-attention BACKWARD for the Qwen3 decoder shape: head_dim 128, causal, GQA (16 q / 8 kv heads)
// (autograd of HF:models/qwen3/modeling_qwen3.py:273-291 through HF:integrations/sdpa_attention.py).
// Successor of attn_bwd_kernel<128,true> (mma.sync, legacy tensor path).
//
// One CTA = one (batch, kv head, 128-key tile); it keeps K and V in shared memory and dK, dV in TMEM, and loops over
// the q heads of the group and the query tiles at or below the diagonal.  Per iteration, five 128x128x128 UMMAs:
//     S^T  = K  Q^T         (TMEM cols   0..127)           dP^T = V  dO^T        (cols 128..255)
//     -- softmax-backward warps: P^T = exp2(S^T c - lse_q), dS^T = P^T (dP^T - D_q) -> bf16 -> smem --
//     dV  += P^T  dO        (cols 256..383)                dK  += dS^T Q         (cols 384..511)
//     dQ   = dS   K         (reuses cols 0..127; A = dS^T read as an MN-major operand, B = K MN-major)
// dQ tiles are scaled and added to the fp32 dQ buffer with vector reductions (several CTAs contribute to a q tile);
// dK (scaled) and dV are written once as bf16 at the end.  All operand tiles are [128 rows x 64 bf16] SWIZZLE_128B
// sub-tiles; "MN-major" operands are the same tiles consumed through the transposing descriptor form.
#include "common.cuh"
#include "kernels.cuh"
#include "tinyaudio_b200.h"

namespace {

constexpr int HD = 128;
constexpr int BT = 128;                    // q tile = kv tile = 128
constexpr int TILE16 = 128 * 64 * 2;       // one [128 x 64] bf16 sub-tile
constexpr int TILE = 2 * TILE16;           // [128 x 128] bf16
constexpr int BWD_THREADS = 384;
constexpr int BAR_OFF = 6 * TILE;          // sK sV sQ sdO sP sdS
constexpr int SMEM_BWD = BAR_OFF + 256 + 4 * 128 * 4 + 1024;   // + barriers + (lse, D) double buffered + align slack
constexpr uint32_t ST_COL = 0, DP_COL = 128, DV_COL = 256, DK_COL = 384;

__device__ __forceinline__ float ex2_approx_b(float x) {
    float y;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint64_t desc_k(uint32_t addr) { return umma_desc_sw128_kmajor(addr); }
__device__ __forceinline__ uint64_t desc_mn(uint32_t smem_addr) {   // MN-major, 64-wide MN atoms TILE16 apart
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((TILE16 >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                   const float* __restrict__ LSE, const float* __restrict__ Dsum, float* __restrict__ dQacc,
                   bf16* __restrict__ dK, bf16* __restrict__ dV, int S, int Hq, int Hkv, long long dq_rs, long long dk_rs,
                   long long dv_rs, float scale, float scale_log2, int dbg) {
    TA_PDL_ENTRY();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
    uint8_t* sK = smem;
    uint8_t* sV = smem + TILE;
    uint8_t* sQ = smem + 2 * TILE;
    uint8_t* sdO = smem + 3 * TILE;
    uint8_t* sP = smem + 4 * TILE;
    uint8_t* sdS = smem + 5 * TILE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
    uint64_t* kv_full = bars;
    uint64_t* qdo_full = bars + 1;
    uint64_t* qdo_empty = bars + 2;
    uint64_t* s_full = bars + 3;
    uint64_t* pds_full = bars + 4;
    uint64_t* dq_full = bars + 5;
    uint64_t* dq_empty = bars + 6;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    float* sVec = reinterpret_cast<float*>(smem + BAR_OFF + 256);   // [2 buffers][lse*log2e (128) | D (128)]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kt = blockIdx.x, hk = blockIdx.y, b = blockIdx.z;
    const int G = Hq / Hkv;
    const int kv0 = kt * BT;
    const int n_q = (S + BT - 1) / BT;
    const int n_it = G * (n_q - kt);                 // causal: query tiles qt >= kt
    const int row_base = b * S;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
        tma_prefetch_desc(&tmdO);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(kv_full, 1);
        mbar_init(qdo_full, 1);
        mbar_init(qdo_empty, 1);
        mbar_init(s_full, 1);
        mbar_init(pds_full, 8);
        mbar_init(dq_full, 1);
        mbar_init(dq_empty, 8);
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(kv_full, 2 * TILE);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                tma_load_2d(sK + u * TILE16, &tmK, kv_full, hk * HD + u * 64, row_base + kv0);
                tma_load_2d(sV + u * TILE16, &tmV, kv_full, hk * HD + u * 64, row_base + kv0);
            }
            for (int it = 0; it < n_it; ++it) {
                const int g = it / (n_q - kt), qt = kt + it % (n_q - kt);
                const int h = hk * G + g;
                mbar_wait(qdo_empty, ((uint32_t)it & 1u) ^ 1u);
                mbar_arrive_expect_tx(qdo_full, 2 * TILE);
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    tma_load_2d(sQ + u * TILE16, &tmQ, qdo_full, h * HD + u * 64, row_base + qt * BT);
                    tma_load_2d(sdO + u * TILE16, &tmdO, qdo_full, h * HD + u * 64, row_base + qt * BT);
                }
            }
        }
    } else if (warp == 1) {
        {   // whole warp, warp-uniform operands, one elected lane issues: inside an `if (lane == 0)` region every UTCHMMA is wrapped into
            // an ELECT / R2UR.BROADCAST / BRA.U.ANY loop (~100 clk of single-thread latency per MMA, tools/attn_trace.py) -- 40 MMAs per
            // (query tile, key tile) here, against 64 clk of tensor-pipe time each
            constexpr uint32_t id_kk = umma_idesc_bf16(BT, BT);                              // A, B K-major
            constexpr uint32_t id_kmn = umma_idesc_bf16(BT, HD) | (1u << 16);                // B MN-major
            constexpr uint32_t id_mnmn = umma_idesc_bf16(BT, HD) | (1u << 15) | (1u << 16);  // A and B MN-major
            const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
            const bool issuer = elect_one();
            const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aQ = smem_u32(sQ), adO = smem_u32(sdO), aP = smem_u32(sP),
                           adS = smem_u32(sdS);
            mbar_wait(kv_full, 0);
            for (int it = 0; it < n_it; ++it) {
                const uint32_t ph = (uint32_t)it & 1u;
                mbar_wait(qdo_full, ph);
                tc_fence_after();
                // dP^T first: its columns were last read by the previous softmax (finished before pds_full), so it does not have to wait
                // for the previous dQ tile to leave TMEM[0,128) and runs under that read-out
                if (issuer) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint32_t off = (k >> 2) * TILE16 + (k & 3) * 32;
                        umma_f16(tb + DP_COL, desc_k(aV + off), desc_k(adO + off), id_kk, k != 0 ? 1u : 0u);
                    }
                }
                __syncwarp();
                mbar_wait(dq_empty, ph ^ 1u);          // previous dQ tile has been read out of TMEM[0,128)
                tc_fence_after();
                if (issuer) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint32_t off = (k >> 2) * TILE16 + (k & 3) * 32;
                        umma_f16(tb + ST_COL, desc_k(aK + off), desc_k(aQ + off), id_kk, k != 0 ? 1u : 0u);
                    }
                    umma_commit(s_full);
                }
                __syncwarp();
                mbar_wait(pds_full, ph);
                tc_fence_after();
                if (issuer) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {      // contraction over the 128 queries, 16 per step
                        const uint32_t offk = (k >> 2) * TILE16 + (k & 3) * 32;
                        umma_f16(tb + DV_COL, desc_k(aP + offk), desc_mn(adO + k * 2048), id_kmn, (it | k) != 0 ? 1u : 0u);
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint32_t offk = (k >> 2) * TILE16 + (k & 3) * 32;
                        umma_f16(tb + DK_COL, desc_k(adS + offk), desc_mn(aQ + k * 2048), id_kmn, (it | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(qdo_empty);            // Q / dO tiles may be overwritten by the next iteration's TMA
#pragma unroll
                    for (int k = 0; k < 8; ++k)        // contraction over the 128 keys
                        umma_f16(tb + ST_COL, desc_mn(adS + k * 2048), desc_mn(aK + k * 2048), id_mnmn, k != 0 ? 1u : 0u);
                    umma_commit(dq_full);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        const int r = q * 32 + lane;                                   // my TMEM lane: key row (softmax) / query row (dQ)
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        uint8_t* p_row = sP + half * TILE16 + r * 128;
        uint8_t* ds_row = sdS + half * TILE16 + r * 128;
        // per-query vectors (lse * log2(e) for half 0, D for half 1) of iteration `it`, loaded ONE ITERATION AHEAD into a register: the
        // global-load latency used to sit on the critical path of every iteration (ncu: long-scoreboard stalls on the smem store below)
        auto load_vec = [&](int it) -> float {
            const int g = it / (n_q - kt), qt = kt + it % (n_q - kt);
            const int qi = qt * BT + r;
            if (qi >= S) return 0.f;
            const long long idx = ((long long)b * Hq + hk * G + g) * S + qi;
            return half == 0 ? LSE[idx] * 1.4426950408889634f : Dsum[idx];
        };
        float vec_next = load_vec(0);
        for (int it = 0; it < n_it; ++it) {
            const uint32_t ph = (uint32_t)it & 1u;
            const int g = it / (n_q - kt), qt = kt + it % (n_q - kt);
            const int h = hk * G + g;
            const int q0 = qt * BT;
            float* vec = sVec + (it & 1) * 256;          // double buffered
            vec[half * 128 + r] = vec_next;
            if (it + 1 < n_it) vec_next = load_vec(it + 1);
            asm volatile("bar.sync 1, 256;" ::: "memory");
            mbar_wait(s_full, ph);
            tc_fence_after();
            // masks: element (kv = kv0 + r, query = q0 + col) is live iff kv < S, query < S, kv <= query
            const int kv = kv0 + r;
            const bool need_mask = (qt == kt) || (q0 + BT > S) || (kv0 + BT > S);
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                uint32_t st[32], dp[32];
                tmem_ld_32x32(lane_base + ST_COL + half * 64 + c * 32, st);
                tmem_ld_32x32(lane_base + DP_COL + half * 64 + c * 32, dp);
                tmem_ld_wait();
                const float* lse_c = vec + half * 64 + c * 32;
                const float* d_c = vec + 128 + half * 64 + c * 32;
                float pv[32], dsv[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float p = ex2_approx_b(fmaf(__uint_as_float(st[i]), scale_log2, -lse_c[i]));
                    if (need_mask) {
                        const int qi = q0 + half * 64 + c * 32 + i;
                        if (kv >= S || qi >= S || kv > qi) p = 0.f;
                    }
                    pv[i] = p;
                    dsv[i] = p * (__uint_as_float(dp[i]) - d_c[i]);
                }
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {
                    uint4 u, w;
                    u.x = pack_bf16x2(pv[8 * qd + 0], pv[8 * qd + 1]);
                    u.y = pack_bf16x2(pv[8 * qd + 2], pv[8 * qd + 3]);
                    u.z = pack_bf16x2(pv[8 * qd + 4], pv[8 * qd + 5]);
                    u.w = pack_bf16x2(pv[8 * qd + 6], pv[8 * qd + 7]);
                    w.x = pack_bf16x2(dsv[8 * qd + 0], dsv[8 * qd + 1]);
                    w.y = pack_bf16x2(dsv[8 * qd + 2], dsv[8 * qd + 3]);
                    w.z = pack_bf16x2(dsv[8 * qd + 4], dsv[8 * qd + 5]);
                    w.w = pack_bf16x2(dsv[8 * qd + 6], dsv[8 * qd + 7]);
                    const int off = (((c * 4 + qd) ^ (r & 7)) << 4);
                    *reinterpret_cast<uint4*>(p_row + off) = u;
                    *reinterpret_cast<uint4*>(ds_row + off) = w;
                }
            }
            tc_fence_before();
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(pds_full);
            // ---- dQ tile: rows = queries, my 64 of the 128 head dims ----
            mbar_wait(dq_full, ph);
            tc_fence_after();
            {
                const int qi = q0 + r;
                float* dst = dQacc + ((long long)row_base + qi) * dq_rs + (long long)h * HD + half * 64;
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32(lane_base + ST_COL + half * 64 + c * 32, v);
                    tmem_ld_wait();
                    if (qi < S && !(dbg & 1)) {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            red_add_v4(dst + c * 32 + 4 * i, __uint_as_float(v[4 * i]) * scale, __uint_as_float(v[4 * i + 1]) * scale,
                                       __uint_as_float(v[4 * i + 2]) * scale, __uint_as_float(v[4 * i + 3]) * scale);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(dq_empty);
        }
        // ---- epilogue: dV, dK (scaled) -> bf16; the last dq_full also covers every earlier UMMA ----
        const int kv = kv0 + r;
#pragma unroll 1
        for (int which = 0; which < 2; ++which) {
            bf16* dst = (which ? dK : dV) + ((long long)row_base + kv) * (which ? dk_rs : dv_rs) + (long long)hk * HD + half * 64;
            const float mul = which ? scale : 1.0f;
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(lane_base + (which ? DK_COL : DV_COL) + half * 64 + c * 32, v);
                tmem_ld_wait();
                if (kv < S) {
#pragma unroll
                    for (int qd = 0; qd < 4; ++qd) {
                        uint4 u;
                        u.x = pack_bf16x2(__uint_as_float(v[8 * qd + 0]) * mul, __uint_as_float(v[8 * qd + 1]) * mul);
                        u.y = pack_bf16x2(__uint_as_float(v[8 * qd + 2]) * mul, __uint_as_float(v[8 * qd + 3]) * mul);
                        u.z = pack_bf16x2(__uint_as_float(v[8 * qd + 4]) * mul, __uint_as_float(v[8 * qd + 5]) * mul);
                        u.w = pack_bf16x2(__uint_as_float(v[8 * qd + 6]) * mul, __uint_as_float(v[8 * qd + 7]) * mul);
                        *reinterpret_cast<uint4*>(dst + c * 32 + qd * 8) = u;
                    }
                }
            }
        }
        tc_fence_before();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<512>(tmem_base);
}

// =====================================================================================================================
// Variant 2: the same five products on 64-QUERY sub-tiles, software-pipelined.
//   The kernel above is a serial chain per (query tile, key tile): S/dP MMAs -> softmax -> dV/dK/dQ MMAs -> dQ read-out, with dQ
//   aliasing S in a full TMEM and single Q / dO buffers, so the TMA round trip, four barrier hand-offs and the read-out all sit on
//   the critical path (ncu: tensor pipe 23 % active, 11.7 k clk per iteration for ~5 k clk of work).  With 64 queries per step:
//     TMEM   S^T [128 keys x 64 q] cols 0..63 | dP^T 64..127 | dQ^T [128 hd x 64 q] x 2 buffers 128..255 | dV 256..383 | dK 384..511
//     smem   K 32 KB | V 32 KB | (Q 16 KB + dO 16 KB) x 2 stages | P^T 16 KB | dS^T 16 KB = 160 KB
//   * S^T / dP^T of step i+1 are issued as soon as the softmax warps have pulled step i's into registers (s_empty), before the
//     dV / dK / dQ products of step i -- the softmax of step i+1 runs under those;
//   * dQ is computed TRANSPOSED, dQ^T = K^T dS^T (A = K as an MN-major operand, M = head dim), so it is a 128-lane tile with its own
//     two TMEM buffers: its read-out (warp-coalesced red.global.add.f32, 32 consecutive head dims per instruction) happens one step
//     late, under the next step's work, and nothing waits for it;
//   * Q / dO are double buffered, so the TMA latency is hidden.
// =====================================================================================================================
constexpr int BQ2 = 64;
constexpr int TILE8 = 64 * 64 * 2;                    // one [64 rows x 64 bf16] sub-tile
constexpr int QDO_STAGE = 4 * TILE8;                  // Q (2 sub-tiles) + dO (2 sub-tiles) of one step
constexpr int B2_K = 0, B2_V = TILE, B2_QDO = 2 * TILE, B2_P = 2 * TILE + 2 * QDO_STAGE, B2_DS = B2_P + TILE16, B2_BAR = B2_DS + TILE16;
constexpr int SMEM_BWD2 = B2_BAR + 256 + 2 * 128 * 4 + 1024;   // + barriers + (lse, D) double buffered + align slack
constexpr uint32_t ST2_COL = 0, DP2_COL = 64, DQ2_COL = 128, DV2_COL = 256, DK2_COL = 384;

__device__ __forceinline__ uint64_t desc_mn_lbo(uint32_t smem_addr, uint32_t lbo_bytes) {   // MN-major, 64-wide MN atoms lbo apart
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_tc_bwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                    const float* __restrict__ LSE, const float* __restrict__ Dsum, float* __restrict__ dQacc,
                    bf16* __restrict__ dK, bf16* __restrict__ dV, int S, int Hq, int Hkv, long long dq_rs, long long dk_rs,
                    long long dv_rs, float scale, float scale_log2, int dbg, int n_kt) {
    TA_PDL_ENTRY();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
    uint8_t* sK = smem + B2_K;
    uint8_t* sV = smem + B2_V;
    uint8_t* sQdO = smem + B2_QDO;          // stage s: Q at + s * QDO_STAGE, dO at + s * QDO_STAGE + 2 * TILE8
    uint8_t* sP = smem + B2_P;
    uint8_t* sdS = smem + B2_DS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + B2_BAR);
    uint64_t* kv_full = bars;
    uint64_t* qdo_full = bars + 1;          // [2]
    uint64_t* qdo_empty = bars + 3;         // [2]
    uint64_t* s_full = bars + 5;
    uint64_t* s_empty = bars + 6;
    uint64_t* pds_full = bars + 7;
    uint64_t* pds_empty = bars + 8;
    uint64_t* dq_full = bars + 9;           // [2]
    uint64_t* dq_empty = bars + 11;         // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
    float* sVec = reinterpret_cast<float*>(smem + B2_BAR + 256);   // [2 buffers][lse*log2e (64) | D (64)]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // linear grid, longest work first: key tile 0 sees every query sub-tile of the sequence, the last key tile only the last few
    const int per_kt = (int)gridDim.x / n_kt;        // = Hkv * B
    const int kt = (int)blockIdx.x / per_kt, hk = ((int)blockIdx.x % per_kt) % Hkv, b = ((int)blockIdx.x % per_kt) / Hkv;
    const int G = Hq / Hkv;
    const int kv0 = kt * BT;
    const int n_qs = (S + BQ2 - 1) / BQ2;            // 64-query sub-tiles of the sequence
    const int qs_lo = 2 * kt;                        // causal: sub-tiles whose last query >= kv0
    const int per_head = n_qs - qs_lo;
    const int n_it = G * per_head;
    const int row_base = b * S;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
        tma_prefetch_desc(&tmdO);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(kv_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&qdo_full[i], 1);
            mbar_init(&qdo_empty[i], 1);
            mbar_init(&dq_full[i], 1);
            mbar_init(&dq_empty[i], 8);
        }
        mbar_init(s_full, 1);
        mbar_init(s_empty, 8);
        mbar_init(pds_full, 8);
        mbar_init(pds_empty, 1);
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(kv_full, 2 * TILE);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                tma_load_2d(sK + u * TILE16, &tmK, kv_full, hk * HD + u * 64, row_base + kv0);
                tma_load_2d(sV + u * TILE16, &tmV, kv_full, hk * HD + u * 64, row_base + kv0);
            }
            for (int it = 0; it < n_it; ++it) {
                const int g = it / per_head, qs = qs_lo + it % per_head;
                const int h = hk * G + g;
                const int st = it & 1;
                mbar_wait(&qdo_empty[st], (((uint32_t)it >> 1) & 1u) ^ 1u);
                mbar_arrive_expect_tx(&qdo_full[st], QDO_STAGE);
                uint8_t* dst = sQdO + st * QDO_STAGE;
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    tma_load_2d(dst + u * TILE8, &tmQ, &qdo_full[st], h * HD + u * 64, row_base + qs * BQ2);
                    tma_load_2d(dst + 2 * TILE8 + u * TILE8, &tmdO, &qdo_full[st], h * HD + u * 64, row_base + qs * BQ2);
                }
            }
        }
    } else if (warp == 1) {
        // whole warp, warp-uniform operands, one elected lane issues (see the kernel above)
        constexpr uint32_t id_s = umma_idesc_bf16(BT, BQ2);                                 // S^T, dP^T: A, B K-major, N = 64
        constexpr uint32_t id_dv = umma_idesc_bf16(BT, HD) | (1u << 16);                    // dV, dK: B MN-major, N = 128
        constexpr uint32_t id_dq = umma_idesc_bf16(HD, BQ2) | (1u << 15) | (1u << 16);      // dQ^T: A and B MN-major, N = 64
        const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
        const bool issuer = elect_one();
        const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP), adS = smem_u32(sdS), aQdO = smem_u32(sQdO);
        auto issue_s = [&](int it) {          // S^T = K Q^T, dP^T = V dO^T for step `it`
            const int st = it & 1;
            mbar_wait(&qdo_full[st], ((uint32_t)it >> 1) & 1u);
            mbar_wait(s_empty, ((uint32_t)it & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t aQ = aQdO + st * QDO_STAGE, adO = aQ + 2 * TILE8;
            if (issuer) {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    umma_f16(tb + ST2_COL, desc_k(aK + (k >> 2) * TILE16 + (k & 3) * 32), desc_k(aQ + (k >> 2) * TILE8 + (k & 3) * 32), id_s,
                             k != 0 ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    umma_f16(tb + DP2_COL, desc_k(aV + (k >> 2) * TILE16 + (k & 3) * 32), desc_k(adO + (k >> 2) * TILE8 + (k & 3) * 32), id_s,
                             k != 0 ? 1u : 0u);
                umma_commit(s_full);
            }
            __syncwarp();
        };
        mbar_wait(kv_full, 0);
        if (n_it > 0) issue_s(0);
        for (int it = 0; it < n_it; ++it) {
            if (it + 1 < n_it) issue_s(it + 1);
            const int st = it & 1;
            const uint32_t aQ = aQdO + st * QDO_STAGE, adO = aQ + 2 * TILE8;
            mbar_wait(pds_full, (uint32_t)it & 1u);
            mbar_wait(&dq_empty[st], (((uint32_t)it >> 1) & 1u) ^ 1u);      // dQ^T buffer `st` was read out (step it - 2)
            tc_fence_after();
            if (issuer) {
#pragma unroll
                for (int k = 0; k < 4; ++k)        // contraction over the 64 queries, 16 per step
                    umma_f16(tb + DV2_COL, desc_k(aP + k * 32), desc_mn_lbo(adO + k * 2048, TILE8), id_dv, (it | k) != 0 ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_f16(tb + DK2_COL, desc_k(adS + k * 32), desc_mn_lbo(aQ + k * 2048, TILE8), id_dv, (it | k) != 0 ? 1u : 0u);
                umma_commit(&qdo_empty[st]);       // this stage's Q / dO may be overwritten
#pragma unroll
                for (int k = 0; k < 8; ++k)        // contraction over the 128 keys: dQ^T [hd x q] = K^T dS^T
                    umma_f16(tb + DQ2_COL + (uint32_t)(st * 64), desc_mn_lbo(aK + k * 2048, TILE16), desc_mn_lbo(adS + k * 2048, TILE16), id_dq,
                             k != 0 ? 1u : 0u);
                umma_commit(&dq_full[st]);
                umma_commit(pds_empty);            // P^T / dS^T tiles may be overwritten
            }
            __syncwarp();
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        const int r = q * 32 + lane;                                   // my TMEM lane: key row (softmax) / head dim (dQ^T)
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        uint8_t* p_row = sP + r * 128;
        uint8_t* ds_row = sdS + r * 128;
        const int etid = (warp - 4) * 32 + lane;                       // 0..255
        // (lse * log2 e | D) of the step's 64 queries: threads 0-63 / 64-127 load one value each, one step ahead, into a register
        auto load_vec = [&](int it) -> float {
            if (etid >= 128) return 0.f;
            const int g = it / per_head, qs = qs_lo + it % per_head;
            const int qi = qs * BQ2 + (etid & 63);
            if (qi >= S) return 0.f;
            const long long idx = ((long long)b * Hq + hk * G + g) * S + qi;
            return etid < 64 ? LSE[idx] * 1.4426950408889634f : Dsum[idx];
        };
        // read-out of dQ^T buffer (step `it`): rows = head dims (my lane), my 32 of the 64 queries
        auto readout_dq = [&](int it) {
            const int st = it & 1;
            const int g = it / per_head, qs = qs_lo + it % per_head;
            const int h = hk * G + g;
            mbar_wait(&dq_full[st], ((uint32_t)it >> 1) & 1u);
            tc_fence_after();
            uint32_t v[32];
            tmem_ld_32x32(lane_base + DQ2_COL + (uint32_t)(st * 64 + half * 32), v);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&dq_empty[st]);
            if (!(dbg & 1)) {
                const int qi0 = qs * BQ2 + half * 32;
                float* dst = dQacc + ((long long)row_base + qi0) * dq_rs + (long long)h * HD + r;
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (qi0 + i < S) atomicAdd(dst + (long long)i * dq_rs, __uint_as_float(v[i]) * scale);
            }
        };
        float vec_next = (n_it > 0) ? load_vec(0) : 0.f;
        for (int it = 0; it < n_it; ++it) {
            const uint32_t ph = (uint32_t)it & 1u;
            const int qs = qs_lo + it % per_head;
            const int q0 = qs * BQ2;
            float* vec = sVec + (it & 1) * 128;          // double buffered: [lse (64) | D (64)]
            if (etid < 128) vec[etid] = vec_next;
            if (it + 1 < n_it) vec_next = load_vec(it + 1);
            asm volatile("bar.sync 1, 256;" ::: "memory");
            mbar_wait(s_full, ph);
            tc_fence_after();
            uint32_t st[32], dp[32];
            tmem_ld_32x32(lane_base + ST2_COL + (uint32_t)(half * 32), st);
            tmem_ld_32x32(lane_base + DP2_COL + (uint32_t)(half * 32), dp);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_empty);          // S^T / dP^T are in registers: the next step's may be issued
            // masks: element (kv = kv0 + r, query = q0 + col) is live iff kv < S, query < S, kv <= query
            const int kv = kv0 + r;
            const bool need_mask = (q0 < kv0 + BT) || (q0 + BQ2 > S) || (kv0 + BT > S);
            const float* lse_c = vec + half * 32;
            const float* d_c = vec + 64 + half * 32;
            float pv[32], dsv[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                float p = ex2_approx_b(fmaf(__uint_as_float(st[i]), scale_log2, -lse_c[i]));
                if (need_mask) {
                    const int qi = q0 + half * 32 + i;
                    if (kv >= S || qi >= S || kv > qi) p = 0.f;
                }
                pv[i] = p;
                dsv[i] = p * (__uint_as_float(dp[i]) - d_c[i]);
            }
            if (it > 0) {
                mbar_wait(pds_empty, (uint32_t)(it - 1) & 1u);          // the previous step's products have consumed P^T / dS^T
            }
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
                uint4 u, w;
                u.x = pack_bf16x2(pv[8 * qd + 0], pv[8 * qd + 1]);
                u.y = pack_bf16x2(pv[8 * qd + 2], pv[8 * qd + 3]);
                u.z = pack_bf16x2(pv[8 * qd + 4], pv[8 * qd + 5]);
                u.w = pack_bf16x2(pv[8 * qd + 6], pv[8 * qd + 7]);
                w.x = pack_bf16x2(dsv[8 * qd + 0], dsv[8 * qd + 1]);
                w.y = pack_bf16x2(dsv[8 * qd + 2], dsv[8 * qd + 3]);
                w.z = pack_bf16x2(dsv[8 * qd + 4], dsv[8 * qd + 5]);
                w.w = pack_bf16x2(dsv[8 * qd + 6], dsv[8 * qd + 7]);
                const int off = (((half * 4 + qd) ^ (r & 7)) << 4);
                *reinterpret_cast<uint4*>(p_row + off) = u;
                *reinterpret_cast<uint4*>(ds_row + off) = w;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(pds_full);
            if (it > 0) readout_dq(it - 1);              // one step late: runs under this step's products
        }
        if (n_it > 0) readout_dq(n_it - 1);
        // ---- epilogue: dV, dK (scaled) -> bf16; the last dq_full also covers every earlier UMMA ----
        const int kv = kv0 + r;
#pragma unroll 1
        for (int which = 0; which < 2; ++which) {
            bf16* dst = (which ? dK : dV) + ((long long)row_base + kv) * (which ? dk_rs : dv_rs) + (long long)hk * HD + half * 64;
            const float mul = which ? scale : 1.0f;
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(lane_base + (which ? DK2_COL : DV2_COL) + half * 64 + c * 32, v);
                tmem_ld_wait();
                if (kv < S) {
#pragma unroll
                    for (int qd = 0; qd < 4; ++qd) {
                        uint4 u;
                        u.x = pack_bf16x2(__uint_as_float(v[8 * qd + 0]) * mul, __uint_as_float(v[8 * qd + 1]) * mul);
                        u.y = pack_bf16x2(__uint_as_float(v[8 * qd + 2]) * mul, __uint_as_float(v[8 * qd + 3]) * mul);
                        u.z = pack_bf16x2(__uint_as_float(v[8 * qd + 4]) * mul, __uint_as_float(v[8 * qd + 5]) * mul);
                        u.w = pack_bf16x2(__uint_as_float(v[8 * qd + 6]) * mul, __uint_as_float(v[8 * qd + 7]) * mul);
                        *reinterpret_cast<uint4*>(dst + c * 32 + qd * 8) = u;
                    }
                }
            }
        }
        tc_fence_before();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<512>(tmem_base);
}

int g_bwd_variant = 2;   // ta_attn_set_bwd_variant: 2 (default) = 64-query pipelined kernel; 1 = the 128-query kernel above
int g_bwd_dbg = 0;
}  // namespace

extern int g_norm_wpb;        // elementwise.cu
int g_wgrad_transposed = 0;   // key 2: 1 = weight-gradient GEMMs through transposed operand copies (A/B reference of the TN kernel)
TA_API int ta_attn_set_bwd_variant(int v) {
    g_bwd_variant = (v == 1) ? 1 : 2;
    return 0;
}
TA_API int ta_debug_set(int key, int value) {   // experiments only (key 1: attention-backward switches)
    if (key == 1) g_bwd_dbg = value;
    if (key == 2) g_wgrad_transposed = value;
    if (key == 3 && value >= 1 && value <= 8) g_norm_wpb = value;
    return 0;
}

// internal: tcgen05 backward for head_dim 128 + causal; *handled = 0 -> caller uses the mma.sync kernel.
// dsum (rowsum(dO*O)) must already be computed and dq_acc zeroed by the caller.
int k_attn_tc_bwd(const bf16* q, const bf16* k, const bf16* v, const bf16* d_o, const float* lse, const float* dsum, float* dq_acc,
                  bf16* dk, bf16* dv, int B, int S, int Hq, int Hkv, int head_dim, long long q_rs, long long k_rs, long long v_rs,
                  long long do_rs, long long dq_rs, long long dk_rs, long long dv_rs, int causal, float scale, cudaStream_t st,
                  int* handled) {
    *handled = 0;
    if (head_dim != HD || !causal || Hq % Hkv != 0 || S < 1) return 0;
    if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
         reinterpret_cast<uintptr_t>(d_o) | reinterpret_cast<uintptr_t>(dq_acc) | reinterpret_cast<uintptr_t>(dk) |
         reinterpret_cast<uintptr_t>(dv)) & 15)
        return 0;
    if ((dq_rs % 4) || (dk_rs % 8) || (dv_rs % 8)) return 0;
    CUtensorMap tq, tk, tv, tdo;
    const long long rows = (long long)B * S;
    int rc = k_make_tensor_map_2d(&tq, q, rows, (long long)Hq * HD, q_rs, BT);
    if (rc) return rc;
    rc = k_make_tensor_map_2d(&tk, k, rows, (long long)Hkv * HD, k_rs, BT);
    if (rc) return rc;
    rc = k_make_tensor_map_2d(&tv, v, rows, (long long)Hkv * HD, v_rs, BT);
    if (rc) return rc;
    rc = k_make_tensor_map_2d(&tdo, d_o, rows, (long long)Hq * HD, do_rs, BT);
    if (rc) return rc;
    static bool done = false;
    if (!done) {
        TA_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BWD));
        TA_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BWD2));
        done = true;
    }
    dim3 grid((S + BT - 1) / BT, Hkv, B);
    if (g_bwd_variant == 2) {
        CUtensorMap tq2, tdo2;                 // 64-row boxes of Q and dO
        rc = k_make_tensor_map_2d(&tq2, q, rows, (long long)Hq * HD, q_rs, BQ2);
        if (rc) return rc;
        rc = k_make_tensor_map_2d(&tdo2, d_o, rows, (long long)Hq * HD, do_rs, BQ2);
        if (rc) return rc;
        const int n_kt = (S + BT - 1) / BT;
        TA_KERNEL_LAUNCH(attn_tc_bwd2_kernel, (unsigned)(n_kt * Hkv * B), BWD_THREADS, SMEM_BWD2, st, tq2, tk, tv, tdo2, lse, dsum, dq_acc, dk, dv, S,
                         Hq, Hkv, dq_rs, dk_rs, dv_rs, scale, scale * 1.4426950408889634f, g_bwd_dbg, n_kt);
        *handled = 1;
        return 0;
    }
    TA_KERNEL_LAUNCH(attn_tc_bwd_kernel, grid, BWD_THREADS, SMEM_BWD, st, tq, tk, tv, tdo, lse, dsum, dq_acc, dk, dv, S, Hq, Hkv, dq_rs, dk_rs,
                     dv_rs, scale, scale * 1.4426950408889634f, g_bwd_dbg);
    *handled = 1;
    return 0;
}
