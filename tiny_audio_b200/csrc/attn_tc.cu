// tcgen05 flash-attention forward for the GLM-ASR encoder shape: head_dim 64, non-causal, no mask, S <= 1500
// (HF:models/glmasr/modeling_glmasr.py:208-221).  Successor of attn_fwd_kernel<64,false> in attn_mma.cu, which tops
// out at the legacy mma.sync rate (~0.5 PFLOP/s on B200; profiles/r01_*): here both GEMMs of the attention run on
// the 5th-gen tensor cores with TMEM accumulators, so the kernel is bound by the exp2 rate of the SFUs instead.
//
// One CTA = one (batch, head, 128-query tile); two CTAs are resident per SM so that one CTA's softmax overlaps the
// other's MMAs.  Roles:  warp 0 TMA producer (Q once; K_j / V_j through a 3-slot ring of 16 KB tiles),
//                        warp 1 MMA issuer   (S_j = Q K_j^T -> TMEM[0,128);  O += P_j V_j -> TMEM[128,192)),
//                        warp 2 TMEM allocator,  warps 4-7 softmax (thread = query row = TMEM lane).
// Softmax keeps a lazily updated reference max (rescale O only when the row max grows by > 2^8, the FA-4 trick),
// reads S twice from TMEM (max pass, exp pass) to stay under 128 registers, and writes P as bf16 into shared memory
// in the K-major SWIZZLE_128B layout the A-operand descriptor expects.  V is consumed as an MN-major B operand
// straight from its row-major [kv][64] TMA tile.
#include <type_traits>

#include "common.cuh"
#include "kernels.cuh"
#include "tinyaudio_b200.h"

namespace {

constexpr int BQ = 128, BKV = 128;
constexpr int TILE16 = 128 * 64 * 2;       // one [128 rows x 64 bf16] SWIZZLE_128B sub-tile
constexpr int ATT_THREADS = 384;           // warps 0-3: TMA / MMA / TMEM alloc / spare; warps 4-11: softmax (2 threads per row)
constexpr int TMEM_COLS_ATT = 256;
constexpr uint32_t S_COL = 0, O_COL = 128;

template <int HD>
struct AttCfg {
    static constexpr int NSUB = HD / 64;
    static constexpr int KV_SLOTS = (HD == 64) ? 3 : 4;               // K/V ring depth (HD 64: two CTAs per SM must fit)
    static constexpr int TILE_BYTES = NSUB * TILE16;                  // Q, K or V tile: 128 rows x HD
    static constexpr int P_BYTES = 2 * TILE16;                        // P: 128 x 128 bf16
    static constexpr int BAR_OFF = TILE_BYTES * (1 + KV_SLOTS) + P_BYTES;
    static constexpr int SMEM = BAR_OFF + 256 + 3 * 1024 + 1024;      // + barriers + exchange floats + 1 KB alignment slack
    static constexpr int CTAS_PER_SM = (HD == 64) ? 2 : 1;
};

// fp32 pair -> packed bf16 pair on the integer pipes (round half up).  Experiment (-DTA_ATTN_PACK_ALU): ncu's
// sm__inst_executed_pipe_xu_realtime (85 %) suggested that cvt.rn.bf16x2.f32 (F2FP) shares the XU pipe with MUFU.EX2; the
// micro-benchmark tools/ubench/sfu4.cu says otherwise -- EX2 issues once per 8.0 clk and sub-partition with or without one F2FP per
// two EX2 -- and the kernel got SLOWER with the integer pack (0.589 -> 0.621 ms: three issue slots instead of one).  Default: cvt.
__device__ __forceinline__ uint32_t pack_bf16x2_alu(float lo, float hi) {
    const uint32_t a = __float_as_uint(lo) + 0x8000u, b = __float_as_uint(hi) + 0x8000u;
    return __byte_perm(a, b, 0x7632);       // (b & 0xffff0000) | (a >> 16)
}
#ifdef TA_ATTN_PACK_ALU
#define PACK_P pack_bf16x2_alu
#else
#define PACK_P pack_bf16x2
#endif

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MN-major SWIZZLE_128B operand (rows = K index, 64 bf16 = 128 B of the MN index per row; 8-row groups 1024 B apart;
// further 64-wide MN atoms `lbo_bytes` apart)
__device__ __forceinline__ uint64_t umma_desc_sw128_mnmajor(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

template <int HD, bool CAUSAL>
__global__ void __launch_bounds__(ATT_THREADS, AttCfg<HD>::CTAS_PER_SM)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, bf16* __restrict__ O, float* __restrict__ LSE, int S, int Hq, int Hkv,
                   long long o_rs, float scale_log2, const int* __restrict__ kv_start) {
    TA_PDL_ENTRY();
    using C = AttCfg<HD>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
    uint8_t* sQ = smem;
    uint8_t* sKV = smem + C::TILE_BYTES;
    uint8_t* sP = smem + C::TILE_BYTES * (1 + C::KV_SLOTS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;
    uint64_t* kv_empty = bars + 1 + C::KV_SLOTS;
    uint64_t* s_full = bars + 1 + 2 * C::KV_SLOTS;
    uint64_t* s_empty = s_full + 1;
    uint64_t* p_full = s_full + 2;
    uint64_t* pv_done = s_full + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 4);
    float* sXch = reinterpret_cast<float*>(smem + C::BAR_OFF + 256);   // [3][2][128]: row max (double buffered) / row sum exchange

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int hk = h / (Hq / Hkv);
    const int q0 = qt * BQ;
    const int n_kv_all = (S + BKV - 1) / BKV;
    const int n_kv = CAUSAL ? min(n_kv_all, qt + 1) : n_kv_all;
    const int row_base = b * S;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < C::KV_SLOTS; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(s_empty, 8);
        mbar_init(p_full, 8);
        mbar_init(pv_done, 1);
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc<TMEM_COLS_ATT>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        if (warp == 0 && lane == 0) {
            mbar_arrive_expect_tx(q_full, C::TILE_BYTES);
#pragma unroll
            for (int u = 0; u < C::NSUB; ++u) tma_load_2d(sQ + u * TILE16, &tmQ, q_full, h * HD + u * 64, row_base + q0);
            for (int i = 0; i < 2 * n_kv; ++i) {
                const int slot = i % C::KV_SLOTS;
                const uint32_t ph = (uint32_t)(i / C::KV_SLOTS) & 1u;
                mbar_wait(&kv_empty[slot], ph ^ 1);
                mbar_arrive_expect_tx(&kv_full[slot], C::TILE_BYTES);
                const int j = i >> 1;
#pragma unroll
                for (int u = 0; u < C::NSUB; ++u)
                    tma_load_2d(sKV + slot * C::TILE_BYTES + u * TILE16, (i & 1) ? &tmV : &tmK, &kv_full[slot], hk * HD + u * 64,
                                row_base + j * BKV);
            }
        } else if (warp == 1) {
            // whole warp, warp-uniform operands, one elected lane issues the tcgen05 instructions (an `if (lane == 0)` region costs
            // ~100 clk of ELECT / R2UR.BROADCAST plumbing per UTCHMMA: tools/attn_trace.py)
            constexpr uint32_t idesc_s = umma_idesc_bf16(BQ, BKV);
            constexpr uint32_t idesc_o = umma_idesc_bf16(BQ, HD) | (1u << 16);   // B operand (V) is MN-major
            const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
            const bool issuer = elect_one();
            const uint32_t q_addr = smem_u32(sQ), p_addr = smem_u32(sP);
            auto issue_s = [&](int j) {
                const int i = 2 * j, slot = i % C::KV_SLOTS;
                mbar_wait(&kv_full[slot], (uint32_t)(i / C::KV_SLOTS) & 1u);
                mbar_wait(s_empty, ((uint32_t)j & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t k_addr = smem_u32(sKV + slot * C::TILE_BYTES);
                if (issuer) {
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k) {
                        const uint32_t off = (k >> 2) * TILE16 + (k & 3) * 32;
                        umma_f16(tb + S_COL, umma_desc_sw128_kmajor(q_addr + off), umma_desc_sw128_kmajor(k_addr + off), idesc_s,
                                 k != 0 ? 1u : 0u);
                    }
                    umma_commit(s_full);
                    umma_commit(&kv_empty[slot]);
                }
                __syncwarp();
            };
            mbar_wait(q_full, 0);
            issue_s(0);
            for (int j = 0; j < n_kv; ++j) {
                if (j + 1 < n_kv) issue_s(j + 1);
                const int i = 2 * j + 1, slot = i % C::KV_SLOTS;
                mbar_wait(p_full, (uint32_t)j & 1u);
                mbar_wait(&kv_full[slot], (uint32_t)(i / C::KV_SLOTS) & 1u);
                tc_fence_after();
                const uint32_t v_addr = smem_u32(sKV + slot * C::TILE_BYTES);
                if (issuer) {
#pragma unroll
                    for (int k = 0; k < BKV / 16; ++k) {
                        const uint32_t pa = p_addr + (k >> 2) * TILE16 + (k & 3) * 32;
                        umma_f16(tb + O_COL, umma_desc_sw128_kmajor(pa), umma_desc_sw128_mnmajor(v_addr + k * 2048, TILE16), idesc_o,
                                 (j | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(pv_done);
                    umma_commit(&kv_empty[slot]);
                }
                __syncwarp();
            }
        }
    } else {
        // two threads per query row: warp (4+q) and warp (8+q) share TMEM lane quadrant q and split the columns
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        const int r = q * 32 + lane;
        constexpr int OW = HD / 2;                   // O columns owned by this thread
        const uint32_t t_s = tmem_base + ((uint32_t)(q * 32) << 16) + S_COL + (uint32_t)(half * 64);
        const uint32_t t_o = tmem_base + ((uint32_t)(q * 32) << 16) + O_COL + (uint32_t)(half * OW);
        float m_ref = -INFINITY, l_sum = 0.f;
        uint8_t* p_row = sP + half * TILE16 + r * 128;
        const uint32_t bar_id = 1 + q;
        // left-padded prompts (generate() with ragged prompts): keys before kv_start[b] are padding.  A real query row (>= kv_start)
        // must not see them; a padding row keeps plain causal attention so that its (unused) output stays finite.
        const int k_lo = (kv_start != nullptr && q0 + r >= kv_start[b]) ? kv_start[b] : 0;
        for (int j = 0; j < n_kv; ++j) {
            mbar_wait(s_full, (uint32_t)j & 1u);
            tc_fence_after();
            // my columns e = lo..lim are real (unmasked) keys
            int lim = S - j * BKV - half * 64 - 1;
            if (CAUSAL && j == qt) lim = min(lim, r - half * 64);
            lim = min(lim, 63);
            const int lo = k_lo - j * BKV - half * 64;
            const bool full_tile = __all_sync(0xffffffffu, lim >= 63 && lo <= 0);
            uint32_t v0[32], v1[32];
            tmem_ld_32x32(t_s, v0);
            tmem_ld_32x32(t_s + 32, v1);
            tmem_ld_wait();
            // S now lives in registers: release its TMEM columns so the MMA warp can issue S_{j+1} during this softmax
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_empty);
            float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            if (full_tile) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(v0[i]));
                    m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(v1[i]));
                }
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (i > lim || i < lo) v0[i] = 0xff800000u;          // -inf
                    if (32 + i > lim || 32 + i < lo) v1[i] = 0xff800000u;
                    m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(v0[i]));
                    m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(v1[i]));
                }
            }
            float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
            float* xch = sXch + (j & 1) * 256;
            xch[half * 128 + r] = mx;
            asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
            mx = fmaxf(mx, xch[(half ^ 1) * 128 + r]) * scale_log2;
            const bool grow = mx > m_ref + 8.0f;
            const float m_new = grow ? mx : m_ref;
            const float alpha = (grow && j > 0) ? ex2_approx(m_ref - m_new) : 1.0f;   // m_ref = -inf (all keys masked so far): 0, O and l are 0
            m_ref = m_new;
            if (j > 0) {
                mbar_wait(pv_done, (uint32_t)(j - 1) & 1u);     // PV_{j-1} has consumed sP and finished updating O
                tc_fence_after();
            }
            // ---- P = exp2(S * scale - m_ref) -> bf16 -> swizzled smem (my 64 columns = sub-tile `half`) ----
            float l4[4] = {0.f, 0.f, 0.f, 0.f};
            const float neg_m = (m_ref == -INFINITY) ? 0.f : -m_ref;     // a tile of padding keys only: exp2(-inf) = 0, never inf - inf
#pragma unroll
            for (int c = 0; c < 2; ++c) {
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {
                    float e[8];
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        const float x = __uint_as_float(c ? v1[8 * qd + t] : v0[8 * qd + t]);
                        e[t] = ex2_approx(fmaf(x, scale_log2, neg_m));     // exp2(-inf) = 0 for masked columns
                        l4[t & 3] += e[t];
                    }
                    uint4 u;
                    u.x = PACK_P(e[0], e[1]);
                    u.y = PACK_P(e[2], e[3]);
                    u.z = PACK_P(e[4], e[5]);
                    u.w = PACK_P(e[6], e[7]);
                    const int k16 = c * 4 + qd;
                    *reinterpret_cast<uint4*>(p_row + ((k16 ^ (r & 7)) << 4)) = u;
                }
            }
            // lazy rescale of O (rare): done after the exp phase so that S's registers are dead by now
            if (j > 0 && __any_sync(0xffffffffu, grow)) {
#pragma unroll 1
                for (int c = 0; c < OW / 32; ++c) {
                    uint32_t o[32];
                    tmem_ld_32x32(t_o + c * 32, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                    tmem_st_32x32(t_o + c * 32, o);
                }
                tmem_st_wait();
                l_sum *= alpha;
            }
            l_sum += (l4[0] + l4[1]) + (l4[2] + l4[3]);
            tc_fence_before();
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        // ---- epilogue: O / l -> bf16 (my OW of the HD head dims) ----
        float* xch = sXch + 512;
        xch[half * 128 + r] = l_sum;
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
        l_sum += xch[(half ^ 1) * 128 + r];
        mbar_wait(pv_done, (uint32_t)(n_kv - 1) & 1u);
        tc_fence_after();
        const int row = q0 + r;
        const float inv = 1.0f / l_sum;
        bf16* orow = O + ((long long)row_base + row) * o_rs + (long long)h * HD + half * OW;
#pragma unroll 1
        for (int c = 0; c < OW / 32; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(t_o + c * 32, v);
            tmem_ld_wait();
            if (row < S) {
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {
                    uint4 u;
                    u.x = pack_bf16x2(__uint_as_float(v[8 * qd + 0]) * inv, __uint_as_float(v[8 * qd + 1]) * inv);
                    u.y = pack_bf16x2(__uint_as_float(v[8 * qd + 2]) * inv, __uint_as_float(v[8 * qd + 3]) * inv);
                    u.z = pack_bf16x2(__uint_as_float(v[8 * qd + 4]) * inv, __uint_as_float(v[8 * qd + 5]) * inv);
                    u.w = pack_bf16x2(__uint_as_float(v[8 * qd + 6]) * inv, __uint_as_float(v[8 * qd + 7]) * inv);
                    *reinterpret_cast<uint4*>(orow + c * 32 + qd * 8) = u;
                }
            }
        }
        if (LSE && half == 0 && row < S) LSE[((long long)b * Hq + h) * S + row] = (m_ref + log2f(l_sum)) * 0.69314718055994531f;
        tc_fence_before();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<TMEM_COLS_ATT>(tmem_base);
}

template <int HD, bool CAUSAL>
int launch_attn_tc(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, bf16* o, float* lse, int B, int S, int Hq,
                   int Hkv, long long o_rs, float scale, cudaStream_t st, const int* kv_start = nullptr) {
    using C = AttCfg<HD>;
    auto kern = attn_tc_fwd_kernel<HD, CAUSAL>;
    static bool done = false;
    if (!done) {
        TA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        TA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        done = true;
    }
    dim3 grid((S + BQ - 1) / BQ, Hq, B);
    TA_KERNEL_LAUNCH(kern, grid, ATT_THREADS, C::SMEM, st, tq, tk, tv, o, lse, S, Hq, Hkv, o_rs, scale * 1.4426950408889634f, kv_start);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Variant 2 (encoder shape, head_dim 64): one thread per query row, 6 warps per CTA, two CTAs per SM.
//   warps 0-3 softmax (thread = query row = TMEM lane; the whole 128-column S row lives in registers, so S's TMEM columns
//   are handed back to the MMA warp right after the load and there is no cross-warp row-max exchange), warp 4 TMA
//   producer + TMEM allocator, warp 5 MMA issuer.
//   POLY > 0: every POLY-th exponential of a full tile is evaluated on the FMA pipe (Cody-Waite range reduction + a
//   degree-3 minimax polynomial, max rel. error 7.5e-5 -- far inside the bf16 rounding of P) instead of MUFU.EX2: the
//   kernel is bound by the 16 exp2/clk/SM of the SFUs (tools/ubench/sfu.cu: a 3:1 mix sustains 20.8/clk/SM).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int ATT1_THREADS = 192;

__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -126.0f);
    const float t = x + 12582912.0f;                 // 1.5 * 2^23: the low mantissa bits now hold round(x)
    const float f = x - (t - 12582912.0f);           // f in [-0.5, 0.5]
    float p = fmaf(f, 0.05517143756151199f, 0.24261081218719482f);
    p = fmaf(p, f, 0.6932609677314758f);
    p = fmaf(p, f, 0.9999281167984009f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

template <int HD>
struct Att1Cfg {
    static constexpr int NSUB = HD / 64;
    // K/V ring depth 4: K_{j+1} re-uses the slot of K_{j-1} (free once S_{j-1} has been issued), so its TMA load starts
    // almost two tiles ahead; with 3 slots it had to wait for PV_{j-1} and S_{j+1} arrived late (ncu: 13 % of the softmax
    // warps' samples were the s_full wait).  2 CTAs x (16 K Q + 64 K ring + 32 K P + 128 B) + 2 x 1 K reserved < 228 K.
    static constexpr int KV_SLOTS = 4;
    static constexpr int TILE_BYTES = NSUB * TILE16;
    static constexpr int P_BYTES = 2 * TILE16;
    static constexpr int BAR_OFF = TILE_BYTES * (1 + KV_SLOTS) + P_BYTES;
    static constexpr int SMEM = BAR_OFF + 128;
};

template <int HD, bool CAUSAL, int POLY>
__global__ void __launch_bounds__(ATT1_THREADS, 2)
attn_tc_fwd1_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, bf16* __restrict__ O, float* __restrict__ LSE, int S, int Hq, int Hkv,
                    long long o_rs, float scale_log2) {
    TA_PDL_ENTRY();
    using C = Att1Cfg<HD>;
    extern __shared__ __align__(1024) uint8_t smem_al[];
    uint8_t* smem = smem_al;
    if (smem_u32(smem) & 1023u) __trap();                 // SWIZZLE_128B tiles need 1024-byte alignment
    uint8_t* sQ = smem;
    uint8_t* sKV = smem + C::TILE_BYTES;
    uint8_t* sP = smem + C::TILE_BYTES * (1 + C::KV_SLOTS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;
    uint64_t* kv_empty = bars + 1 + C::KV_SLOTS;
    uint64_t* s_full = bars + 1 + 2 * C::KV_SLOTS;
    uint64_t* s_empty = s_full + 1;
    uint64_t* p_full = s_full + 2;
    uint64_t* pv_done = s_full + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int hk = h / (Hq / Hkv);
    const int q0 = qt * BQ;
    const int n_kv_all = (S + BKV - 1) / BKV;
    const int n_kv = CAUSAL ? min(n_kv_all, qt + 1) : n_kv_all;
    const int row_base = b * S;

    if (warp == 4) {
        if (lane == 0) {
            tma_prefetch_desc(&tmQ);
            tma_prefetch_desc(&tmK);
            tma_prefetch_desc(&tmV);
            mbar_init(q_full, 1);
            for (int s = 0; s < C::KV_SLOTS; ++s) {
                mbar_init(&kv_full[s], 1);
                mbar_init(&kv_empty[s], 1);
            }
            mbar_init(s_full, 1);
            mbar_init(s_empty, 4);
            mbar_init(p_full, 4);
            mbar_init(pv_done, 1);
            mbar_fence_init();
        }
        __syncwarp();
        tmem_alloc<TMEM_COLS_ATT>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, C::TILE_BYTES);
#pragma unroll
            for (int u = 0; u < C::NSUB; ++u) tma_load_2d(sQ + u * TILE16, &tmQ, q_full, h * HD + u * 64, row_base + q0);
            for (int i = 0; i < 2 * n_kv; ++i) {
                const int slot = i % C::KV_SLOTS;
                const uint32_t ph = (uint32_t)(i / C::KV_SLOTS) & 1u;
                mbar_wait(&kv_empty[slot], ph ^ 1);
                mbar_arrive_expect_tx(&kv_full[slot], C::TILE_BYTES);
                const int j = i >> 1;
#pragma unroll
                for (int u = 0; u < C::NSUB; ++u)
                    tma_load_2d(sKV + slot * C::TILE_BYTES + u * TILE16, (i & 1) ? &tmV : &tmK, &kv_full[slot], hk * HD + u * 64,
                                row_base + j * BKV);
            }
        }
    } else if (warp == 5) {
        {   // whole warp, warp-uniform operands, one elected lane issues (see attn_tc_fwd_kernel)
            constexpr uint32_t idesc_s = umma_idesc_bf16(BQ, BKV);
            constexpr uint32_t idesc_o = umma_idesc_bf16(BQ, HD) | (1u << 16);   // B operand (V) is MN-major
            const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
            const bool issuer = elect_one();
            const uint32_t q_addr = smem_u32(sQ), p_addr = smem_u32(sP);
            auto issue_s = [&](int j) {
                const int i = 2 * j, slot = i % C::KV_SLOTS;
                mbar_wait(&kv_full[slot], (uint32_t)(i / C::KV_SLOTS) & 1u);
                mbar_wait(s_empty, ((uint32_t)j & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t k_addr = smem_u32(sKV + slot * C::TILE_BYTES);
                if (issuer) {
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k) {
                        const uint32_t off = (k >> 2) * TILE16 + (k & 3) * 32;
                        umma_f16(tb + S_COL, umma_desc_sw128_kmajor(q_addr + off), umma_desc_sw128_kmajor(k_addr + off), idesc_s,
                                 k != 0 ? 1u : 0u);
                    }
                    umma_commit(s_full);
                    umma_commit(&kv_empty[slot]);
                }
                __syncwarp();
            };
            mbar_wait(q_full, 0);
            issue_s(0);
            for (int j = 0; j < n_kv; ++j) {
                if (j + 1 < n_kv) issue_s(j + 1);
                const int i = 2 * j + 1, slot = i % C::KV_SLOTS;
                mbar_wait(p_full, (uint32_t)j & 1u);
                mbar_wait(&kv_full[slot], (uint32_t)(i / C::KV_SLOTS) & 1u);
                tc_fence_after();
                const uint32_t v_addr = smem_u32(sKV + slot * C::TILE_BYTES);
                if (issuer) {
#pragma unroll
                    for (int k = 0; k < BKV / 16; ++k) {
                        const uint32_t pa = p_addr + (k >> 2) * TILE16 + (k & 3) * 32;
                        umma_f16(tb + O_COL, umma_desc_sw128_kmajor(pa), umma_desc_sw128_mnmajor(v_addr + k * 2048, TILE16), idesc_o,
                                 (j | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(pv_done);
                    umma_commit(&kv_empty[slot]);
                }
                __syncwarp();
            }
        }
    } else {
        const int r = warp * 32 + lane;                     // query row of the tile = TMEM lane
        const uint32_t t_s = tmem_base + ((uint32_t)(warp * 32) << 16) + S_COL;
        const uint32_t t_o = tmem_base + ((uint32_t)(warp * 32) << 16) + O_COL;
        float m_ref = -INFINITY, l_sum = 0.f;
        uint8_t* p_row = sP + r * 128;
        for (int j = 0; j < n_kv; ++j) {
            mbar_wait(s_full, (uint32_t)j & 1u);
            tc_fence_after();
            int lim = S - j * BKV - 1;                      // my columns e = 0..127 are real (unmasked) keys iff e <= lim
            if (CAUSAL && j == qt) lim = min(lim, r);
            lim = min(lim, 127);
            const bool full_tile = __all_sync(0xffffffffu, lim >= 127);
            uint32_t v[4][32];
#pragma unroll
            for (int c = 0; c < 4; ++c) tmem_ld_32x32(t_s + c * 32, v[c]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_empty);            // S_j is in registers: the MMA warp may issue S_{j+1} now
            float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            if (!full_tile) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (c * 32 + i > lim) v[c][i] = 0xff800000u;      // -inf
            }
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int i = 0; i < 32; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(v[c][i]));
            const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * scale_log2;
            const bool grow = mx > m_ref + 8.0f;
            const float m_new = grow ? mx : m_ref;
            const float alpha = (grow && j > 0) ? ex2_approx(m_ref - m_new) : 1.0f;
            m_ref = m_new;
            if (j > 0) {
                mbar_wait(pv_done, (uint32_t)(j - 1) & 1u);     // PV_{j-1} has consumed sP and finished updating O
                tc_fence_after();
            }
            float l4[4] = {0.f, 0.f, 0.f, 0.f};
            const float neg_m = -m_ref;
            auto exp_phase = [&](auto use_poly) {
                constexpr bool UP = decltype(use_poly)::value;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
#pragma unroll
                    for (int qd = 0; qd < 4; ++qd) {
                        float e[8];
#pragma unroll
                        for (int t = 0; t < 8; ++t) {
                            const float x = fmaf(__uint_as_float(v[c][8 * qd + t]), scale_log2, neg_m);
                            if (UP && POLY > 0 && (t % (POLY > 0 ? POLY : 1)) == (POLY > 0 ? POLY : 1) - 1) e[t] = ex2_poly(x);
                            else e[t] = ex2_approx(x);                    // exp2(-inf) = 0 for masked columns
                            l4[t & 3] += e[t];
                        }
                        uint4 u;
                        u.x = PACK_P(e[0], e[1]);
                        u.y = PACK_P(e[2], e[3]);
                        u.z = PACK_P(e[4], e[5]);
                        u.w = PACK_P(e[6], e[7]);
                        const int k16 = c * 4 + qd;                       // 16-byte chunk of the 256-byte P row
                        *reinterpret_cast<uint4*>(p_row + (k16 >> 3) * TILE16 + (((k16 & 7) ^ (r & 7)) << 4)) = u;
                    }
                }
            };
            if (POLY > 0 && full_tile) exp_phase(std::true_type{});
            else exp_phase(std::false_type{});
            if (j > 0 && __any_sync(0xffffffffu, grow)) {                 // lazy rescale of O (rare)
#pragma unroll 1
                for (int c = 0; c < HD / 32; ++c) {
                    uint32_t o[32];
                    tmem_ld_32x32(t_o + c * 32, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                    tmem_st_32x32(t_o + c * 32, o);
                }
                tmem_st_wait();
                l_sum *= alpha;
            }
            l_sum += (l4[0] + l4[1]) + (l4[2] + l4[3]);
            tc_fence_before();
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        mbar_wait(pv_done, (uint32_t)(n_kv - 1) & 1u);
        tc_fence_after();
        const int row = q0 + r;
        const float inv = 1.0f / l_sum;
        bf16* orow = O + ((long long)row_base + row) * o_rs + (long long)h * HD;
#pragma unroll 1
        for (int c = 0; c < HD / 32; ++c) {
            uint32_t o[32];
            tmem_ld_32x32(t_o + c * 32, o);
            tmem_ld_wait();
            if (row < S) {
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {
                    uint4 u;
                    u.x = pack_bf16x2(__uint_as_float(o[8 * qd + 0]) * inv, __uint_as_float(o[8 * qd + 1]) * inv);
                    u.y = pack_bf16x2(__uint_as_float(o[8 * qd + 2]) * inv, __uint_as_float(o[8 * qd + 3]) * inv);
                    u.z = pack_bf16x2(__uint_as_float(o[8 * qd + 4]) * inv, __uint_as_float(o[8 * qd + 5]) * inv);
                    u.w = pack_bf16x2(__uint_as_float(o[8 * qd + 6]) * inv, __uint_as_float(o[8 * qd + 7]) * inv);
                    *reinterpret_cast<uint4*>(orow + c * 32 + qd * 8) = u;
                }
            }
        }
        if (LSE && row < S) LSE[((long long)b * Hq + h) * S + row] = (m_ref + log2f(l_sum)) * 0.69314718055994531f;
        tc_fence_before();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc<TMEM_COLS_ATT>(tmem_base);
}

template <int HD, bool CAUSAL, int POLY>
int launch_attn_tc1(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, bf16* o, float* lse, int B, int S, int Hq,
                    int Hkv, long long o_rs, float scale, cudaStream_t st) {
    using C = Att1Cfg<HD>;
    auto kern = attn_tc_fwd1_kernel<HD, CAUSAL, POLY>;
    static bool done = false;
    if (!done) {
        TA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        TA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        done = true;
    }
    dim3 grid((S + BQ - 1) / BQ, Hq, B);
    TA_KERNEL_LAUNCH(kern, grid, ATT1_THREADS, C::SMEM, st, tq, tk, tv, o, lse, S, Hq, Hkv, o_rs, scale * 1.4426950408889634f);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Variant 3 (encoder shape, head_dim 64, non-causal): two threads per query row that never talk inside the loop.
//   Thread `half` of a row owns key columns [64 half, 64 half + 64) of every 128-key tile as an independent KV stream: its own
//   running reference max / sum and its OWN output accumulator in TMEM (O_a at columns [128,192), O_b at [192,256)), i.e. a
//   split-KV (flash-decoding style) decomposition inside the CTA.  The PV product of a tile is issued as two K = 64 groups, one
//   per accumulator.  The two streams are merged once, after the last tile.  Compared with variant 1 there is no per-tile row-max
//   exchange (bar.sync) and compared with variant 2 each thread keeps 64 instead of 128 scores in registers, so 8 softmax warps
//   per CTA x 2 CTAs = 4 warps per SM sub-partition hide the TMEM-load / barrier / MUFU latencies (tools/ubench/sfu2.cu: the
//   softmax instruction mix needs > 2 warps per sub-partition to approach the 16 exp2/clk/SM ceiling).
//   warps 0-7 softmax (q = warp & 3 = TMEM lane quadrant, half = warp >> 2), warp 8 TMA producer + TMEM allocator, warp 9 MMA.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int ATT2_THREADS = 320;

template <int HD>
__global__ void __launch_bounds__(ATT2_THREADS, 2)
attn_tc_fwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, bf16* __restrict__ O, float* __restrict__ LSE, int S, int Hq, int Hkv,
                    long long o_rs, float scale_log2) {
    static_assert(HD == 64, "variant 3 is the encoder shape");
    using C = Att1Cfg<HD>;
    extern __shared__ __align__(1024) uint8_t smem_al[];
    uint8_t* smem = smem_al;
    if (smem_u32(smem) & 1023u) __trap();                 // SWIZZLE_128B tiles need 1024-byte alignment
    uint8_t* sQ = smem;
    uint8_t* sKV = smem + C::TILE_BYTES;
    uint8_t* sP = smem + C::TILE_BYTES * (1 + C::KV_SLOTS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;
    uint64_t* kv_empty = bars + 1 + C::KV_SLOTS;
    uint64_t* s_full = bars + 1 + 2 * C::KV_SLOTS;
    uint64_t* s_empty = s_full + 1;
    uint64_t* p_full = s_full + 2;
    uint64_t* pv_done = s_full + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int hk = h / (Hq / Hkv);
    const int q0 = qt * BQ;
    const int n_kv = (S + BKV - 1) / BKV;
    const int row_base = b * S;
    constexpr uint32_t OA_COL = 128, OB_COL = 192;

    if (warp == 8) {
        if (lane == 0) {
            tma_prefetch_desc(&tmQ);
            tma_prefetch_desc(&tmK);
            tma_prefetch_desc(&tmV);
            mbar_init(q_full, 1);
            for (int s = 0; s < C::KV_SLOTS; ++s) {
                mbar_init(&kv_full[s], 1);
                mbar_init(&kv_empty[s], 1);
            }
            mbar_init(s_full, 1);
            mbar_init(s_empty, 8);
            mbar_init(p_full, 8);
            mbar_init(pv_done, 1);
            mbar_fence_init();
        }
        __syncwarp();
        tmem_alloc<TMEM_COLS_ATT>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 8) {
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, C::TILE_BYTES);
            tma_load_2d(sQ, &tmQ, q_full, h * HD, row_base + q0);
            for (int i = 0; i < 2 * n_kv; ++i) {
                const int slot = i % C::KV_SLOTS;
                const uint32_t ph = (uint32_t)(i / C::KV_SLOTS) & 1u;
                mbar_wait(&kv_empty[slot], ph ^ 1);
                mbar_arrive_expect_tx(&kv_full[slot], C::TILE_BYTES);
                tma_load_2d(sKV + slot * C::TILE_BYTES, (i & 1) ? &tmV : &tmK, &kv_full[slot], hk * HD, row_base + (i >> 1) * BKV);
            }
        }
    } else if (warp == 9) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = umma_idesc_bf16(BQ, BKV);
            constexpr uint32_t idesc_o = umma_idesc_bf16(BQ, HD) | (1u << 16);   // B operand (V) is MN-major
            const uint32_t q_addr = smem_u32(sQ), p_addr = smem_u32(sP);
            auto issue_s = [&](int j) {
                const int i = 2 * j, slot = i % C::KV_SLOTS;
                mbar_wait(&kv_full[slot], (uint32_t)(i / C::KV_SLOTS) & 1u);
                mbar_wait(s_empty, ((uint32_t)j & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t k_addr = smem_u32(sKV + slot * C::TILE_BYTES);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k)
                    umma_f16(tmem_base + S_COL, umma_desc_sw128_kmajor(q_addr + k * 32), umma_desc_sw128_kmajor(k_addr + k * 32), idesc_s,
                             k != 0 ? 1u : 0u);
                umma_commit(s_full);
                umma_commit(&kv_empty[slot]);
            };
            mbar_wait(q_full, 0);
            issue_s(0);
            for (int j = 0; j < n_kv; ++j) {
                if (j + 1 < n_kv) issue_s(j + 1);
                const int i = 2 * j + 1, slot = i % C::KV_SLOTS;
                mbar_wait(p_full, (uint32_t)j & 1u);
                mbar_wait(&kv_full[slot], (uint32_t)(i / C::KV_SLOTS) & 1u);
                tc_fence_after();
                const uint32_t v_addr = smem_u32(sKV + slot * C::TILE_BYTES);
#pragma unroll
                for (int k = 0; k < BKV / 16; ++k) {      // keys [0,64) -> O_a, keys [64,128) -> O_b
                    const uint32_t pa = p_addr + (k >> 2) * TILE16 + (k & 3) * 32;
                    umma_f16(tmem_base + ((k < 4) ? OA_COL : OB_COL), umma_desc_sw128_kmajor(pa),
                             umma_desc_sw128_mnmajor(v_addr + k * 2048, TILE16), idesc_o, (j | (k & 3)) != 0 ? 1u : 0u);
                }
                umma_commit(pv_done);
                umma_commit(&kv_empty[slot]);
            }
        }
    } else {
        const int q = warp & 3, half = warp >> 2;
        const int r = q * 32 + lane;                        // query row of the tile = TMEM lane
        const uint32_t t_s = tmem_base + ((uint32_t)(q * 32) << 16) + S_COL + (uint32_t)(half * 64);
        const uint32_t t_o = tmem_base + ((uint32_t)(q * 32) << 16) + (half ? OB_COL : OA_COL);
        float m_ref = -INFINITY, l_sum = 0.f;
        uint8_t* p_row = sP + half * TILE16 + r * 128;
        for (int j = 0; j < n_kv; ++j) {
            mbar_wait(s_full, (uint32_t)j & 1u);
            tc_fence_after();
            const int lim = min(S - j * BKV - half * 64 - 1, 63);       // my columns e = 0..63 are real keys iff e <= lim
            uint32_t v[2][32];
            tmem_ld_32x32(t_s, v[0]);
            tmem_ld_32x32(t_s + 32, v[1]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_empty);            // S_j is in registers: the MMA warp may issue S_{j+1} now
            if (lim < 63) {                                 // only the last tile(s): warp-uniform
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (c * 32 + i > lim) v[c][i] = 0xff800000u;      // -inf
            }
            float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int i = 0; i < 32; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(v[c][i]));
            const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * scale_log2;
            // a stream whose keys are all masked so far keeps m_ref = -inf; exp2(-inf - (-inf)) must not happen: use 0 as reference
            const bool grow = mx > m_ref + 8.0f;
            const float m_new = grow ? mx : m_ref;
            const float alpha = (grow && m_ref != -INFINITY) ? ex2_approx(m_ref - m_new) : 1.0f;
            m_ref = m_new;
            if (j > 0) {
                mbar_wait(pv_done, (uint32_t)(j - 1) & 1u);     // PV_{j-1} has consumed sP and finished updating O_a / O_b
                tc_fence_after();
            }
            float l4[4] = {0.f, 0.f, 0.f, 0.f};
            const float neg_m = (m_ref == -INFINITY) ? 0.f : -m_ref;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {
                    float e[8];
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        e[t] = ex2_approx(fmaf(__uint_as_float(v[c][8 * qd + t]), scale_log2, neg_m));     // exp2(-inf) = 0 when masked
                        l4[t & 3] += e[t];
                    }
                    uint4 u;
                    u.x = PACK_P(e[0], e[1]);
                    u.y = PACK_P(e[2], e[3]);
                    u.z = PACK_P(e[4], e[5]);
                    u.w = PACK_P(e[6], e[7]);
                    const int k16 = c * 4 + qd;                           // 16-byte chunk of my 128-byte half row
                    *reinterpret_cast<uint4*>(p_row + ((k16 ^ (r & 7)) << 4)) = u;
                }
            }
            if (j > 0 && __any_sync(0xffffffffu, grow)) {                 // lazy rescale of my accumulator (rare)
#pragma unroll 1
                for (int c = 0; c < HD / 32; ++c) {
                    uint32_t o[32];
                    tmem_ld_32x32(t_o + c * 32, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                    tmem_st_32x32(t_o + c * 32, o);
                }
                tmem_st_wait();
                l_sum *= alpha;
            }
            l_sum += (l4[0] + l4[1]) + (l4[2] + l4[3]);
            tc_fence_before();
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        mbar_wait(pv_done, (uint32_t)(n_kv - 1) & 1u);
        tc_fence_after();
        // ---- merge the two KV streams of each row: stream b hands (m, l, O_b) to stream a through shared memory (sP is free now) ----
        float* xO = reinterpret_cast<float*>(sP);               // [64 cols][128 rows] fp32 = 32 KB, column-major: conflict free
        float* xML = reinterpret_cast<float*>(sQ);              // [2][128]  (Q is dead: every S has been issued and completed)
        const uint32_t pair_bar = 1 + q;                        // named barrier per lane quadrant: warps q and q + 4
        if (half == 1) {
#pragma unroll 1
            for (int c = 0; c < HD / 32; ++c) {
                uint32_t o[32];
                tmem_ld_32x32(t_o + c * 32, o);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) xO[(c * 32 + i) * 128 + r] = __uint_as_float(o[i]);
            }
            xML[r] = m_ref;
            xML[128 + r] = l_sum;
            asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
        } else {
            asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
            const float m_b = xML[r], l_b = xML[128 + r];
            const float m = fmaxf(m_ref, m_b);                  // stream a always has real keys (S >= 1), so m is finite
            const float wa = ex2_approx(m_ref - m), wb = (m_b == -INFINITY) ? 0.f : ex2_approx(m_b - m);
            const float l_tot = l_sum * wa + l_b * wb;
            const float inv = 1.0f / l_tot;
            const int row = q0 + r;
            bf16* orow = O + ((long long)row_base + row) * o_rs + (long long)h * HD;
#pragma unroll 1
            for (int c = 0; c < HD / 32; ++c) {
                uint32_t o[32];
                tmem_ld_32x32(t_o + c * 32, o);
                tmem_ld_wait();
                float f[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) f[i] = (__uint_as_float(o[i]) * wa + xO[(c * 32 + i) * 128 + r] * wb) * inv;
                if (row < S) {
#pragma unroll
                    for (int qd = 0; qd < 4; ++qd) {
                        uint4 u;
                        u.x = pack_bf16x2(f[8 * qd + 0], f[8 * qd + 1]);
                        u.y = pack_bf16x2(f[8 * qd + 2], f[8 * qd + 3]);
                        u.z = pack_bf16x2(f[8 * qd + 4], f[8 * qd + 5]);
                        u.w = pack_bf16x2(f[8 * qd + 6], f[8 * qd + 7]);
                        *reinterpret_cast<uint4*>(orow + c * 32 + qd * 8) = u;
                    }
                }
            }
            if (LSE && row < S) LSE[((long long)b * Hq + h) * S + row] = (m + log2f(l_tot)) * 0.69314718055994531f;
        }
        tc_fence_before();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc<TMEM_COLS_ATT>(tmem_base);
}

template <int HD>
int launch_attn_tc2(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, bf16* o, float* lse, int B, int S, int Hq,
                    int Hkv, long long o_rs, float scale, cudaStream_t st) {
    using C = Att1Cfg<HD>;
    auto kern = attn_tc_fwd2_kernel<HD>;
    static bool done = false;
    if (!done) {
        TA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        TA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        done = true;
    }
    dim3 grid((S + BQ - 1) / BQ, Hq, B);
    kern<<<grid, ATT2_THREADS, C::SMEM, st>>>(tq, tk, tv, o, lse, S, Hq, Hkv, o_rs, scale * 1.4426950408889634f);
    TA_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Variant 6 (encoder shape, head_dim 64, non-causal): persistent CTA, two query tiles in ping-pong.
//
// ncu of variant 2 (profiles/r02_c01_ncu_attn_tc_fwd1.txt): the SFU pipe is 61 % busy and the tensor pipe 30 % -- the two resident
// CTAs drift INTO phase (both softmax warps of an SM sub-partition sit in their exp2 phase together, sharing the 4 MUFU lanes, and
// then both wait for their S / PV products together: 25 % of a softmax warp's samples are s_full / pv_done waits), and every CTA
// pays its own prologue (TMEM allocation, Q and first K tile from HBM) 52 times per SM.
//
// Here ONE CTA per SM stays resident and walks work items (batch, head, pair of 128-query tiles):
//   * warps 0-3 = softmax group A (query rows 0..127 of the pair), warps 4-7 = group B (rows 128..255); warp i and warp i+4 sit on
//     the same SM sub-partition and take STRICT TURNS in the exp2 phase (mbarrier token per sub-partition): while A's warp runs its
//     128 MUFU.EX2 per row, B's warp loads its next S row from TMEM, takes the row max, and waits for its PV product -- the SFUs see
//     one warp at a time, back to back;
//   * both groups share every K / V tile (one TMA load, one shared-memory copy for two S and two PV products);
//   * warp 8 = TMA producer (Q double-buffered: the next item's Q and first K/V tiles land while the current item finishes),
//     warp 9 = MMA issuer; TMEM: S_A [0,128) S_B [128,256) O_A [256,320) O_B [320,384);
//   * the O / l epilogue of an item runs in the slack the other group's exp2 phase leaves.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int ATT3_THREADS = 352;      // 8 softmax warps, TMA producer, one MMA-issuing warp per query group

template <bool P_TMEM>
struct Att3Cfg {
    // P_TMEM: the probabilities go back into TENSOR MEMORY (bf16 pairs, 64 columns per group) and the PV product takes its A operand
    // from there (tcgen05.mma with A in TMEM) -- no P store / P operand read through shared memory (64 KB of the 144 KB a tile moves
    // through the 128 B/clk shared-memory port), and the freed 64 KB deepen the K / V ring to 8 tiles.
    static constexpr int KV_SLOTS = P_TMEM ? 8 : 4;
    static constexpr int Q_ITEM = 2 * TILE16;                       // Q of one work item: tile A + tile B (128 x 64 bf16 each)
    static constexpr int KV_OFF = 2 * Q_ITEM;                       // two Q buffers
    static constexpr int P_OFF = KV_OFF + KV_SLOTS * TILE16;
    static constexpr int BAR_OFF = P_OFF + (P_TMEM ? 0 : 2 * (2 * TILE16));   // P_A, P_B: 128 x 128 bf16 each (shared-memory variant)
    static constexpr int SMEM = BAR_OFF + 512;
    static constexpr int TMEM_COLS = 512;
    static constexpr uint32_t S_COL0 = 0, O_COL0 = 256, P_COL0 = 384;          // S_A S_B | O_A O_B | P_A P_B (64 columns each)
};

// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (128 lanes x K, two bf16 per 32-bit column, even k in the low half) comes from
// tensor memory
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

template <int POLY, bool TURNS, bool P_TMEM>
__global__ void __launch_bounds__(ATT3_THREADS, 1)
attn_tc_fwd3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, bf16* __restrict__ O, float* __restrict__ LSE, int S, int Hq, int Hkv,
                    long long o_rs, float scale_log2, int n_items, int n_qpairs, long long* __restrict__ trace, int trace_steps) {
    TA_PDL_ENTRY();
    constexpr int HD = 64;
    using C = Att3Cfg<P_TMEM>;
    // optional timeline of CTA 0 (ta_attn_set_trace): SM clock at the pipeline events of softmax warps 0 (group A) and 4 (group B) and
    // of the MMA thread, 8 slots per (kv tile) step -- tools/attn_trace.py turns it into per-phase durations
    const bool tr_on = trace != nullptr && blockIdx.x == 0;
#define TR(slot, step, ev)                                                                   \
    do {                                                                                     \
        if (tr_on && (threadIdx.x & 31) == 0 && (step) < trace_steps) trace[((slot) * trace_steps + (step)) * 8 + (ev)] = clock64(); \
    } while (0)
    extern __shared__ __align__(1024) uint8_t smem_al[];
    uint8_t* smem = smem_al;
    if (smem_u32(smem) & 1023u) __trap();                 // SWIZZLE_128B tiles need 1024-byte alignment
    uint8_t* sQ = smem;
    uint8_t* sKV = smem + C::KV_OFF;
    uint8_t* sP = smem + C::P_OFF;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
    uint64_t* q_full = bars;                  // [2]
    uint64_t* q_empty = bars + 2;             // [2]
    uint64_t* kv_full = bars + 4;             // [8]
    uint64_t* kv_empty = bars + 12;           // [8]
    uint64_t* s_full = bars + 20;             // [2] per group
    uint64_t* s_empty = bars + 22;            // [2]
    uint64_t* p_full = bars + 24;             // [2]
    uint64_t* pv_done = bars + 26;            // [2]
    uint64_t* o_free = bars + 28;             // [2]
    uint64_t* turn = bars + 30;               // [2][4]: turn[g * 4 + q] completes when the OTHER group's warp q has finished an exp2 phase
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 38);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_kv = (S + BKV - 1) / BKV;
    const int n_mine = (n_items > (int)blockIdx.x) ? (n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int total = n_mine * n_kv;          // (item, kv tile) steps of this CTA

    if (warp == 8) {
        if (lane == 0) {
            tma_prefetch_desc(&tmQ);
            tma_prefetch_desc(&tmK);
            tma_prefetch_desc(&tmV);
            for (int i = 0; i < 2; ++i) {
                mbar_init(&q_full[i], 1);
                mbar_init(&q_empty[i], 2);                 // both groups' MMA threads are done with the item's Q
                mbar_init(&s_full[i], 1);
                mbar_init(&s_empty[i], 4);
                mbar_init(&p_full[i], 4);
                mbar_init(&pv_done[i], 1);
                mbar_init(&o_free[i], 4);
            }
            for (int s = 0; s < C::KV_SLOTS; ++s) {
                mbar_init(&kv_full[s], 1);
                mbar_init(&kv_empty[s], 2);                // both groups have consumed the tile
            }
            for (int i = 0; i < 8; ++i) mbar_init(&turn[i], 1);
            mbar_fence_init();
        }
        __syncwarp();
        tmem_alloc<C::TMEM_COLS>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 8) {
        // ------------------------------------------------ TMA producer ------------------------------------------------
        if (lane == 0) {
            int ld = 0;                                    // K / V tile loads issued so far (ring position)
            for (int it = 0; it < n_mine; ++it) {
                const int w = (int)blockIdx.x + it * (int)gridDim.x;
                const int qp = w % n_qpairs, bh = w / n_qpairs, h = bh % Hq, b = bh / Hq;
                const int hk = h / (Hq / Hkv), row_base = b * S, q0 = qp * 2 * BQ;
                const int buf = it & 1;
                mbar_wait(&q_empty[buf], (((uint32_t)it >> 1) & 1u) ^ 1u);
                mbar_arrive_expect_tx(&q_full[buf], C::Q_ITEM);
                tma_load_2d(sQ + buf * C::Q_ITEM, &tmQ, &q_full[buf], h * HD, row_base + q0);
                tma_load_2d(sQ + buf * C::Q_ITEM + TILE16, &tmQ, &q_full[buf], h * HD, row_base + q0 + BQ);
                for (int i = 0; i < 2 * n_kv; ++i, ++ld) {
                    const int slot = ld % C::KV_SLOTS;
                    mbar_wait(&kv_empty[slot], (((uint32_t)ld / C::KV_SLOTS) & 1u) ^ 1u);
                    mbar_arrive_expect_tx(&kv_full[slot], TILE16);
                    tma_load_2d(sKV + slot * TILE16, (i & 1) ? &tmV : &tmK, &kv_full[slot], hk * HD, row_base + (i >> 1) * BKV);
                }
            }
        }
    } else if (warp >= 9) {
        // ------------------------------------------------ MMA issuers ------------------------------------------------
        // one issuing thread PER GROUP (warp 9: group A, warp 10: group B).  A single in-order thread serving both groups
        // (S_A, PV_A, S_B, PV_B, ...) couples them: the timeline (tools/attn_trace.py) showed PV_A issued 1 700 clk after A's P was
        // ready because the thread was parked on B's barrier, and every softmax warp waiting ~700 clk for an S product that takes
        // 120 clk once issued.
        const int g = __shfl_sync(0xffffffffu, warp - 9, 0);
        if (total > 0) {
            // whole warp, warp-uniform operands, one elected lane issues (see the 64-key kernel below: ~95 clk per MMA issue otherwise)
            constexpr uint32_t idesc_s = umma_idesc_bf16(BQ, BKV);
            constexpr uint32_t idesc_o = umma_idesc_bf16(BQ, HD) | (1u << 16);   // B operand (V) is MN-major
            const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
            const bool leader = elect_one();
            const uint64_t p_desc = umma_desc_sw128_kmajor(smem_u32(sP + g * 2 * TILE16));
            auto issue_s = [&](int t) {                    // S_g(t) = Q_g K(t)^T
                const int it = t / n_kv, j = t - it * n_kv, buf = it & 1;
                const int ld = 2 * t, slot = ld % C::KV_SLOTS;
                if (j == 0) mbar_wait(&q_full[buf], ((uint32_t)it >> 1) & 1u);
                mbar_wait(&kv_full[slot], ((uint32_t)ld / C::KV_SLOTS) & 1u);
                mbar_wait(&s_empty[g], ((uint32_t)t & 1u) ^ 1u);
                tc_fence_after();
                const uint64_t q_desc = umma_desc_sw128_kmajor(smem_u32(sQ + buf * C::Q_ITEM + g * TILE16));
                const uint64_t k_desc = umma_desc_sw128_kmajor(smem_u32(sKV + slot * TILE16));
                if (leader) {
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k)
                        umma_f16(tb + C::S_COL0 + g * 128, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
                    umma_commit(&s_full[g]);
                    umma_commit(&kv_empty[slot]);
                    if (j == n_kv - 1) umma_commit(&q_empty[buf]);
                }
                __syncwarp();
            };
            auto issue_pv = [&](int t) {                   // O_g (+)= P_g(t) V(t)
                const int it = t / n_kv, j = t - it * n_kv;
                const int ld = 2 * t + 1, slot = ld % C::KV_SLOTS;
                mbar_wait(&p_full[g], (uint32_t)t & 1u);
                TR(8 + g, t, 3);
                mbar_wait(&kv_full[slot], ((uint32_t)ld / C::KV_SLOTS) & 1u);
                if (j == 0 && it > 0) mbar_wait(&o_free[g], ((uint32_t)(it - 1)) & 1u);   // the previous item's O has been read out
                TR(8 + g, t, 4);
                tc_fence_after();
                const uint64_t v_desc = umma_desc_sw128_mnmajor(smem_u32(sKV + slot * TILE16), TILE16);
                if (leader) {
                    if constexpr (P_TMEM) {
#pragma unroll
                        for (int k = 0; k < BKV / 16; ++k)      // 16 keys = 8 columns of packed bf16 pairs
                            umma_f16_ts(tb + C::O_COL0 + g * HD, tb + C::P_COL0 + g * 64 + k * 8, v_desc + 128 * k, idesc_o,
                                        (j | k) != 0 ? 1u : 0u);
                    } else {
#pragma unroll
                        for (int k = 0; k < BKV / 16; ++k)      // P tile = two 64-key SWIZZLE_128B sub-tiles, 32 bytes per 16 keys inside each
                            umma_f16(tb + C::O_COL0 + g * HD, p_desc + (k >> 2) * (TILE16 >> 4) + (k & 3) * 2, v_desc + 128 * k, idesc_o,
                                     (j | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&pv_done[g]);
                    umma_commit(&kv_empty[slot]);
                }
                __syncwarp();
            };
            issue_s(0);
            for (int t = 0; t < total; ++t) {
                TR(8 + g, t, 0);
                if (t + 1 < total) issue_s(t + 1);
                TR(8 + g, t, 1);
                issue_pv(t);
                TR(8 + g, t, 2);
            }
        }
    } else {
        // ------------------------------------------------ softmax groups ------------------------------------------------
        const int g = warp >> 2, q = warp & 3;
        const int r = q * 32 + lane;                        // query row of my tile = TMEM lane
        const uint32_t t_s = tmem_base + ((uint32_t)(q * 32) << 16) + C::S_COL0 + (uint32_t)g * 128;
        const uint32_t t_o = tmem_base + ((uint32_t)(q * 32) << 16) + C::O_COL0 + (uint32_t)g * HD;
        uint8_t* p_row = sP + (P_TMEM ? 0 : g * 2 * TILE16 + r * 128);
        const uint32_t t_p = tmem_base + ((uint32_t)(q * 32) << 16) + C::P_COL0 + (uint32_t)g * 64;
        uint64_t* my_turn = &turn[g * 4 + q];
        uint64_t* other_turn = &turn[(g ^ 1) * 4 + q];
        int t = 0;
        for (int it = 0; it < n_mine; ++it) {
            const int w = (int)blockIdx.x + it * (int)gridDim.x;
            const int qp = w % n_qpairs, bh = w / n_qpairs, h = bh % Hq, b = bh / Hq;
            const int row_base = b * S, q0 = qp * 2 * BQ + g * BQ;
            float m_ref = -INFINITY, l_sum = 0.f;
            for (int j = 0; j < n_kv; ++j, ++t) {
                TR(warp, t, 0);
                mbar_wait(&s_full[g], (uint32_t)t & 1u);
                tc_fence_after();
                TR(warp, t, 1);
                const int lim = min(S - j * BKV - 1, 127);  // my columns e = 0..127 are real (unmasked) keys iff e <= lim
                uint32_t v[4][32];
#pragma unroll
                for (int c = 0; c < 4; ++c) tmem_ld_32x32(t_s + c * 32, v[c]);
                tmem_ld_wait();
                TR(warp, t, 2);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_empty[g]);    // S is in registers: the MMA warp may issue my next S product now
                if (lim < 127) {                            // last tile of the sequence only: warp-uniform
#pragma unroll
                    for (int c = 0; c < 4; ++c)
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (c * 32 + i > lim) v[c][i] = 0xff800000u;      // -inf
                }
                float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int i = 0; i < 32; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(v[c][i]));
                const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * scale_log2;
                const bool grow = mx > m_ref + 8.0f;
                const float m_new = grow ? mx : m_ref;
                const float alpha = (grow && j > 0) ? ex2_approx(m_ref - m_new) : 1.0f;
                m_ref = m_new;
                TR(warp, t, 3);
                if (j > 0) {
                    mbar_wait(&pv_done[g], (uint32_t)(t - 1) & 1u);     // PV of my previous tile has consumed P and updated O
                    tc_fence_after();
                }
                TR(warp, t, 4);
                // ---- TURNS: my turn on this sub-partition's SFUs (group A starts; then strictly alternating with warp q of the other group) ----
                if constexpr (TURNS) mbar_wait(my_turn, g == 0 ? (((uint32_t)t & 1u) ^ 1u) : ((uint32_t)t & 1u));
                TR(warp, t, 5);
                float l4[4] = {0.f, 0.f, 0.f, 0.f};
                const float neg_m = -m_ref;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
#pragma unroll
                    for (int qd = 0; qd < 4; ++qd) {
                        float e[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const float x = fmaf(__uint_as_float(v[c][8 * qd + u]), scale_log2, neg_m);
                            if (POLY > 0 && (u % (POLY > 0 ? POLY : 1)) == (POLY > 0 ? POLY : 1) - 1) e[u] = ex2_poly(x);
                            else e[u] = ex2_approx(x);                    // exp2(-inf) = 0 for masked columns
                            l4[u & 3] += e[u];
                        }
                        uint4 pk;
                        pk.x = PACK_P(e[0], e[1]);
                        pk.y = PACK_P(e[2], e[3]);
                        pk.z = PACK_P(e[4], e[5]);
                        pk.w = PACK_P(e[6], e[7]);
                        const int k16 = c * 4 + qd;                       // 16-byte chunk of the 256-byte P row
                        if constexpr (P_TMEM) {
                            // keep the packed pairs in the (dead) S registers: columns [4 k16, +4) of my P row, stored 8 columns at a time
                            v[c][8 * qd + 0] = pk.x; v[c][8 * qd + 1] = pk.y; v[c][8 * qd + 2] = pk.z; v[c][8 * qd + 3] = pk.w;
                            if (qd & 1) {
                                const uint32_t w8[8] = {v[c][8 * (qd - 1) + 0], v[c][8 * (qd - 1) + 1], v[c][8 * (qd - 1) + 2], v[c][8 * (qd - 1) + 3],
                                                        pk.x, pk.y, pk.z, pk.w};
                                tmem_st_32x8(t_p + (k16 - 1) * 4, w8);
                            }
                        } else {
                            *reinterpret_cast<uint4*>(p_row + (k16 >> 3) * TILE16 + (((k16 & 7) ^ (r & 7)) << 4)) = pk;
                        }
                    }
                }
                if constexpr (TURNS) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(other_turn);               // hand the SFUs to the other group's warp
                }
                if (j > 0 && __any_sync(0xffffffffu, grow)) {             // lazy rescale of O (rare)
#pragma unroll 1
                    for (int c = 0; c < HD / 32; ++c) {
                        uint32_t o[32];
                        tmem_ld_32x32(t_o + c * 32, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        tmem_st_32x32(t_o + c * 32, o);
                    }
                    tmem_st_wait();
                    l_sum *= alpha;
                }
                l_sum += (l4[0] + l4[1]) + (l4[2] + l4[3]);
                if constexpr (P_TMEM) tmem_st_wait();
                TR(warp, t, 6);
                tc_fence_before();
                if constexpr (!P_TMEM) fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[g]);
            }
            // ---- item epilogue: O / l -> bf16 (runs while the other group has the SFUs) ----
            mbar_wait(&pv_done[g], (uint32_t)(t - 1) & 1u);
            tc_fence_after();
            const int row = q0 + r;
            const float inv = 1.0f / l_sum;
            bf16* orow = O + ((long long)row_base + row) * o_rs + (long long)h * HD;
#pragma unroll 1
            for (int c = 0; c < HD / 32; ++c) {
                uint32_t o[32];
                tmem_ld_32x32(t_o + c * 32, o);
                tmem_ld_wait();
                if (row < S) {
#pragma unroll
                    for (int qd = 0; qd < 4; ++qd) {
                        uint4 u;
                        u.x = pack_bf16x2(__uint_as_float(o[8 * qd + 0]) * inv, __uint_as_float(o[8 * qd + 1]) * inv);
                        u.y = pack_bf16x2(__uint_as_float(o[8 * qd + 2]) * inv, __uint_as_float(o[8 * qd + 3]) * inv);
                        u.z = pack_bf16x2(__uint_as_float(o[8 * qd + 4]) * inv, __uint_as_float(o[8 * qd + 5]) * inv);
                        u.w = pack_bf16x2(__uint_as_float(o[8 * qd + 6]) * inv, __uint_as_float(o[8 * qd + 7]) * inv);
                        *reinterpret_cast<uint4*>(orow + c * 32 + qd * 8) = u;
                    }
                }
            }
            if (LSE && row < S) LSE[((long long)b * Hq + h) * S + row] = (m_ref + log2f(l_sum)) * 0.69314718055994531f;
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&o_free[g]);                       // the MMA warp may overwrite O with the next item's first PV
        }
    }

#undef TR
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

long long* g_attn_trace = nullptr;
int g_attn_trace_steps = 0;

template <int POLY, bool TURNS, bool P_TMEM>
int launch_attn_tc3(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, bf16* o, float* lse, int B, int S, int Hq,
                    int Hkv, long long o_rs, float scale, cudaStream_t st) {
    using C = Att3Cfg<P_TMEM>;
    auto kern = attn_tc_fwd3_kernel<POLY, TURNS, P_TMEM>;
    static bool done = false;
    static int n_sm = 0;
    if (!done) {
        TA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        int dev = 0;
        TA_CHECK_CUDA(cudaGetDevice(&dev));
        TA_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        done = true;
    }
    const int n_qpairs = ((S + BQ - 1) / BQ + 1) / 2;
    const long long n_items = (long long)B * Hq * n_qpairs;
    TA_REQUIRE(n_items < (1LL << 30), "attention: too many work items");
    const int grid = (int)((n_items < n_sm) ? n_items : n_sm);
    TA_KERNEL_LAUNCH(kern, grid, ATT3_THREADS, C::SMEM, st, tq, tk, tv, o, lse, S, Hq, Hkv, o_rs, scale * 1.4426950408889634f, (int)n_items,
                     n_qpairs, g_attn_trace, g_attn_trace_steps);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Variant "64-key tiles, three CTAs per SM" (encoder shape, head_dim 64, non-causal).
//
// What the measurements of this round say (tools/ubench/sfu4.cu, tools/attn_trace.py, ncu):
//   * MUFU.EX2 issues once per 8.0 clk and SM sub-partition = 16 exp2 / clk / SM, and that is the kernel's floor: 1024 clk per
//     128 x 128 score tile;  cvt.rn.bf16x2 is free next to it;
//   * one softmax warp alone runs its exp2 phase at ~11.7 clk per EX2, two warps of a sub-partition together at ~8.8 (91 % of the
//     pipe) -- but every warp also spends ~1 100 clk per tile NOT issuing exp2 (TMEM load, row max, waiting for its S / PV products
//     and for the barrier wake-ups), and with two softmax warps per sub-partition nothing fills the SFUs meanwhile: 8.2 of 16
//     exp2 / clk / SM, whatever the arrangement (two CTAs, one persistent CTA with two query tiles, strict turns, P in TMEM).
// So: THREE softmax warps per sub-partition.  Tensor memory (512 columns) and registers (64 K) do not allow a third 128-column S
// tile, hence 64-key tiles: S 64 + O 64 columns and 64 fp32 scores per thread -> 112 registers, 64 KB of shared memory
// (Q 16 K, four 8 KB K / V tiles, P 16 K) and 128 TMEM columns per CTA, three CTAs per SM.
//   warps 0-3 softmax (thread = query row = TMEM lane), warp 4 TMA producer + TMEM allocator, warp 5 MMA issuer.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int BKV5 = 64;
constexpr int TILE8 = BKV5 * 64 * 2;       // one K or V tile: 64 keys x 64 bf16
struct Att5Cfg {
    static constexpr int KV_SLOTS = 4;
    static constexpr int KV_OFF = TILE16;                            // after Q (128 x 64 bf16)
    static constexpr int P_OFF = KV_OFF + KV_SLOTS * TILE8;
    static constexpr int BAR_OFF = P_OFF + TILE16;                   // P: 128 x 64 bf16 = one SWIZZLE_128B tile
    static constexpr int SMEM = BAR_OFF + 128;
    static constexpr int TMEM_COLS = 128;
};

template <int POLY>
__global__ void __launch_bounds__(ATT1_THREADS, 3)
attn_tc_fwd5_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, bf16* __restrict__ O, float* __restrict__ LSE, int S, int Hq, int Hkv,
                    long long o_rs, float scale_log2) {
    TA_PDL_ENTRY();
    constexpr int HD = 64;
    using C = Att5Cfg;
    extern __shared__ __align__(1024) uint8_t smem_al[];
    uint8_t* smem = smem_al;
    if (smem_u32(smem) & 1023u) __trap();                 // SWIZZLE_128B tiles need 1024-byte alignment
    uint8_t* sQ = smem;
    uint8_t* sKV = smem + C::KV_OFF;
    uint8_t* sP = smem + C::P_OFF;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;
    uint64_t* kv_empty = bars + 1 + C::KV_SLOTS;
    uint64_t* s_full = bars + 1 + 2 * C::KV_SLOTS;
    uint64_t* s_empty = s_full + 1;
    uint64_t* p_full = s_full + 2;
    uint64_t* pv_done = s_full + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int hk = h / (Hq / Hkv);
    const int q0 = qt * BQ;
    const int n_kv = (S + BKV5 - 1) / BKV5;
    const int row_base = b * S;
    constexpr uint32_t S5_COL = 0, O5_COL = 64;

    if (warp == 4) {
        if (lane == 0) {
            tma_prefetch_desc(&tmQ);
            tma_prefetch_desc(&tmK);
            tma_prefetch_desc(&tmV);
            mbar_init(q_full, 1);
            for (int s = 0; s < C::KV_SLOTS; ++s) {
                mbar_init(&kv_full[s], 1);
                mbar_init(&kv_empty[s], 1);
            }
            mbar_init(s_full, 1);
            mbar_init(s_empty, 4);
            mbar_init(p_full, 4);
            mbar_init(pv_done, 1);
            mbar_fence_init();
        }
        __syncwarp();
        tmem_alloc<C::TMEM_COLS>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, TILE16);
            tma_load_2d(sQ, &tmQ, q_full, h * HD, row_base + q0);
            for (int i = 0; i < 2 * n_kv; ++i) {
                const int slot = i % C::KV_SLOTS;
                mbar_wait(&kv_empty[slot], (((uint32_t)i / C::KV_SLOTS) & 1u) ^ 1u);
                mbar_arrive_expect_tx(&kv_full[slot], TILE8);
                tma_load_2d(sKV + slot * TILE8, (i & 1) ? &tmV : &tmK, &kv_full[slot], hk * HD, row_base + (i >> 1) * BKV5);
            }
        }
    } else if (warp == 5) {
        // The WHOLE warp walks the issue loop (warp-uniform control flow and operands); one elected lane executes the tcgen05
        // instructions.  Inside an `if (lane == 0)` region ptxas cannot keep the descriptors in uniform registers and wraps every
        // UTCHMMA into an ELECT / 4 x R2UR.BROADCAST / BRA.U.ANY loop: the kernel timeline (tools/attn_trace.py) showed ~95 clk per
        // MMA *issue* -- 760 clk per key tile on the softmax -> PV -> softmax critical path, while an M128 N64 K16 MMA executes in 32.
        constexpr uint32_t idesc_s = umma_idesc_bf16(BQ, BKV5);
        constexpr uint32_t idesc_o = umma_idesc_bf16(BQ, HD) | (1u << 16);   // B operand (V) is MN-major
        const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
        const bool leader = elect_one();
        const uint64_t q_desc = umma_desc_sw128_kmajor(smem_u32(sQ)), p_desc = umma_desc_sw128_kmajor(smem_u32(sP));
        auto issue_s = [&](int j) {
            const int i = 2 * j, slot = i % C::KV_SLOTS;
            mbar_wait(&kv_full[slot], ((uint32_t)i / C::KV_SLOTS) & 1u);
            mbar_wait(s_empty, ((uint32_t)j & 1u) ^ 1u);
            tc_fence_after();
            const uint64_t k_desc = umma_desc_sw128_kmajor(smem_u32(sKV + slot * TILE8));
            if (leader) {
#pragma unroll
                for (int k = 0; k < HD / 16; ++k)      // + 32 bytes (>> 4 = 2) per 16-wide k step
                    umma_f16(tb + S5_COL, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
                umma_commit(s_full);
                umma_commit(&kv_empty[slot]);
            }
            __syncwarp();
        };
        mbar_wait(q_full, 0);
        issue_s(0);
        for (int j = 0; j < n_kv; ++j) {
            if (j + 1 < n_kv) issue_s(j + 1);
            const int i = 2 * j + 1, slot = i % C::KV_SLOTS;
            mbar_wait(p_full, (uint32_t)j & 1u);
            mbar_wait(&kv_full[slot], ((uint32_t)i / C::KV_SLOTS) & 1u);
            tc_fence_after();
            const uint64_t v_desc = umma_desc_sw128_mnmajor(smem_u32(sKV + slot * TILE8), TILE8);
            if (leader) {
#pragma unroll
                for (int k = 0; k < BKV5 / 16; ++k)    // P: + 32 bytes per 16 keys; V: + 16 rows x 128 B = 2048 bytes (>> 4 = 128)
                    umma_f16(tb + O5_COL, p_desc + 2 * k, v_desc + 128 * k, idesc_o, (j | k) != 0 ? 1u : 0u);
                umma_commit(pv_done);
                umma_commit(&kv_empty[slot]);
            }
            __syncwarp();
        }
    } else {
        const int r = warp * 32 + lane;                     // query row of the tile = TMEM lane
        const uint32_t t_s = tmem_base + ((uint32_t)(warp * 32) << 16) + S5_COL;
        const uint32_t t_o = tmem_base + ((uint32_t)(warp * 32) << 16) + O5_COL;
        float m_ref = -INFINITY, l_sum = 0.f;
        uint8_t* p_row = sP + r * 128;
        for (int j = 0; j < n_kv; ++j) {
            mbar_wait(s_full, (uint32_t)j & 1u);
            tc_fence_after();
            const int lim = min(S - j * BKV5 - 1, BKV5 - 1);      // my columns e = 0..63 are real (unmasked) keys iff e <= lim
            uint32_t v[2][32];
            tmem_ld_32x32(t_s, v[0]);
            tmem_ld_32x32(t_s + 32, v[1]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_empty);            // S_j is in registers: the MMA warp may issue S_{j+1} now
            if (lim < BKV5 - 1) {                           // last tile only: warp-uniform
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (c * 32 + i > lim) v[c][i] = 0xff800000u;      // -inf
            }
            float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int i = 0; i < 32; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(v[c][i]));
            const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * scale_log2;
            const bool grow = mx > m_ref + 8.0f;
            const float m_new = grow ? mx : m_ref;
            const float alpha = (grow && j > 0) ? ex2_approx(m_ref - m_new) : 1.0f;
            m_ref = m_new;
            if (j > 0) {
                mbar_wait(pv_done, (uint32_t)(j - 1) & 1u);     // PV_{j-1} has consumed sP and finished updating O
                tc_fence_after();
            }
            float l4[4] = {0.f, 0.f, 0.f, 0.f};
            const float neg_m = -m_ref;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {
                    float e[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const float x = fmaf(__uint_as_float(v[c][8 * qd + u]), scale_log2, neg_m);
                        if (POLY > 0 && (u % (POLY > 0 ? POLY : 1)) == (POLY > 0 ? POLY : 1) - 1) e[u] = ex2_poly(x);
                        else e[u] = ex2_approx(x);                    // exp2(-inf) = 0 for masked columns
                        l4[u & 3] += e[u];
                    }
                    uint4 pk;
                    pk.x = PACK_P(e[0], e[1]);
                    pk.y = PACK_P(e[2], e[3]);
                    pk.z = PACK_P(e[4], e[5]);
                    pk.w = PACK_P(e[6], e[7]);
                    const int k16 = c * 4 + qd;                       // 16-byte chunk of the 128-byte P row
                    *reinterpret_cast<uint4*>(p_row + ((k16 ^ (r & 7)) << 4)) = pk;
                }
            }
            if (j > 0 && __any_sync(0xffffffffu, grow)) {             // lazy rescale of O (rare)
#pragma unroll 1
                for (int c = 0; c < HD / 32; ++c) {
                    uint32_t o[32];
                    tmem_ld_32x32(t_o + c * 32, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                    tmem_st_32x32(t_o + c * 32, o);
                }
                tmem_st_wait();
                l_sum *= alpha;
            }
            l_sum += (l4[0] + l4[1]) + (l4[2] + l4[3]);
            tc_fence_before();
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        mbar_wait(pv_done, (uint32_t)(n_kv - 1) & 1u);
        tc_fence_after();
        const int row = q0 + r;
        const float inv = 1.0f / l_sum;
        bf16* orow = O + ((long long)row_base + row) * o_rs + (long long)h * HD;
#pragma unroll 1
        for (int c = 0; c < HD / 32; ++c) {
            uint32_t o[32];
            tmem_ld_32x32(t_o + c * 32, o);
            tmem_ld_wait();
            if (row < S) {
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {
                    uint4 u;
                    u.x = pack_bf16x2(__uint_as_float(o[8 * qd + 0]) * inv, __uint_as_float(o[8 * qd + 1]) * inv);
                    u.y = pack_bf16x2(__uint_as_float(o[8 * qd + 2]) * inv, __uint_as_float(o[8 * qd + 3]) * inv);
                    u.z = pack_bf16x2(__uint_as_float(o[8 * qd + 4]) * inv, __uint_as_float(o[8 * qd + 5]) * inv);
                    u.w = pack_bf16x2(__uint_as_float(o[8 * qd + 6]) * inv, __uint_as_float(o[8 * qd + 7]) * inv);
                    *reinterpret_cast<uint4*>(orow + c * 32 + qd * 8) = u;
                }
            }
        }
        if (LSE && row < S) LSE[((long long)b * Hq + h) * S + row] = (m_ref + log2f(l_sum)) * 0.69314718055994531f;
        tc_fence_before();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

template <int POLY>
int launch_attn_tc5(const bf16* q, const bf16* k, const bf16* v, bf16* o, float* lse, int B, int S, int Hq, int Hkv, long long q_rs,
                    long long k_rs, long long v_rs, long long o_rs, float scale, cudaStream_t st) {
    using C = Att5Cfg;
    auto kern = attn_tc_fwd5_kernel<POLY>;
    static bool done = false;
    if (!done) {
        TA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        TA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        done = true;
    }
    CUtensorMap tq, tk, tv;                                  // K / V boxes hold 64 rows here
    const long long rows = (long long)B * S;
    int rc = k_make_tensor_map_2d(&tq, q, rows, (long long)Hq * 64, q_rs, BQ);
    if (rc) return rc;
    rc = k_make_tensor_map_2d(&tk, k, rows, (long long)Hkv * 64, k_rs, BKV5);
    if (rc) return rc;
    rc = k_make_tensor_map_2d(&tv, v, rows, (long long)Hkv * 64, v_rs, BKV5);
    if (rc) return rc;
    dim3 grid((S + BQ - 1) / BQ, Hq, B);
    TA_KERNEL_LAUNCH(kern, grid, ATT1_THREADS, C::SMEM, st, tq, tk, tv, o, lse, S, Hq, Hkv, o_rs, scale * 1.4426950408889634f);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Variant 6: variant 5's structure (64-key tiles, one thread per query row, warps 0-3 softmax, warp 4 TMA + TMEM, warp 5 MMA issuer)
// for head_dim 128, causal, GQA, left-padded keys (kv_start) -- the decoder's attention.  Q 32 KB + 4 x 16 KB K/V ring + P 16 KB =
// 112 KB, TMEM 64 (S) + 128 (O) columns: TWO CTAs per SM.  The decoder's problems are short (S = 464: a query tile sees 2-8 key tiles),
// so a CTA's prologue / first-load latency / epilogue are as long as its main loop; with one 192 KB CTA per SM (attn_tc_fwd_kernel<128>)
// nothing ran under them and ncu showed no unit above 24 % (profiles/r02_c23_ncu_top_kernels_summary.txt: 115.6 us, tensor pipe 19 %).
// Heavy query tiles are scheduled first (linear grid ordered by query tile, last tile first).
// ---------------------------------------------------------------------------------------------------------------------
template <int HD, int SLOTS>
struct Att6Cfg {
    static constexpr int NSUB = HD / 64;
    static constexpr int Q_BYTES = NSUB * TILE16;                    // 128 rows x HD
    static constexpr int KV_BYTES = NSUB * TILE8;                    // 64 keys x HD
    static constexpr int KV_SLOTS = SLOTS;
    static constexpr int KV_OFF = Q_BYTES;
    static constexpr int P_OFF = KV_OFF + KV_SLOTS * KV_BYTES;
    static constexpr int BAR_OFF = P_OFF + TILE16;
    static constexpr int SMEM = BAR_OFF + (SLOTS <= 4 ? 128 : 256);  // 5 + 2 SLOTS barriers + the TMEM slot
    static constexpr int TMEM_COLS = (HD == 64) ? 128 : 256;         // S: 64 columns, O: HD columns
    static constexpr int CTAS = (HD == 64) ? 3 : (SLOTS <= 4 ? 2 : 1);
};

template <int HD, bool CAUSAL, int SLOTS>
__global__ void __launch_bounds__(ATT1_THREADS, Att6Cfg<HD, SLOTS>::CTAS)
attn_tc_fwd6_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, bf16* __restrict__ O, float* __restrict__ LSE, int S, int Hq, int Hkv,
                    long long o_rs, float scale_log2, const int* __restrict__ kv_start) {
    TA_PDL_ENTRY();
    using C = Att6Cfg<HD, SLOTS>;
    extern __shared__ __align__(1024) uint8_t smem_al[];
    uint8_t* smem = smem_al;
    if (smem_u32(smem) & 1023u) __trap();                 // SWIZZLE_128B tiles need 1024-byte alignment
    uint8_t* sQ = smem;
    uint8_t* sKV = smem + C::KV_OFF;
    uint8_t* sP = smem + C::P_OFF;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;
    uint64_t* kv_empty = bars + 1 + C::KV_SLOTS;
    uint64_t* s_full = bars + 1 + 2 * C::KV_SLOTS;
    uint64_t* s_empty = s_full + 1;
    uint64_t* p_full = s_full + 2;
    uint64_t* pv_done = s_full + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // linear grid, longest work first (causal: the last query tile sees every key tile)
    const int n_qt = (S + BQ - 1) / BQ, per_qt = (int)gridDim.x / n_qt;      // per_qt = Hq * B
    const int qt = CAUSAL ? n_qt - 1 - (int)blockIdx.x / per_qt : (int)blockIdx.x / per_qt;
    const int h = ((int)blockIdx.x % per_qt) % Hq, b = ((int)blockIdx.x % per_qt) / Hq;
    const int hk = h / (Hq / Hkv);
    const int q0 = qt * BQ;
    const int n_all = (S + BKV5 - 1) / BKV5;
    const int n_kv = CAUSAL ? min(n_all, 2 * qt + 2) : n_all;       // keys <= q0 + 127
    const int row_base = b * S;
    constexpr uint32_t S6_COL = 0, O6_COL = 64;

    if (warp == 4) {
        if (lane == 0) {
            tma_prefetch_desc(&tmQ);
            tma_prefetch_desc(&tmK);
            tma_prefetch_desc(&tmV);
            mbar_init(q_full, 1);
            for (int s = 0; s < C::KV_SLOTS; ++s) {
                mbar_init(&kv_full[s], 1);
                mbar_init(&kv_empty[s], 1);
            }
            mbar_init(s_full, 1);
            mbar_init(s_empty, 4);
            mbar_init(p_full, 4);
            mbar_init(pv_done, 1);
            mbar_fence_init();
        }
        __syncwarp();
        tmem_alloc<C::TMEM_COLS>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, C::Q_BYTES);
#pragma unroll
            for (int u = 0; u < C::NSUB; ++u) tma_load_2d(sQ + u * TILE16, &tmQ, q_full, h * HD + u * 64, row_base + q0);
            for (int i = 0; i < 2 * n_kv; ++i) {
                const int slot = i % C::KV_SLOTS;
                mbar_wait(&kv_empty[slot], (((uint32_t)i / C::KV_SLOTS) & 1u) ^ 1u);
                mbar_arrive_expect_tx(&kv_full[slot], C::KV_BYTES);
#pragma unroll
                for (int u = 0; u < C::NSUB; ++u)
                    tma_load_2d(sKV + slot * C::KV_BYTES + u * TILE8, (i & 1) ? &tmV : &tmK, &kv_full[slot], hk * HD + u * 64,
                                row_base + (i >> 1) * BKV5);
            }
        }
    } else if (warp == 5) {
        // whole warp, warp-uniform operands, one elected lane issues (see variant 5)
        constexpr uint32_t idesc_s = umma_idesc_bf16(BQ, BKV5);
        constexpr uint32_t idesc_o = umma_idesc_bf16(BQ, HD) | (1u << 16);   // B operand (V) is MN-major
        const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
        const bool leader = elect_one();
        const uint32_t q_addr = smem_u32(sQ);
        const uint64_t p_desc = umma_desc_sw128_kmajor(smem_u32(sP));
        auto issue_s = [&](int j) {
            const int i = 2 * j, slot = i % C::KV_SLOTS;
            mbar_wait(&kv_full[slot], ((uint32_t)i / C::KV_SLOTS) & 1u);
            mbar_wait(s_empty, ((uint32_t)j & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t k_addr = smem_u32(sKV + slot * C::KV_BYTES);
            if (leader) {
#pragma unroll
                for (int k = 0; k < HD / 16; ++k)      // 64-wide sub-tile (k >> 2), + 32 bytes per 16-wide k step inside it
                    umma_f16(tb + S6_COL, umma_desc_sw128_kmajor(q_addr + (k >> 2) * TILE16 + (k & 3) * 32),
                             umma_desc_sw128_kmajor(k_addr + (k >> 2) * TILE8 + (k & 3) * 32), idesc_s, k != 0 ? 1u : 0u);
                umma_commit(s_full);
                umma_commit(&kv_empty[slot]);
            }
            __syncwarp();
        };
        mbar_wait(q_full, 0);
        issue_s(0);
        for (int j = 0; j < n_kv; ++j) {
            if (j + 1 < n_kv) issue_s(j + 1);
            const int i = 2 * j + 1, slot = i % C::KV_SLOTS;
            mbar_wait(p_full, (uint32_t)j & 1u);
            mbar_wait(&kv_full[slot], ((uint32_t)i / C::KV_SLOTS) & 1u);
            tc_fence_after();
            const uint32_t v_addr = smem_u32(sKV + slot * C::KV_BYTES);
            if (leader) {
#pragma unroll
                for (int k = 0; k < BKV5 / 16; ++k)    // P: + 32 bytes per 16 keys; V: + 16 rows x 128 B; 64-wide head-dim atoms TILE8 apart
                    umma_f16(tb + O6_COL, p_desc + 2 * k, umma_desc_sw128_mnmajor(v_addr + k * 2048, TILE8), idesc_o, (j | k) != 0 ? 1u : 0u);
                umma_commit(pv_done);
                umma_commit(&kv_empty[slot]);
            }
            __syncwarp();
        }
    } else {
        const int r = warp * 32 + lane;                     // query row of the tile = TMEM lane
        const uint32_t t_s = tmem_base + ((uint32_t)(warp * 32) << 16) + S6_COL;
        const uint32_t t_o = tmem_base + ((uint32_t)(warp * 32) << 16) + O6_COL;
        float m_ref = -INFINITY, l_sum = 0.f;
        uint8_t* p_row = sP + r * 128;
        // left-padded prompts: keys before kv_start[b] are padding.  A real query row (>= kv_start) must not see them; a padding row
        // keeps plain causal attention so that its (unused) output stays finite.
        const int k_lo = (kv_start != nullptr && q0 + r >= kv_start[b]) ? kv_start[b] : 0;
        for (int j = 0; j < n_kv; ++j) {
            mbar_wait(s_full, (uint32_t)j & 1u);
            tc_fence_after();
            // my columns e = lo..lim of this tile are real (unmasked) keys
            int lim = S - j * BKV5 - 1;
            if (CAUSAL) lim = min(lim, q0 + r - j * BKV5);
            lim = min(lim, BKV5 - 1);
            const int lo = k_lo - j * BKV5;
            const bool full_tile = __all_sync(0xffffffffu, lim >= BKV5 - 1 && lo <= 0);
            uint32_t v[2][32];
            tmem_ld_32x32(t_s, v[0]);
            tmem_ld_32x32(t_s + 32, v[1]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_empty);            // S_j is in registers: the MMA warp may issue S_{j+1} now
            if (!full_tile) {
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (c * 32 + i > lim || c * 32 + i < lo) v[c][i] = 0xff800000u;      // -inf
            }
            float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int i = 0; i < 32; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(v[c][i]));
            const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * scale_log2;
            const bool grow = mx > m_ref + 8.0f;
            const float m_new = grow ? mx : m_ref;
            const float alpha = (grow && j > 0) ? ex2_approx(m_ref - m_new) : 1.0f;   // m_ref = -inf (all keys masked so far): 0, O and l are 0
            m_ref = m_new;
            if (j > 0) {
                mbar_wait(pv_done, (uint32_t)(j - 1) & 1u);     // PV_{j-1} has consumed sP and finished updating O
                tc_fence_after();
            }
            float l4[4] = {0.f, 0.f, 0.f, 0.f};
            const float neg_m = (m_ref == -INFINITY) ? 0.f : -m_ref;     // a tile of masked keys only: exp2(-inf) = 0, never inf - inf
#pragma unroll
            for (int c = 0; c < 2; ++c) {
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {
                    float e[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        e[u] = ex2_approx(fmaf(__uint_as_float(v[c][8 * qd + u]), scale_log2, neg_m));     // exp2(-inf) = 0 for masked columns
                        l4[u & 3] += e[u];
                    }
                    uint4 pk;
                    pk.x = PACK_P(e[0], e[1]);
                    pk.y = PACK_P(e[2], e[3]);
                    pk.z = PACK_P(e[4], e[5]);
                    pk.w = PACK_P(e[6], e[7]);
                    const int k16 = c * 4 + qd;                       // 16-byte chunk of the 128-byte P row
                    *reinterpret_cast<uint4*>(p_row + ((k16 ^ (r & 7)) << 4)) = pk;
                }
            }
            if (j > 0 && __any_sync(0xffffffffu, grow)) {             // lazy rescale of O (rare)
#pragma unroll 1
                for (int c = 0; c < HD / 32; ++c) {
                    uint32_t o[32];
                    tmem_ld_32x32(t_o + c * 32, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                    tmem_st_32x32(t_o + c * 32, o);
                }
                tmem_st_wait();
                l_sum *= alpha;
            }
            l_sum += (l4[0] + l4[1]) + (l4[2] + l4[3]);
            tc_fence_before();
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        mbar_wait(pv_done, (uint32_t)(n_kv - 1) & 1u);
        tc_fence_after();
        const int row = q0 + r;
        const float inv = 1.0f / l_sum;
        bf16* orow = O + ((long long)row_base + row) * o_rs + (long long)h * HD;
#pragma unroll 1
        for (int c = 0; c < HD / 32; ++c) {
            uint32_t o[32];
            tmem_ld_32x32(t_o + c * 32, o);
            tmem_ld_wait();
            if (row < S) {
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {
                    uint4 u;
                    u.x = pack_bf16x2(__uint_as_float(o[8 * qd + 0]) * inv, __uint_as_float(o[8 * qd + 1]) * inv);
                    u.y = pack_bf16x2(__uint_as_float(o[8 * qd + 2]) * inv, __uint_as_float(o[8 * qd + 3]) * inv);
                    u.z = pack_bf16x2(__uint_as_float(o[8 * qd + 4]) * inv, __uint_as_float(o[8 * qd + 5]) * inv);
                    u.w = pack_bf16x2(__uint_as_float(o[8 * qd + 6]) * inv, __uint_as_float(o[8 * qd + 7]) * inv);
                    *reinterpret_cast<uint4*>(orow + c * 32 + qd * 8) = u;
                }
            }
        }
        if (LSE && row < S) LSE[((long long)b * Hq + h) * S + row] = (m_ref + log2f(l_sum)) * 0.69314718055994531f;
        tc_fence_before();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

template <int HD, bool CAUSAL, int SLOTS>
int launch_attn_tc6s(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, bf16* o, float* lse, int B, int S, int Hq, int Hkv,
                     long long o_rs, float scale, cudaStream_t st, const int* kv_start) {
    using C = Att6Cfg<HD, SLOTS>;
    auto kern = attn_tc_fwd6_kernel<HD, CAUSAL, SLOTS>;
    static bool done = false;
    if (!done) {
        TA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        TA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        done = true;
    }
    const unsigned grid = (unsigned)(((S + BQ - 1) / BQ) * Hq * B);
    TA_KERNEL_LAUNCH(kern, grid, ATT1_THREADS, C::SMEM, st, tq, tk, tv, o, lse, S, Hq, Hkv, o_rs, scale * 1.4426950408889634f, kv_start);
    return 0;
}

// K/V ring depth.  Measured at B = 32, S = 464 (profiles/r02_c26_time_lm_attention.log): 4 slots 80.4 us, 3 slots 87.6 us (K_{j+1} cannot be
// requested before PV_{j-1} has released its slot), 6 / 8 slots with one CTA per SM 119 us, the 128-key-tile kernel 113.7 us.
// (cudaOccupancyMaxActiveBlocksPerMultiprocessor reports ONE resident CTA for every depth, 3 included; the timings say two are
// resident for 3 and 4 -- it is not used for the choice.)
int g_attn6_slots = 4;

template <int HD, bool CAUSAL>
int launch_attn_tc6(const bf16* q, const bf16* k, const bf16* v, bf16* o, float* lse, int B, int S, int Hq, int Hkv, long long q_rs,
                    long long k_rs, long long v_rs, long long o_rs, float scale, cudaStream_t st, const int* kv_start) {
    CUtensorMap tq, tk, tv;                                  // K / V boxes hold 64 rows here
    const long long rows = (long long)B * S;
    int rc = k_make_tensor_map_2d(&tq, q, rows, (long long)Hq * HD, q_rs, BQ);
    if (rc) return rc;
    rc = k_make_tensor_map_2d(&tk, k, rows, (long long)Hkv * HD, k_rs, BKV5);
    if (rc) return rc;
    rc = k_make_tensor_map_2d(&tv, v, rows, (long long)Hkv * HD, v_rs, BKV5);
    if (rc) return rc;
    if (g_attn6_slots == 4) return launch_attn_tc6s<HD, CAUSAL, 4>(tq, tk, tv, o, lse, B, S, Hq, Hkv, o_rs, scale, st, kv_start);
    if (g_attn6_slots == 6) return launch_attn_tc6s<HD, CAUSAL, 6>(tq, tk, tv, o, lse, B, S, Hq, Hkv, o_rs, scale, st, kv_start);
    if (g_attn6_slots == 8) return launch_attn_tc6s<HD, CAUSAL, 8>(tq, tk, tv, o, lse, B, S, Hq, Hkv, o_rs, scale, st, kv_start);
    return launch_attn_tc6s<HD, CAUSAL, 3>(tq, tk, tv, o, lse, B, S, Hq, Hkv, o_rs, scale, st, kv_start);
}

int g_attn_tc_lm = 1;   // 1 (default): decoder shape (head_dim 128, causal) on variant 6; 0: on attn_tc_fwd_kernel<128, true> (ta_attn_set_tc_lm)

// Forward-kernel selection (ta_attn_set_tc).  0: mma.sync; 1: tcgen05, two threads per query row, every shape; the other modes choose
// the encoder-shape kernel (head_dim 64, non-causal) and leave the rest on mode 1's kernel:
//   2: one thread per row, 128-key tiles, two CTAs per SM (round-1 default: 0.590 ms per encoder layer at B = 32, S = 1500);
//      3 / 4 / 11: the same with every 4th / 2nd / 8th exp2 on the FMA pipe;  5: two 64-key streams per row;
//   6-10, 12, 13: persistent CTA with two query tiles (turns / free-running, P through shared or tensor memory; 0.58-0.64 ms);
//   14 (default): 64-key tiles, three CTAs per SM -- three softmax warps per SM sub-partition keep the SFUs busier: 0.527 ms;
//      15 / 16: the same with every 8th / 4th exp2 on the FMA pipe (0.521 / slower).
// Measurements: profiles/r02_attention_*.txt.
int g_attn_tc = 14;

}  // namespace

// diagnostic: the persistent encoder-attention kernel (modes 6+) records the SM clock at its pipeline events for CTA 0 into
// buf [10 slots: softmax warps 0-7 (group A = 0-3, B = 4-7), MMA threads of A and B][steps][8] (int64, device memory; NULL switches it off)
TA_API int ta_attn_set_trace(void* buf, int steps) {
    g_attn_trace = reinterpret_cast<long long*>(buf);
    g_attn_trace_steps = buf ? steps : 0;
    return 0;
}

TA_API int ta_attn_set_tc_lm(int variant) {      // 0: 128-key tiles; 1: 64-key tiles, 4-slot ring (default); 3 / 4 / 6 / 8: that ring depth
    g_attn_tc_lm = variant ? 1 : 0;
    g_attn6_slots = (variant == 3 || variant == 6 || variant == 8) ? variant : 4;
    return 0;
}
TA_API int ta_attn_tc_lm_ring_slots(void) { return g_attn6_slots; }
TA_API int ta_attn_set_tc(int on) {
    g_attn_tc = (on < 0 || on > 16) ? 14 : on;
    return 0;
}
int k_attn_tc_enabled() { return g_attn_tc; }

// internal: *handled = 1 if this shape runs on the tcgen05 kernel (and was launched), 0 if the caller should use mma.sync
int k_attn_tc_fwd(const bf16* q, const bf16* k, const bf16* v, bf16* o, float* lse, int B, int S, int Hq, int Hkv, int head_dim,
                  long long q_rs, long long k_rs, long long v_rs, long long o_rs, int causal, float scale, cudaStream_t st,
                  int* handled, const int* kv_start) {
    *handled = 0;
    if (kv_start != nullptr && !(g_attn_tc && head_dim == 128 && causal)) {
        ta_set_error("attention with left-padded keys (kv_start) runs on the tcgen05 causal head_dim-128 kernel only");
        return -1;
    }
    if (!g_attn_tc || (head_dim != 64 && head_dim != 128) || S < 1 || Hq % Hkv != 0) return 0;
    if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
         reinterpret_cast<uintptr_t>(o)) & 15)
        return 0;
    CUtensorMap tq, tk, tv;
    const long long rows = (long long)B * S;
    int rc = k_make_tensor_map_2d(&tq, q, rows, (long long)Hq * head_dim, q_rs, BQ);
    if (rc) return rc;
    rc = k_make_tensor_map_2d(&tk, k, rows, (long long)Hkv * head_dim, k_rs, BKV);
    if (rc) return rc;
    rc = k_make_tensor_map_2d(&tv, v, rows, (long long)Hkv * head_dim, v_rs, BKV);
    if (rc) return rc;
    *handled = 1;
    if (head_dim == 64 && !causal && (g_attn_tc == 14 || g_attn_tc == 15 || g_attn_tc == 16)) {      // 64-key tiles, three CTAs per SM
        if (g_attn_tc == 14) return launch_attn_tc5<0>(q, k, v, o, lse, B, S, Hq, Hkv, q_rs, k_rs, v_rs, o_rs, scale, st);
        if (g_attn_tc == 15) return launch_attn_tc5<8>(q, k, v, o, lse, B, S, Hq, Hkv, q_rs, k_rs, v_rs, o_rs, scale, st);
        return launch_attn_tc5<4>(q, k, v, o, lse, B, S, Hq, Hkv, q_rs, k_rs, v_rs, o_rs, scale, st);
    }
    if (head_dim == 64 && !causal && g_attn_tc >= 2) {
        // persistent two-tile kernel: 6 = strict SFU turns, P through shared memory; 7 = free-running groups, P through shared memory;
        // 8 = free-running, P in tensor memory; 9 = turns + P in tensor memory; 10 = 8 with every 4th exp2 on the FMA pipe
        if (g_attn_tc == 6) return launch_attn_tc3<0, true, false>(tq, tk, tv, o, lse, B, S, Hq, Hkv, o_rs, scale, st);
        if (g_attn_tc == 7) return launch_attn_tc3<0, false, false>(tq, tk, tv, o, lse, B, S, Hq, Hkv, o_rs, scale, st);
        if (g_attn_tc == 8) return launch_attn_tc3<0, false, true>(tq, tk, tv, o, lse, B, S, Hq, Hkv, o_rs, scale, st);
        if (g_attn_tc == 9) return launch_attn_tc3<0, true, true>(tq, tk, tv, o, lse, B, S, Hq, Hkv, o_rs, scale, st);
        if (g_attn_tc == 10) return launch_attn_tc3<4, false, true>(tq, tk, tv, o, lse, B, S, Hq, Hkv, o_rs, scale, st);
        if (g_attn_tc == 11) return launch_attn_tc1<64, false, 8>(tq, tk, tv, o, lse, B, S, Hq, Hkv, o_rs, scale, st);   // mode 2 + every 8th exp2 on the FMA pipe
        if (g_attn_tc == 12) return launch_attn_tc3<8, false, true>(tq, tk, tv, o, lse, B, S, Hq, Hkv, o_rs, scale, st);
        if (g_attn_tc == 13) return launch_attn_tc3<8, false, false>(tq, tk, tv, o, lse, B, S, Hq, Hkv, o_rs, scale, st);
        if (g_attn_tc == 5) return launch_attn_tc2<64>(tq, tk, tv, o, lse, B, S, Hq, Hkv, o_rs, scale, st);
        if (g_attn_tc == 2) return launch_attn_tc1<64, false, 0>(tq, tk, tv, o, lse, B, S, Hq, Hkv, o_rs, scale, st);
        if (g_attn_tc == 3) return launch_attn_tc1<64, false, 4>(tq, tk, tv, o, lse, B, S, Hq, Hkv, o_rs, scale, st);
        return launch_attn_tc1<64, false, 2>(tq, tk, tv, o, lse, B, S, Hq, Hkv, o_rs, scale, st);
    }
    if (head_dim == 64) {
        if (causal) return launch_attn_tc<64, true>(tq, tk, tv, o, lse, B, S, Hq, Hkv, o_rs, scale, st);
        return launch_attn_tc<64, false>(tq, tk, tv, o, lse, B, S, Hq, Hkv, o_rs, scale, st);
    }
    if (causal && g_attn_tc_lm) return launch_attn_tc6<128, true>(q, k, v, o, lse, B, S, Hq, Hkv, q_rs, k_rs, v_rs, o_rs, scale, st, kv_start);
    if (causal) return launch_attn_tc<128, true>(tq, tk, tv, o, lse, B, S, Hq, Hkv, o_rs, scale, st, kv_start);
    return launch_attn_tc<128, false>(tq, tk, tv, o, lse, B, S, Hq, Hkv, o_rs, scale, st);
}
