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

enum { SK_BF16 = TA_SKINNY_BF16, SK_F32_RESID = TA_SKINNY_F32_RESID, SK_SWIGLU = TA_SKINNY_SWIGLU, SK_PARTIAL = TA_SKINNY_PARTIAL };

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
constexpr int SK_GROUP = SK_WARPS * SK_UNROLL;      // 32-wide k chunks one CTA consumes per pipeline step (= 1024 k)
constexpr int SK_STAGES = 6;                        // cp.async ring depth: 5 steps x 32 KB of weights in flight per CTA
constexpr int SK_STAGE_BYTES = SK_UNROLL * 2 * SK_WARPS * 32 * 16;
constexpr int SK_SMEM = SK_STAGES * SK_STAGE_BYTES;

// TT = number of 8-token tiles (M <= 8 * TT).  Work item = (16-row weight tile, k split); persistent CTAs walk the items
// with stride gridDim.x.  Weights stream through a per-thread cp.async ring in shared memory (each thread copies exactly
// the 16-byte pieces it will itself feed to its MMAs, so the ring needs no barriers): SK_STAGES - 1 steps = 160 KB per SM
// in flight, across item boundaries and -- the weights being frozen -- already before the PDL wait.  When one step covers
// the whole contraction (K <= 1024, no split) the token fragments are loaded once per CTA and stay in registers.
//   SK_PARTIAL: out = fp32 [k_splits, M, ldo] raw partial sums (reduced in fixed order by decode_resid_rmsnorm_kernel).
//   SIMPLE (K == 1024, no split): one pipeline step per tile and every chunk valid -- the item / tile index arithmetic, its integer
//   divisions and all per-chunk predicates fold away at compile time (ncu: the general form executed 555 instructions per warp
//   and tile, most of them this bookkeeping, and was issue-bound at 2.75 TB/s on the lm_head).
template <int TT, int EPI, bool SIMPLE>
__global__ void __launch_bounds__(SK_WARPS * 32)
skinny_gemm_kernel(const bf16* __restrict__ X, long long ldx, const bf16* __restrict__ W, long long ldw, int M, int n_tiles, int K,
                   int k_splits_rt, void* __restrict__ out, long long ldo, const float* __restrict__ resid) {
    __shared__ float red[SK_WARPS][SK_ROWS][TT * 8 + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int k_splits = SIMPLE ? 1 : k_splits_rt;
    const int cps = SIMPLE ? SK_GROUP : (K >> 5) / k_splits;   // chunks per split
    const int gpi = SIMPLE ? 1 : (cps + SK_GROUP - 1) / SK_GROUP;   // pipeline steps per item
    const int n_items = n_tiles * k_splits;
    const int my_items = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int n_steps = my_items * gpi;
    const bool keep_x = (gpi == 1 && k_splits == 1);

    auto rows_of = [&](int tile, long long& row_lo, long long& row_hi) {
        if (EPI == SK_SWIGLU) {
            const int f0 = tile * 8;                                         // h features f0 .. f0 + 7
            row_lo = (long long)(f0 / 64) * 128 + (f0 % 64) + g;             // gate row of feature f0 + g
            row_hi = row_lo + 64;                                            // its up row
        } else {
            row_lo = (long long)tile * SK_ROWS + g;
            row_hi = row_lo + 8;
        }
    };
    extern __shared__ uint4 sk_ring[];                           // [SK_STAGES][2 * SK_UNROLL][256 threads]
    auto load_w = [&](int step) {                                // issue (never wait for) the weight pieces of one step
        if (step < n_steps) {
            const int item = (int)blockIdx.x + (step / gpi) * (int)gridDim.x;
            const int tile = item / k_splits, split = item % k_splits;
            long long row_lo, row_hi;
            rows_of(tile, row_lo, row_hi);
            const int c_end = (split + 1) * cps;
            const int c0 = split * cps + (step % gpi) * SK_GROUP + warp;
            const bf16* w_lo = W + row_lo * ldw + t * 8;
            const bf16* w_hi = W + row_hi * ldw + t * 8;
            uint4* slot = sk_ring + (step % SK_STAGES) * (2 * SK_UNROLL * SK_WARPS * 32) + threadIdx.x;
#pragma unroll
            for (int u = 0; u < SK_UNROLL; ++u) {
                const int c = c0 + u * SK_WARPS;
                if (SIMPLE || c < c_end) {
                    cp_async16(slot + (2 * u) * (SK_WARPS * 32), w_lo + c * 32, true);
                    cp_async16(slot + (2 * u + 1) * (SK_WARPS * 32), w_hi + c * 32, true);
                }
            }
        }
        cp_async_commit();                                       // one group per step, empty past the end: uniform counting
    };
    const bf16* x_row[TT];
#pragma unroll
    for (int j = 0; j < TT; ++j) x_row[j] = X + (long long)min(8 * j + g, M - 1) * ldx + t * 8;
    uint4 xa[SK_UNROLL][TT];
    auto load_x = [&](int c0, int c_end) {
#pragma unroll
        for (int u = 0; u < SK_UNROLL; ++u) {
            const int c = c0 + u * SK_WARPS;
            if (SIMPLE || c < c_end) {
#pragma unroll
                for (int j = 0; j < TT; ++j) xa[u][j] = *reinterpret_cast<const uint4*>(x_row[j] + c * 32);
            }
        }
    };
    float acc[TT][4];
#pragma unroll
    for (int j = 0; j < TT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;

    // the weights are frozen: start streaming them before waiting for the kernel that produces X (PDL overlap)
#pragma unroll 1
    for (int st = 0; st < SK_STAGES - 1; ++st) load_w(st);
    pdl_launch_dependents();
    pdl_wait();
    if (keep_x) load_x(warp, cps);

    for (int step = 0; step < n_steps; ++step) {
        const int item = (int)blockIdx.x + (step / gpi) * (int)gridDim.x;
        const int tile = item / k_splits, split = item % k_splits;
        const int c_end = (split + 1) * cps;
        const int c0 = split * cps + (step % gpi) * SK_GROUP + warp;
        if (!keep_x) load_x(c0, c_end);
        cp_async_wait<SK_STAGES - 2>();                          // this thread's pieces of `step` have landed
        uint4 ca[SK_UNROLL], cb[SK_UNROLL];
        {
            const uint4* slot = sk_ring + (step % SK_STAGES) * (2 * SK_UNROLL * SK_WARPS * 32) + threadIdx.x;
#pragma unroll
            for (int u = 0; u < SK_UNROLL; ++u) {
                if (SIMPLE || c0 + u * SK_WARPS < c_end) {
                    ca[u] = slot[(2 * u) * (SK_WARPS * 32)];
                    cb[u] = slot[(2 * u + 1) * (SK_WARPS * 32)];
                }
            }
        }
        load_w(step + SK_STAGES - 1);                            // refill the slot consumed one step ago
#pragma unroll
        for (int u = 0; u < SK_UNROLL; ++u) {
            if (SIMPLE || c0 + u * SK_WARPS < c_end) {
#pragma unroll
                for (int j = 0; j < TT; ++j) {
                    mma_bf16_16816(acc[j], ca[u].x, cb[u].x, ca[u].y, cb[u].y, xa[u][j].x, xa[u][j].y);
                    mma_bf16_16816(acc[j], ca[u].z, cb[u].z, ca[u].w, cb[u].w, xa[u][j].z, xa[u][j].w);
                }
            }
        }
        if ((step % gpi) != gpi - 1) continue;
        // ---- item finished: fixed-order reduction over the 8 warps' k slices, then the epilogue ----
#pragma unroll
        for (int j = 0; j < TT; ++j) {
            red[warp][g][8 * j + 2 * t] = acc[j][0];
            red[warp][g][8 * j + 2 * t + 1] = acc[j][1];
            red[warp][g + 8][8 * j + 2 * t] = acc[j][2];
            red[warp][g + 8][8 * j + 2 * t + 1] = acc[j][3];
            acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
        }
        __syncthreads();
        auto total = [&](int i, int m) {
            float sum = 0.f;
#pragma unroll
            for (int w = 0; w < SK_WARPS; ++w) sum += red[w][i][m];     // fixed order: bit-reproducible
            return sum;
        };
        if (EPI == SK_SWIGLU) {
            const int f0 = tile * 8;
            for (int idx = threadIdx.x; idx < 8 * TT * 8; idx += blockDim.x) {
                const int m = idx >> 3, i = idx & 7;
                if (m < M) {
                    const float gt = bf16_round(total(i, m)), up = bf16_round(total(i + 8, m));
                    reinterpret_cast<bf16*>(out)[(long long)m * ldo + f0 + i] = __float2bfloat16_rn(bf16_round(gt * sigmoidf_(gt)) * up);
                }
            }
        } else {
            const long long n0 = (long long)tile * SK_ROWS;
            for (int idx = threadIdx.x; idx < SK_ROWS * TT * 8; idx += blockDim.x) {
                const int m = idx >> 4, i = idx & 15;
                if (m < M) {
                    const float sum = total(i, m);
                    const long long o = (long long)m * ldo + n0 + i;
                    if (EPI == SK_BF16) reinterpret_cast<bf16*>(out)[o] = __float2bfloat16_rn(sum);
                    else if (EPI == SK_F32_RESID) reinterpret_cast<float*>(out)[o] = resid[o] + bf16_round(sum);
                    else reinterpret_cast<float*>(out)[(long long)split * M * ldo + o] = sum;
                }
            }
        }
        __syncthreads();      // `red` is rewritten by the next item
    }
}

template <int EPI>
int skinny_launch(const bf16* X, long long ldx, const bf16* W, long long ldw, int M, int N, int K, int k_splits, void* out,
                  long long ldo, const float* resid, cudaStream_t st) {
    const int n_tiles = N / SK_ROWS;
    const int items = n_tiles * k_splits;
    const int grid = items < 148 ? items : 148;      // one persistent CTA per SM (the ring takes most of its shared memory)
    const bool simple = (K == 32 * SK_GROUP && k_splits == 1);
    static bool attr_done = false;
    if (!attr_done) {
#define TA_SK_ATTR(TT_, S_) TA_CHECK_CUDA(cudaFuncSetAttribute(skinny_gemm_kernel<TT_, EPI, S_>, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_SMEM))
        TA_SK_ATTR(1, false); TA_SK_ATTR(2, false); TA_SK_ATTR(4, false);
        TA_SK_ATTR(1, true); TA_SK_ATTR(2, true); TA_SK_ATTR(4, true);
#undef TA_SK_ATTR
        attr_done = true;
    }
#define TA_SK_LAUNCH(TT_, S_)                                                                                                          \
    TA_CHECK_CUDA(launch_pdl(skinny_gemm_kernel<TT_, EPI, S_>, grid, SK_WARPS * 32, SK_SMEM, st, X, ldx, W, ldw, M, n_tiles, K, k_splits, out, \
                             ldo, resid))
    if (simple) {
        if (M <= 8) TA_SK_LAUNCH(1, true); else if (M <= 16) TA_SK_LAUNCH(2, true); else TA_SK_LAUNCH(4, true);
    } else {
        if (M <= 8) TA_SK_LAUNCH(1, false); else if (M <= 16) TA_SK_LAUNCH(2, false); else TA_SK_LAUNCH(4, false);
    }
#undef TA_SK_LAUNCH
    TA_LAUNCH_CHECK();
    return 0;
}

// x_out = x_in + bf16(sum_s partial[s])   (the linear's output is rounded to bf16 before the residual add, like the training
// path);  y = RMSNorm(x_out) * w -> bf16.  One block per token row; `partial` may be null (plain RMSNorm of x_in, x_out unused).
__global__ void __launch_bounds__(128)
decode_resid_rmsnorm_kernel(const float* __restrict__ x_in, const float* __restrict__ partial, int n_splits, int rows,
                            float* __restrict__ x_out, const float* __restrict__ w, bf16* __restrict__ y, int D, float eps,
                            long long ldy) {
    __shared__ float red[33];
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x;
    float v[2][4];
    float ssq = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int c = (i * 128 + threadIdx.x) * 4;
        v[i][0] = v[i][1] = v[i][2] = v[i][3] = 0.f;
        if (c < D) {
            const float4 a = *reinterpret_cast<const float4*>(x_in + (long long)row * D + c);
            v[i][0] = a.x; v[i][1] = a.y; v[i][2] = a.z; v[i][3] = a.w;
            if (partial) {
                float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int sp = 0; sp < n_splits; ++sp) {
                    const float4 p = *reinterpret_cast<const float4*>(partial + ((long long)sp * rows + row) * D + c);
                    sum.x += p.x; sum.y += p.y; sum.z += p.z; sum.w += p.w;
                }
                v[i][0] += bf16_round(sum.x); v[i][1] += bf16_round(sum.y); v[i][2] += bf16_round(sum.z); v[i][3] += bf16_round(sum.w);
                *reinterpret_cast<float4*>(x_out + (long long)row * D + c) = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
            }
            ssq += v[i][0] * v[i][0] + v[i][1] * v[i][1] + v[i][2] * v[i][2] + v[i][3] * v[i][3];
        }
    }
    const float rstd = rsqrtf(block_sum(ssq, red) / D + eps);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int c = (i * 128 + threadIdx.x) * 4;
        if (c < D) {
            const float4 gw = *reinterpret_cast<const float4*>(w + c);
            uint2 u;
            u.x = pack_bf16x2(gw.x * (v[i][0] * rstd), gw.y * (v[i][1] * rstd));
            u.y = pack_bf16x2(gw.z * (v[i][2] * rstd), gw.w * (v[i][3] * rstd));
            *reinterpret_cast<uint2*>(y + (long long)row * ldy + c) = u;
        }
    }
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
// decode attention, fused with what precedes it in the layer: per-head RMSNorm + RoPE of the new token's q / k (same
// arithmetic as lm_qknorm_rope_fwd_kernel; HF:models/qwen3/modeling_qwen3.py:263-268), the cache append of k / v at row
// pos = *pos_ptr (device side: graph-replayable), and single-query attention over cache rows [0, pos]
// (HF:integrations/sdpa_attention.py with a 1-token query; GQA: G = Hq / Hkv query heads share one K/V head).
// grid (Hkv, B): the CTA owns kv head blockIdx.x of sequence blockIdx.y and its G query heads, so everything it needs of
// the new token is local.  8 warps x 4 key slots, 8 lanes per key (16 head dims each); online softmax per slot, slots merged
// through shared memory.  Reads each cached K/V row exactly once: HBM-bound, 512 B per cached token and kv head.
// QKV != nullptr: fused mode (q / k / v taken from the raw projection row).  QKV == nullptr: q is ready, the cache holds rows
// [0, pos] already (unit test of the attention arithmetic alone).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int DA_WARPS = 8, DA_SLOTS = DA_WARPS * 4;

template <int G, int KPI>      // KPI = keys per slot and iteration (independent loads in flight; 2 costs registers -> 1 CTA / SM)
__global__ void __launch_bounds__(DA_WARPS * 32)
decode_attn_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ q_ready, bf16* k_cache, bf16* v_cache,
                   bf16* __restrict__ out, long long ld_out, const float* __restrict__ qw, const float* __restrict__ kw,
                   const float* __restrict__ cosT, const float* __restrict__ sinT, const int* __restrict__ pos_ptr, int Hq, int Hkv,
                   int max_seq, float eps, float scale_log2, const int* __restrict__ kv_start) {
    const int HD = 128;
    __shared__ float s_acc[DA_SLOTS][G][HD];
    __shared__ float s_m[DA_SLOTS][G], s_l[DA_SLOTS][G];
    __shared__ float s_q[G][HD];
    const int kvh = blockIdx.x, b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = warp * 4 + (lane >> 3), l8 = lane & 7;
    const int KD = Hkv * HD;
    pdl_launch_dependents();
    pdl_wait();
    const int pos = *pos_ptr;                          // cache row the new token is appended to (same for every sequence)
    // left-padded prompts (generate() with ragged prompts): sequence b's real tokens start at cache row ks; its rotary position
    // is counted from there (HF: position_ids = cumsum(attention_mask) - 1) and the padding rows are never attended to
    const int ks = kv_start ? kv_start[b] : 0;
    const int rpos = pos - ks;
    const int n_keys = rpos + 1;
    bf16* kb_w = k_cache + (long long)b * max_seq * KD + (long long)kvh * HD;
    bf16* vb_w = v_cache + (long long)b * max_seq * KD + (long long)kvh * HD;

    if (qkv != nullptr) {
        // warps 0 .. G-1: query heads; warp G: key head; warp G+1: value head
        const int heads = Hq + 2 * Hkv;
        if (warp < G + 2) {
            const int h = (warp < G) ? kvh * G + warp : (warp == G ? Hq + kvh : Hq + Hkv + kvh);
            const bf16* src = qkv + (long long)b * heads * HD + (long long)h * HD;
            const uint32_t ra = *reinterpret_cast<const uint32_t*>(src + 2 * lane);
            const uint32_t rb = *reinterpret_cast<const uint32_t*>(src + 64 + 2 * lane);
            if (warp == G + 1) {
                bf16* dst = vb_w + (long long)pos * KD;
                *reinterpret_cast<uint32_t*>(dst + 2 * lane) = ra;
                *reinterpret_cast<uint32_t*>(dst + 64 + 2 * lane) = rb;
            } else {
                const float* w = (warp < G) ? qw : kw;
                const float2 a = unpack_bf16x2(ra), bb = unpack_bf16x2(rb);
                const float ss = warp_sum(a.x * a.x + a.y * a.y + bb.x * bb.x + bb.y * bb.y);
                const float rstd = rsqrtf(ss / HD + eps);
                // normed value is cast back to bf16 before the gain is applied (Qwen3RMSNorm)
                const float n0 = bf16_round(a.x * rstd) * w[2 * lane], n1 = bf16_round(a.y * rstd) * w[2 * lane + 1];
                const float n2 = bf16_round(bb.x * rstd) * w[64 + 2 * lane], n3 = bf16_round(bb.y * rstd) * w[64 + 2 * lane + 1];
                const float c0 = cosT[rpos * 64 + 2 * lane], c1 = cosT[rpos * 64 + 2 * lane + 1];
                const float s0 = sinT[rpos * 64 + 2 * lane], s1 = sinT[rpos * 64 + 2 * lane + 1];
                const uint32_t lo = pack_bf16x2(n0 * c0 - n2 * s0, n1 * c1 - n3 * s1);
                const uint32_t hi = pack_bf16x2(n2 * c0 + n0 * s0, n3 * c1 + n1 * s1);
                if (warp == G) {
                    bf16* dst = kb_w + (long long)pos * KD;
                    *reinterpret_cast<uint32_t*>(dst + 2 * lane) = lo;
                    *reinterpret_cast<uint32_t*>(dst + 64 + 2 * lane) = hi;
                } else {      // the roped query is rounded to bf16 like the training path's qk buffer
                    const float2 l2 = unpack_bf16x2(lo), h2 = unpack_bf16x2(hi);
                    s_q[warp][2 * lane] = l2.x * scale_log2;
                    s_q[warp][2 * lane + 1] = l2.y * scale_log2;
                    s_q[warp][64 + 2 * lane] = h2.x * scale_log2;
                    s_q[warp][64 + 2 * lane + 1] = h2.y * scale_log2;
                }
            }
        }
    } else {
        for (int i = threadIdx.x; i < G * HD; i += blockDim.x)
            s_q[i / HD][i % HD] = __bfloat162float(q_ready[(long long)b * (Hq * HD) + (long long)(kvh * G) * HD + i]) * scale_log2;
    }
    __syncthreads();      // also makes this CTA's cache-row writes visible to its own loads below

    float qf[G][16];
#pragma unroll
    for (int gq = 0; gq < G; ++gq)
#pragma unroll
        for (int i = 0; i < 16; ++i) qf[gq][i] = s_q[gq][l8 * 16 + i];
    float m[G], l[G], acc[G][16];
#pragma unroll
    for (int gq = 0; gq < G; ++gq) {
        m[gq] = -INFINITY;
        l[gq] = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[gq][i] = 0.f;
    }
    const bf16* kb = kb_w + (long long)ks * KD + l8 * 16;       // keys / values of the real tokens: cache rows [ks, pos]
    const bf16* vb = vb_w + (long long)ks * KD + l8 * 16;
    uint4 nk[KPI][2], nv[KPI][2];
    auto load_kv = [&](int key0) {
#pragma unroll
        for (int e = 0; e < KPI; ++e) {
            const int kk = min(key0 + e * DA_SLOTS + slot, n_keys - 1);
            nk[e][0] = *reinterpret_cast<const uint4*>(kb + (long long)kk * KD);
            nk[e][1] = *reinterpret_cast<const uint4*>(kb + (long long)kk * KD + 8);
            nv[e][0] = *reinterpret_cast<const uint4*>(vb + (long long)kk * KD);
            nv[e][1] = *reinterpret_cast<const uint4*>(vb + (long long)kk * KD + 8);
        }
    };
    load_kv(0);
    for (int key0 = 0; key0 < n_keys; key0 += KPI * DA_SLOTS) {      // uniform trip count: the shuffles below need the full warp
        uint4 ck[KPI][2], cv[KPI][2];
#pragma unroll
        for (int e = 0; e < KPI; ++e) { ck[e][0] = nk[e][0]; ck[e][1] = nk[e][1]; cv[e][0] = nv[e][0]; cv[e][1] = nv[e][1]; }
        if (key0 + KPI * DA_SLOTS < n_keys) load_kv(key0 + KPI * DA_SLOTS);      // next rows in flight during this softmax update
#pragma unroll
        for (int e = 0; e < KPI; ++e) {
            const bool live = key0 + e * DA_SLOTS + slot < n_keys;
            const uint32_t kw8[8] = {ck[e][0].x, ck[e][0].y, ck[e][0].z, ck[e][0].w, ck[e][1].x, ck[e][1].y, ck[e][1].z, ck[e][1].w};
            const uint32_t vw8[8] = {cv[e][0].x, cv[e][0].y, cv[e][0].z, cv[e][0].w, cv[e][1].x, cv[e][1].y, cv[e][1].z, cv[e][1].w};
            float kf[16], vf[16];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float2 a = unpack_bf16x2(kw8[i]), c = unpack_bf16x2(vw8[i]);
                kf[2 * i] = a.x; kf[2 * i + 1] = a.y;
                vf[2 * i] = c.x; vf[2 * i + 1] = c.y;
            }
#pragma unroll
            for (int gq = 0; gq < G; ++gq) {
                float sc = 0.f;
#pragma unroll
                for (int i = 0; i < 16; ++i) sc = fmaf(qf[gq][i], kf[i], sc);
                sc += __shfl_xor_sync(0xffffffffu, sc, 1);
                sc += __shfl_xor_sync(0xffffffffu, sc, 2);
                sc += __shfl_xor_sync(0xffffffffu, sc, 4);
                if (live) {
                    const float m_new = fmaxf(m[gq], sc);
                    const float corr = exp2f(m[gq] - m_new), pr = exp2f(sc - m_new);
                    m[gq] = m_new;
                    l[gq] = l[gq] * corr + pr;
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc[gq][i] = fmaf(acc[gq][i], corr, pr * vf[i]);
                }
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
    pdl_launch_dependents();
    pdl_wait();
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
    pdl_launch_dependents();
    pdl_wait();
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
                  const float* resid, cudaStream_t st, int k_splits) {
    TA_REQUIRE(M >= 1 && M <= 32, "skinny GEMM: M = %d not in [1, 32]", M);
    TA_REQUIRE(N % 16 == 0 && K % 32 == 0 && ldx % 8 == 0 && ldw % 8 == 0, "skinny GEMM: N %% 16, K %% 32, ld %% 8 must be 0 (N %d K %d)", N, K);
    TA_REQUIRE(mode != SK_SWIGLU || N % 128 == 0, "skinny SwiGLU: N must be a multiple of 128");
    TA_REQUIRE(mode != SK_F32_RESID || resid, "skinny GEMM: residual missing");
    TA_REQUIRE(k_splits >= 1 && (K / 32) % k_splits == 0, "skinny GEMM: k_splits %d does not divide K / 32 = %d", k_splits, K / 32);
    TA_REQUIRE(k_splits == 1 || mode == SK_PARTIAL, "skinny GEMM: split-K needs the partial-sum mode");
    switch (mode) {
        case SK_BF16: return skinny_launch<SK_BF16>(X, ldx, W, ldw, M, N, K, 1, out, ldo, resid, st);
        case SK_F32_RESID: return skinny_launch<SK_F32_RESID>(X, ldx, W, ldw, M, N, K, 1, out, ldo, resid, st);
        case SK_SWIGLU: return skinny_launch<SK_SWIGLU>(X, ldx, W, ldw, M, N, K, 1, out, ldo, resid, st);
        case SK_PARTIAL: return skinny_launch<SK_PARTIAL>(X, ldx, W, ldw, M, N, K, k_splits, out, ldo, resid, st);
    }
    TA_REQUIRE(false, "skinny GEMM: unknown mode %d", mode);
}

int k_skinny_splits(int N, int K) {      // enough work items to fill the GPU when N is small
    const int chunks = K / 32;
    if (N / SK_ROWS >= 256) return 1;
    for (int sp = 4; sp > 1; --sp)
        if (chunks % sp == 0) return sp;
    return 1;
}

int k_kv_cache_store(const bf16* k_src, long long k_ld, const bf16* v_src, long long v_ld, bf16* k_cache, bf16* v_cache, int B, int S,
                     int KD, int max_seq, cudaStream_t st) {
    const long long total = (long long)B * S * (KD / 8);
    const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    kv_cache_store_kernel<<<grid, 256, 0, st>>>(k_src, k_ld, v_src, v_ld, k_cache, v_cache, B, S, KD, max_seq);
    TA_LAUNCH_CHECK();
    return 0;
}

int k_decode_attn(const bf16* qkv, const bf16* q_ready, bf16* k_cache, bf16* v_cache, bf16* out, long long ld_out, const float* qw,
                  const float* kw, const float* cosT, const float* sinT, const int* pos, int B, int Hq, int Hkv, int max_seq, float eps,
                  float scale, cudaStream_t st, const int* kv_start) {
    const float sl2 = scale * 1.4426950408889634f;
    dim3 grid(Hkv, B);
    const int G = Hq / Hkv;
    TA_REQUIRE(Hq % Hkv == 0 && (G == 1 || G == 2), "decode attention: Hq / Hkv = %d not supported (1 or 2)", G);
    // few CTAs (small batch): latency-bound, two keys per slot in flight; many CTAs: keep two CTAs resident per SM instead
    const bool deep = (long long)B * Hkv <= 148;
#define TA_DA_LAUNCH(GG, KK)                                                                                                     \
    TA_CHECK_CUDA(launch_pdl(decode_attn_kernel<GG, KK>, grid, DA_WARPS * 32, 0, st, qkv, q_ready, k_cache, v_cache, out, ld_out, qw, kw, \
                             cosT, sinT, pos, Hq, Hkv, max_seq, eps, sl2, kv_start))
    if (G == 1) {
        if (deep) TA_DA_LAUNCH(1, 2); else TA_DA_LAUNCH(1, 1);
    } else {
        if (deep) TA_DA_LAUNCH(2, 2); else TA_DA_LAUNCH(2, 1);
    }
#undef TA_DA_LAUNCH
    TA_LAUNCH_CHECK();
    return 0;
}

int k_decode_resid_rmsnorm(const float* x_in, const float* partial, int n_splits, int rows, float* x_out, const float* w, bf16* y, int D,
                           float eps, long long ldy, cudaStream_t st) {
    TA_REQUIRE(D % 4 == 0 && D <= 1024 && ldy % 4 == 0, "decode rmsnorm: D must be a multiple of 4 and <= 1024 (D %d)", D);
    TA_CHECK_CUDA(launch_pdl(decode_resid_rmsnorm_kernel, rows, 128, 0, st, x_in, partial, n_splits, rows, x_out, w, y, D, eps, ldy));
    TA_LAUNCH_CHECK();
    return 0;
}

int k_embed_rows(const long long* ids, const float* table, float* out, int B, int D, long long vocab, cudaStream_t st) {
    TA_CHECK_CUDA(launch_pdl(embed_rows_kernel, B, 256, 0, st, ids, table, out, D, vocab));
    TA_LAUNCH_CHECK();
    return 0;
}

int k_argmax_rows(const bf16* logits, long long ld, int rows, int V, long long* next_ids, int* pos_inc, cudaStream_t st) {
    TA_CHECK_CUDA(launch_pdl(argmax_rows_kernel, rows, 1024, 0, st, logits, ld, V, next_ids, pos_inc));
    TA_LAUNCH_CHECK();
    return 0;
}

// ---- exported for unit parity tests ----
TA_API int ta_skinny_gemm_bf16(const void* X, long long ldx, const void* W, long long ldw, int M, int N, int K, int mode, void* out,
                               long long ldo, const float* resid, int k_splits, void* stream) {
    TA_REQUIRE(X && W && out, "ta_skinny_gemm_bf16: null pointer");
    return k_skinny_gemm((const bf16*)X, ldx, (const bf16*)W, ldw, M, N, K, mode, out, ldo, resid, reinterpret_cast<cudaStream_t>(stream),
                         k_splits < 1 ? 1 : k_splits);
}
TA_API int ta_decode_resid_rmsnorm(const float* x_in, const float* partial, int n_splits, int rows, float* x_out, const float* w, void* y,
                                   int D, float eps, long long ldy, void* stream) {
    TA_REQUIRE(x_in && w && y && (!partial || x_out), "ta_decode_resid_rmsnorm: null pointer");
    return k_decode_resid_rmsnorm(x_in, partial, n_splits, rows, x_out, w, (bf16*)y, D, eps, ldy, reinterpret_cast<cudaStream_t>(stream));
}
TA_API int ta_decode_attn(const void* q, const void* k_cache, const void* v_cache, void* out, long long ld_out, const int* pos, int B,
                          int Hq, int Hkv, int max_seq, float scale, void* stream) {
    TA_REQUIRE(q && k_cache && v_cache && out && pos, "ta_decode_attn: null pointer");
    return k_decode_attn(nullptr, (const bf16*)q, (bf16*)const_cast<void*>(k_cache), (bf16*)const_cast<void*>(v_cache), (bf16*)out, ld_out,
                         nullptr, nullptr, nullptr, nullptr, pos, B, Hq, Hkv, max_seq, 0.f, scale, reinterpret_cast<cudaStream_t>(stream));
}
TA_API int ta_argmax_rows(const void* logits, long long ld, int rows, int V, long long* next_ids, void* stream) {
    TA_REQUIRE(logits && next_ids, "ta_argmax_rows: null pointer");
    return k_argmax_rows((const bf16*)logits, ld, rows, V, next_ids, nullptr, reinterpret_cast<cudaStream_t>(stream));
}
