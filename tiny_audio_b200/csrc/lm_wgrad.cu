// Kernels of the unfrozen-LM recipe (reference: configs/experiments/embedded.yaml:19-33
// `freeze_language_model: false`, tiny_audio/asr_modeling.py:251-254 -- the whole Qwen3 decoder is trained):
//   * RMSNorm / per-head q,k-norm WEIGHT gradients (the frozen path only needs the input gradients),
//   * the embed_tokens gradient (scatter-add of d(inputs_embeds) over the text positions; the table is tied to lm_head,
//     whose wgrad is a GEMM in engine.cu),
//   * fp32 master weight -> packed bf16 operand refresh after every optimiser step (row-block remap for the fused qkv and
//     the 64-row gate/up interleave, plus the transposed dgrad copy) in one pass.
// The linear-weight gradients themselves are tcgen05 GEMMs (engine.cu: dW = dY^T X).
#include "common.cuh"
#include "kernels.cuh"
#include "tinyaudio_b200.h"

namespace {

__device__ __forceinline__ void ld8_bf16(const bf16* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ void ld8_f32(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// dw[c] += sum_rows dy[row, c] * x[src, c] * rstd(x[src])       (Qwen3RMSNorm: y = w * (x * rstd), HF:models/qwen3/modeling_qwen3.py:59-64)
// one warp per row (lane owns 8-element vectors), grid-stride over rows, register accumulators, one shared-memory reduction
// over the block's warps and one atomicAdd per column and block at the end.  D multiple of 256, D <= 256 * MAXV.
template <int MAXV>
__global__ void __launch_bounds__(256)
rmsnorm_dw_kernel(const bf16* __restrict__ dy, const float* __restrict__ x, const int* __restrict__ row_index, long long rows, int D,
                  float eps, float* __restrict__ dw) {
    __shared__ float s_acc[8][MAXV * 256 + 8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nv = D / 256;
    float acc[MAXV][8];
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (long long row = (long long)blockIdx.x * 8 + warp; row < rows; row += (long long)gridDim.x * 8) {
        const long long src = row_index ? (long long)row_index[row] : row;
        float xv[MAXV][8];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i)
            if (i < nv) {
                ld8_f32(x + src * D + (i * 32 + lane) * 8, xv[i]);
#pragma unroll
                for (int j = 0; j < 8; ++j) s += xv[i][j] * xv[i][j];
            }
        const float rstd = rsqrtf(warp_sum(s) / D + eps);
#pragma unroll
        for (int i = 0; i < MAXV; ++i)
            if (i < nv) {
                float d8[8];
                ld8_bf16(dy + row * D + (i * 32 + lane) * 8, d8);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(d8[j], xv[i][j] * rstd, acc[i][j]);
            }
    }
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
        if (i < nv) {
#pragma unroll
            for (int j = 0; j < 8; ++j) s_acc[warp][(i * 32 + lane) * 8 + j] = acc[i][j];
        }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += s_acc[w][c];
        atomicAdd(dw + c, s);
    }
}

// q_norm / k_norm weight gradients of one layer:  dw[d] += sum_{row, head} z[d] * bf16(x[d] * rstd)  with z = rope^T(dy)
// (forward: lm_qknorm_rope_fwd_kernel).  One warp per (row, head), lane owns dims {2l, 2l+1, 2l+64, 2l+65}.
__global__ void __launch_bounds__(256)
qknorm_dw_kernel(const bf16* __restrict__ qkv, const float* __restrict__ dq, const bf16* __restrict__ dk, const float* __restrict__ cosT,
                 const float* __restrict__ sinT, long long M, int S, int Hq, int Hkv, float eps, float* __restrict__ dqw,
                 float* __restrict__ dkw) {
    const int HD = 128;
    __shared__ float s_acc[8][2][HD];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int heads = Hq + Hkv;
    float aq[4] = {0.f, 0.f, 0.f, 0.f}, ak[4] = {0.f, 0.f, 0.f, 0.f};
    const long long total = M * heads;
    for (long long wid = (long long)blockIdx.x * 8 + warp; wid < total; wid += (long long)gridDim.x * 8) {
        const int h = (int)(wid % heads);
        const long long row = wid / heads;
        const int pos = (int)(row % S);
        float g0, g1, g2, g3;
        if (h < Hq) {
            const float* s = dq + row * (long long)(Hq * HD) + (long long)h * HD;
            const float2 lo = *reinterpret_cast<const float2*>(s + 2 * lane), hi = *reinterpret_cast<const float2*>(s + 64 + 2 * lane);
            g0 = lo.x; g1 = lo.y; g2 = hi.x; g3 = hi.y;
        } else {
            const bf16* s = dk + row * (long long)(Hkv * HD) + (long long)(h - Hq) * HD;
            const float2 lo = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(s + 2 * lane));
            const float2 hi = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(s + 64 + 2 * lane));
            g0 = lo.x; g1 = lo.y; g2 = hi.x; g3 = hi.y;
        }
        const float c0 = cosT[pos * 64 + 2 * lane], c1 = cosT[pos * 64 + 2 * lane + 1];
        const float s0 = sinT[pos * 64 + 2 * lane], s1 = sinT[pos * 64 + 2 * lane + 1];
        const float z0 = g0 * c0 + g2 * s0, z1 = g1 * c1 + g3 * s1;
        const float z2 = g2 * c0 - g0 * s0, z3 = g3 * c1 - g1 * s1;
        const bf16* src = qkv + row * (long long)((Hq + 2 * Hkv) * HD) + (long long)h * HD;
        const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(src + 2 * lane));
        const float2 b = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(src + 64 + 2 * lane));
        const float ss = warp_sum(a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y);
        const float rstd = rsqrtf(ss / HD + eps);
        const float n0 = bf16_round(a.x * rstd), n1 = bf16_round(a.y * rstd), n2 = bf16_round(b.x * rstd), n3 = bf16_round(b.y * rstd);
        if (h < Hq) { aq[0] = fmaf(z0, n0, aq[0]); aq[1] = fmaf(z1, n1, aq[1]); aq[2] = fmaf(z2, n2, aq[2]); aq[3] = fmaf(z3, n3, aq[3]); }
        else { ak[0] = fmaf(z0, n0, ak[0]); ak[1] = fmaf(z1, n1, ak[1]); ak[2] = fmaf(z2, n2, ak[2]); ak[3] = fmaf(z3, n3, ak[3]); }
    }
    s_acc[warp][0][2 * lane] = aq[0]; s_acc[warp][0][2 * lane + 1] = aq[1]; s_acc[warp][0][64 + 2 * lane] = aq[2]; s_acc[warp][0][64 + 2 * lane + 1] = aq[3];
    s_acc[warp][1][2 * lane] = ak[0]; s_acc[warp][1][2 * lane + 1] = ak[1]; s_acc[warp][1][64 + 2 * lane] = ak[2]; s_acc[warp][1][64 + 2 * lane + 1] = ak[3];
    __syncthreads();
    {
        const int which = threadIdx.x >> 7, d = threadIdx.x & 127;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += s_acc[w][which][d];
        atomicAdd((which ? dkw : dqw) + d, s);
    }
}

// d(embed_tokens)[id] += d(inputs_embeds)[p]  for every text position p (audio positions receive the projector output instead:
// tiny_audio/asr_modeling.py:497-515).  One block per token position.
__global__ void embed_grad_scatter_kernel(const long long* __restrict__ ids, const float* __restrict__ d_emb, float* __restrict__ d_table,
                                          int D, long long vocab, long long audio_id) {
    const long long p = blockIdx.x;
    const long long id = ids[p];
    if (id == audio_id || id < 0 || id >= vocab) return;
    for (int c = threadIdx.x; c < D; c += blockDim.x) atomicAdd(d_table + id * D + c, d_emb[p * D + c]);
}

// master fp32 [R, C] -> bf16 operand rows  dst[(r / blk) * blk_stride + r % blk + row_off, 0:C]  (leading dim ld) and, optionally, the
// transposed copy  dstT[c, same mapped row]  (leading dim ldT).  32 x 32 tiles through shared memory: both outputs coalesced.
__global__ void pack_weight_kernel(const float* __restrict__ src, int R, int C, bf16* __restrict__ dst, long long ld, bf16* __restrict__ dstT,
                                   long long ldT, int blk, int blk_stride, int row_off) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        const float v = (r < R && c < C) ? src[(long long)r * C + c] : 0.f;
        tile[i][threadIdx.x] = v;
        if (r < R && c < C) dst[(long long)((r / blk) * blk_stride + r % blk + row_off) * ld + c] = __float2bfloat16_rn(v);
    }
    if (dstT == nullptr) return;
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (c < C && r < R) dstT[(long long)c * ldT + (r / blk) * blk_stride + r % blk + row_off] = __float2bfloat16_rn(tile[threadIdx.x][i]);
    }
}

}  // namespace

int k_rmsnorm_dw(const bf16* dy, const float* x, const int* row_index, long long rows, int D, float eps, float* dw, cudaStream_t st) {
    TA_REQUIRE(D % 256 == 0 && D <= 1024, "rmsnorm dw: D must be a multiple of 256 and <= 1024 (D %d)", D);
    if (rows == 0) return 0;
    const long long want = (rows + 7) / 8;
    const int grid = (int)(want < 148 * 2 ? want : 148 * 2);
    rmsnorm_dw_kernel<4><<<grid, 256, 0, st>>>(dy, x, row_index, rows, D, eps, dw);
    TA_LAUNCH_CHECK();
    return 0;
}

int k_qknorm_dw(const bf16* qkv, const float* dq, const bf16* dk, const float* cosT, const float* sinT, long long M, int S, int Hq, int Hkv,
                float eps, float* dqw, float* dkw, cudaStream_t st) {
    if (M == 0) return 0;
    const long long want = (M * (Hq + Hkv) + 7) / 8;
    const int grid = (int)(want < 148 * 4 ? want : 148 * 4);
    qknorm_dw_kernel<<<grid, 256, 0, st>>>(qkv, dq, dk, cosT, sinT, M, S, Hq, Hkv, eps, dqw, dkw);
    TA_LAUNCH_CHECK();
    return 0;
}

int k_embed_grad_scatter(const long long* ids, const float* d_emb, float* d_table, long long n_tok, int D, long long vocab,
                         long long audio_id, cudaStream_t st) {
    if (n_tok == 0) return 0;
    embed_grad_scatter_kernel<<<(unsigned)n_tok, 256, 0, st>>>(ids, d_emb, d_table, D, vocab, audio_id);
    TA_LAUNCH_CHECK();
    return 0;
}

TA_API int ta_pack_weight(const float* src, int R, int C, void* dst, long long ld, void* dstT, long long ldT, int blk, int blk_stride,
                          int row_off, void* stream) {
    TA_REQUIRE(src && dst && R > 0 && C > 0 && blk > 0 && blk_stride >= blk, "ta_pack_weight: bad arguments");
    dim3 grid((C + 31) / 32, (R + 31) / 32), block(32, 8);
    pack_weight_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, R, C, (bf16*)dst, ld, (bf16*)dstT, ldT, blk, blk_stride,
                                                                                  row_off);
    TA_LAUNCH_CHECK();
    return 0;
}
TA_API int ta_rmsnorm_dw(const void* dy, const float* x, const int* row_index, long long rows, int D, float eps, float* dw, void* stream) {
    TA_REQUIRE(dy && x && dw, "ta_rmsnorm_dw: null pointer");
    return k_rmsnorm_dw((const bf16*)dy, x, row_index, rows, D, eps, dw, reinterpret_cast<cudaStream_t>(stream));
}
TA_API int ta_qknorm_dw(const void* qkv, const float* dq, const void* dk, const float* cos_t, const float* sin_t, long long M, int S, int Hq,
                        int Hkv, float eps, float* dqw, float* dkw, void* stream) {
    TA_REQUIRE(qkv && dq && dk && dqw && dkw, "ta_qknorm_dw: null pointer");
    return k_qknorm_dw((const bf16*)qkv, dq, (const bf16*)dk, cos_t, sin_t, M, S, Hq, Hkv, eps, dqw, dkw, reinterpret_cast<cudaStream_t>(stream));
}
TA_API int ta_embed_grad_scatter(const long long* ids, const float* d_emb, float* d_table, long long n_tok, int D, long long vocab,
                                 long long audio_id, void* stream) {
    TA_REQUIRE(ids && d_emb && d_table, "ta_embed_grad_scatter: null pointer");
    return k_embed_grad_scatter(ids, d_emb, d_table, n_tok, D, vocab, audio_id, reinterpret_cast<cudaStream_t>(stream));
}
