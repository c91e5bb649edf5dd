// HBM-bound pieces of the path: norms, rotary embeddings, im2col, embedding lookup + <audio> scatter, cross entropy,
// transposes, global-norm clip + AdamW.  All fp32 math on bf16 / fp32 storage, vectorised 16-byte accesses,
// one warp (or one block) per row.  Reference call sites are cited per function in include/tinyaudio_b200.h.
#include "common.cuh"
#include "kernels.cuh"
#include "tinyaudio_b200.h"

namespace {

__device__ __forceinline__ void ld8_bf16(const bf16* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ void st8_bf16(bf16* p, const float (&v)[8]) {
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]); u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void ld8_f32(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void st8_f32(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

// =============================================================================================================
// conv2 im2col: x [B, T, C] bf16 (time-major) -> out [B*T2, 3C], row (b,t') = [x(2t'-1) | x(2t') | x(2t'+1)]
// (conv k=3, stride 2, pad 1:  HF:models/glmasr/modeling_glmasr.py:303,318)
// =============================================================================================================
__global__ void im2col_k3_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int B, int T, int T2, int C, int stride) {
    const int chunks = C / 8;
    const long long total = (long long)B * T2 * 3 * chunks;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % chunks);
        const int tap = (int)((i / chunks) % 3);
        const long long bt = i / (3LL * chunks);
        const int t2 = (int)(bt % T2), b = (int)(bt / T2);
        const int t = t2 * stride + tap - 1;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (t >= 0 && t < T) v = *reinterpret_cast<const uint4*>(x + ((long long)b * T + t) * C + c * 8);
        *reinterpret_cast<uint4*>(out + bt * (3LL * C) + (long long)tap * C + c * 8) = v;
    }
}

// =============================================================================================================
// LayerNorm, bf16 -> bf16, fp32 statistics  (nn.LayerNorm eps 1e-5; autocast keeps layer_norm in fp32)
// one warp per row; D multiple of 256 up to 2048
// =============================================================================================================
template <int MAXV>   // MAXV = max 8-element vectors per lane
__global__ void __launch_bounds__(256)
layernorm_bf16_kernel(const bf16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                      bf16* __restrict__ y, long long rows, int D, float eps, int reverse) {
    TA_PDL_ENTRY();
    long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    // reverse: walk the rows from the end.  The producing GEMM wrote x in ascending row order, so its LAST rows are the ones still in the
    // 126 MB L2 (x alone is 123 MB at 48 000 x 1280); and the consuming GEMM reads y from row 0, which this order writes last.
    // Same-box A/B: step 122.4 / 123.1 -> 121.8 / 121.8 ms (profiles/r02_c21_*).  (A persistent variant that prefetches the next row
    // into registers was NOT faster: 3.65 vs 3.53 ms per step for the 65 launches, call 22.)
    if (reverse) row = rows - 1 - row;
    const int lane = threadIdx.x & 31;
    const int nv = D / 256;   // vectors per lane
    // the row stays PACKED (bf16) in registers between the three passes: 4 instead of 8 registers per vector, so twice as many
    // rows are in flight per SM (the kernel is HBM-latency bound: ncu showed 32 % active warps at 61 % of the copy bandwidth)
    uint4 raw[MAXV];
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
        if (i < nv) raw[i] = *reinterpret_cast<const uint4*>(x + row * D + (i * 32 + lane) * 8);
    auto unpack8 = [](const uint4& u, float (&v)[8]) {
        const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
    };
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
        if (i < nv) {
            float v[8];
            unpack8(raw[i], v);
#pragma unroll
            for (int j = 0; j < 8; ++j) s += v[j];
        }
    const float mean = warp_sum(s) / D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
        if (i < nv) {
            float v[8];
            asm volatile("" : "+r"(raw[i].x), "+r"(raw[i].y), "+r"(raw[i].z), "+r"(raw[i].w));   // re-unpack: do not keep 8 floats alive
            unpack8(raw[i], v);
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float d = v[j] - mean; q += d * d; }
        }
    const float rstd = rsqrtf(warp_sum(q) / D + eps);
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
        if (i < nv) {
            const int c = (i * 32 + lane) * 8;
            float v[8], ww[8], bb[8], o[8];
            asm volatile("" : "+r"(raw[i].x), "+r"(raw[i].y), "+r"(raw[i].z), "+r"(raw[i].w));
            unpack8(raw[i], v);
            ld8_f32(w + c, ww);
            ld8_f32(bias + c, bb);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = (v[j] - mean) * rstd * ww[j] + bb[j];
            st8_bf16(y + row * D + c, o);
        }
}

// =============================================================================================================
// RMSNorm fp32 residual stream -> bf16 (Qwen3RMSNorm, HF:models/qwen3/modeling_qwen3.py:59-64) ; optional row gather
// =============================================================================================================
template <int MAXV>
__global__ void rmsnorm_f32_kernel(const float* __restrict__ x, const float* __restrict__ w, bf16* __restrict__ y,
                                   const int* __restrict__ row_index, long long rows, int D, float eps, long long ldy) {
    TA_PDL_ENTRY();
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const long long src = row_index ? (long long)row_index[row] : row;
    const int nv = D / 256;
    float v[MAXV][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
        if (i < nv) {
            ld8_f32(x + src * D + (i * 32 + lane) * 8, v[i]);
#pragma unroll
            for (int j = 0; j < 8; ++j) s += v[i][j] * v[i][j];
        }
    const float rstd = rsqrtf(warp_sum(s) / D + eps);
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
        if (i < nv) {
            const int c = (i * 32 + lane) * 8;
            float ww[8], o[8];
            ld8_f32(w + c, ww);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = ww[j] * (v[i][j] * rstd);
            st8_bf16(y + row * ldy + c, o);
        }
}

// dx_out[dst] (+)= rmsnorm_bwd(dy (bf16), x[dst] (fp32), w);  dst = row_index ? row_index[row] : row
// accumulate != 0: dx_out += ...  (dx_out may alias the incoming residual gradient)
template <int MAXV>
__global__ void rmsnorm_f32_bwd_kernel(const bf16* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ w,
                                       float* __restrict__ dx, const int* __restrict__ row_index, long long rows, int D,
                                       float eps, int accumulate, bf16* __restrict__ dx_bf16, long long ld_b) {
    TA_PDL_ENTRY();
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const long long dst = row_index ? (long long)row_index[row] : row;
    const int nv = D / 256;
    float xv[MAXV][8], gv[MAXV][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
        if (i < nv) {
            const int c = (i * 32 + lane) * 8;
            ld8_f32(x + dst * D + c, xv[i]);
            float d8[8], ww[8];
            ld8_bf16(dy + row * D + c, d8);
            ld8_f32(w + c, ww);
#pragma unroll
            for (int j = 0; j < 8; ++j) { gv[i][j] = d8[j] * ww[j]; s += xv[i][j] * xv[i][j]; }
        }
    const float rstd = rsqrtf(warp_sum(s) / D + eps);
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
        if (i < nv) {
#pragma unroll
            for (int j = 0; j < 8; ++j) dot += gv[i][j] * xv[i][j] * rstd;
        }
    dot = warp_sum(dot) / D;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
        if (i < nv) {
            const int c = (i * 32 + lane) * 8;
            float o[8];
            if (accumulate) ld8_f32(dx + dst * D + c, o);
            else {
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] += rstd * (gv[i][j] - xv[i][j] * rstd * dot);
            st8_f32(dx + dst * D + c, o);
            if (dx_bf16) st8_bf16(dx_bf16 + dst * ld_b + c, o);   // operand of the next dgrad GEMM (autocast rounds the branch grad)
        }
}

// =============================================================================================================
// encoder partial RoPE in place on the fused qkv buffer (HF:models/glmasr/modeling_glmasr.py:156-171)
// qkv [rows, 3*H*hd]; q at col 0, k at col H*hd; rotary dims [0, rd) of every head, pairs (i, i + rd/2)
// cos/sin tables [S, rd/2] fp32 (already rounded to the activation dtype by the host, as the reference does)
// =============================================================================================================
__global__ void enc_rope_kernel(bf16* __restrict__ qkv, const float* __restrict__ cosT, const float* __restrict__ sinT,
                                long long rows, int S, int H, int hd, int rd) {
    const int half = rd / 2;
    const long long total = rows * 2 * H * half;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(i % half);
        const int h = (int)((i / half) % H);
        const int qk = (int)((i / ((long long)half * H)) % 2);
        const long long row = i / ((long long)half * H * 2);
        const int pos = (int)(row % S);
        bf16* base = qkv + row * (3LL * H * hd) + (long long)qk * H * hd + (long long)h * hd;
        const float c = cosT[pos * half + p], s = sinT[pos * half + p];
        const float x1 = __bfloat162float(base[p]), x2 = __bfloat162float(base[p + half]);
        base[p] = __float2bfloat16_rn(x1 * c - x2 * s);
        base[p + half] = __float2bfloat16_rn(x2 * c + x1 * s);
    }
}

// =============================================================================================================
// Qwen3 per-head RMSNorm (q_norm / k_norm) + full RoPE  (HF:models/qwen3/modeling_qwen3.py:263-268, 159-181)
// in : qkv [M, (Hq+2Hkv)*128] bf16 (raw projections)     out: qk [M, (Hq+Hkv)*128] bf16 (normed + roped)
// one warp per (row, head); lane owns dims {2l, 2l+1, 2l+64, 2l+65}
// =============================================================================================================
__global__ void lm_qknorm_rope_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ qk, const float* __restrict__ qw,
                                          const float* __restrict__ kw, const float* __restrict__ cosT,
                                          const float* __restrict__ sinT, long long M, int S, int Hq, int Hkv, float eps,
                                          const int* __restrict__ pos_ids) {
    TA_PDL_ENTRY();
    // 8 lanes per (row, head): lane l8 owns dims [8 l8, +8) and [64 + 8 l8, +8) -- the RoPE pairs (d, d + 64) stay in one thread and every
    // access is a 16-byte vector (a warp covers 4 consecutive heads = 1 KB contiguous); 4-byte accesses ran at 40 % of the HBM rate
    const int HD = 128;
    const int heads = Hq + Hkv;
    const long long item = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int l8 = threadIdx.x & 7;
    const bool live = item < M * heads;
    const long long it = live ? item : 0;
    const int h = (int)(it % heads);
    const long long row = it / heads;
    // position ids are arange(S) (HF:qwen3:397-400) unless the caller supplies them (left-padded prompts in generate())
    const int pos = pos_ids ? pos_ids[row] : (int)(row % S);
    const bf16* src = qkv + row * (long long)((Hq + 2 * Hkv) * HD) + (long long)h * HD + 8 * l8;
    const float* w = ((h < Hq) ? qw : kw) + 8 * l8;
    float a[8], b[8];
    ld8_bf16(src, a);
    ld8_bf16(src + 64, b);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) ss += a[i] * a[i] + b[i] * b[i];
    ss += __shfl_xor_sync(0xffffffffu, ss, 1);
    ss += __shfl_xor_sync(0xffffffffu, ss, 2);
    ss += __shfl_xor_sync(0xffffffffu, ss, 4);
    const float rstd = rsqrtf(ss / HD + eps);
    float wl[8], wh[8], c[8], sn[8], lo[8], hi[8];
    ld8_f32(w, wl);
    ld8_f32(w + 64, wh);
    ld8_f32(cosT + pos * 64 + 8 * l8, c);
    ld8_f32(sinT + pos * 64 + 8 * l8, sn);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        // normed value is cast back to the input dtype (bf16) before the gain is applied
        const float n_lo = bf16_round(a[i] * rstd) * wl[i], n_hi = bf16_round(b[i] * rstd) * wh[i];
        lo[i] = n_lo * c[i] - n_hi * sn[i];
        hi[i] = n_hi * c[i] + n_lo * sn[i];
    }
    if (live) {
        bf16* dst = qk + row * (long long)(heads * HD) + (long long)h * HD + 8 * l8;
        st8_bf16(dst, lo);
        st8_bf16(dst + 64, hi);
    }
}

// backward of the above: (dq fp32 [M,Hq*128], dk bf16 [M,Hkv*128], dv bf16 [M,Hkv*128]) -> d_qkv bf16 [M,(Hq+2Hkv)*128]
__global__ void lm_qknorm_rope_bwd_kernel(const bf16* __restrict__ qkv, const float* __restrict__ dq, const bf16* __restrict__ dk,
                                          const bf16* __restrict__ dv, bf16* __restrict__ dqkv, const float* __restrict__ qw,
                                          const float* __restrict__ kw, const float* __restrict__ cosT,
                                          const float* __restrict__ sinT, long long M, int S, int Hq, int Hkv, float eps,
                                          long long ld_out) {
    TA_PDL_ENTRY();
    // 8 lanes per (row, head), lane l8 owns dims [8 l8, +8) and [64 + 8 l8, +8): 16-byte accesses throughout (see the forward kernel)
    const int HD = 128;
    const int heads = Hq + 2 * Hkv;
    const long long item = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int l8 = threadIdx.x & 7;
    const bool live = item < M * heads;
    const long long it = live ? item : 0;
    const int h = (int)(it % heads);
    const long long row = it / heads;
    bf16* dst = dqkv + row * ld_out + (long long)h * HD + 8 * l8;
    const bool is_v = h >= Hq + Hkv;
    float g_lo[8], g_hi[8];   // dy for my 8 + 8 dims
    if (is_v) {               // v: plain copy (still takes part in the shuffles below with zeros)
        const bf16* sv = dv + row * (long long)(Hkv * HD) + (long long)(h - Hq - Hkv) * HD + 8 * l8;
        if (live) {
            *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(sv);
            *reinterpret_cast<uint4*>(dst + 64) = *reinterpret_cast<const uint4*>(sv + 64);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) { g_lo[i] = 0.f; g_hi[i] = 0.f; }
    } else if (h < Hq) {
        const float* sq = dq + row * (long long)(Hq * HD) + (long long)h * HD + 8 * l8;
        ld8_f32(sq, g_lo);
        ld8_f32(sq + 64, g_hi);
    } else {
        const bf16* sk = dk + row * (long long)(Hkv * HD) + (long long)(h - Hq) * HD + 8 * l8;
        ld8_bf16(sk, g_lo);
        ld8_bf16(sk + 64, g_hi);
    }
    const int pos = (int)(row % S);
    const int hx = is_v ? 0 : h;                                   // any valid head for the (unused) v-lane loads
    const float* w = ((hx < Hq) ? qw : kw) + 8 * l8;
    const bf16* src = qkv + row * (long long)(heads * HD) + (long long)hx * HD + 8 * l8;
    float c[8], sn[8], wl[8], wh[8], a[8], b[8];
    ld8_f32(cosT + pos * 64 + 8 * l8, c);
    ld8_f32(sinT + pos * 64 + 8 * l8, sn);
    ld8_f32(w, wl);
    ld8_f32(w + 64, wh);
    ld8_bf16(src, a);
    ld8_bf16(src + 64, b);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) ss += a[i] * a[i] + b[i] * b[i];
    ss += __shfl_xor_sync(0xffffffffu, ss, 1);
    ss += __shfl_xor_sync(0xffffffffu, ss, 2);
    ss += __shfl_xor_sync(0xffffffffu, ss, 4);
    const float rstd = rsqrtf(ss / HD + eps);
    float d_lo[8], d_hi[8], n_lo[8], n_hi[8];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        // rope^T, then through the gain
        d_lo[i] = (g_lo[i] * c[i] + g_hi[i] * sn[i]) * wl[i];
        d_hi[i] = (g_hi[i] * c[i] - g_lo[i] * sn[i]) * wh[i];
        n_lo[i] = a[i] * rstd;
        n_hi[i] = b[i] * rstd;
        dot += d_lo[i] * n_lo[i] + d_hi[i] * n_hi[i];
    }
    dot += __shfl_xor_sync(0xffffffffu, dot, 1);
    dot += __shfl_xor_sync(0xffffffffu, dot, 2);
    dot += __shfl_xor_sync(0xffffffffu, dot, 4);
    dot /= HD;
    if (live && !is_v) {
        float o_lo[8], o_hi[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            o_lo[i] = rstd * (d_lo[i] - n_lo[i] * dot);
            o_hi[i] = rstd * (d_hi[i] - n_hi[i] * dot);
        }
        st8_bf16(dst, o_lo);
        st8_bf16(dst + 64, o_hi);
    }
}

// =============================================================================================================
// projector RMSNorm (+GELU) forward / backward on bf16 GEMM outputs (tiny_audio/projectors.py:66-71; LlamaRMSNorm)
//   fwd  : y = act( w * bf16(x * rstd) )            ACT: 0 = identity (fp32 out), 1 = exact GELU (bf16 out)
//   bwd  : dx (bf16) and dw (fp32, atomically accumulated over rows)
// one block per row, blockDim = 256, D <= 2048 (multiple of 256)
// =============================================================================================================
template <int ACT>
__global__ void __launch_bounds__(256)
proj_norm_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ w, void* __restrict__ y, int D, float eps) {
    __shared__ float red[40];
    const long long row = blockIdx.x;
    const int c = threadIdx.x * 8;
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float s = 0.f;
    if (c < D) {
        ld8_bf16(x + row * D + c, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[j] * v[j];
    }
    const float rstd = rsqrtf(block_sum(s, red) / D + eps);
    if (c < D) {
        float ww[8], o[8];
        ld8_f32(w + c, ww);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float z = ww[j] * bf16_round(v[j] * rstd);
            o[j] = ACT ? gelu_erf(z) : z;
        }
        if (ACT) st8_bf16(reinterpret_cast<bf16*>(y) + row * D + c, o);
        else st8_f32(reinterpret_cast<float*>(y) + row * D + c, o);
    }
}

// ROWS rows per block: dw accumulated in registers across the block's rows, then one atomicAdd per column
template <int ACT, typename GT>
__global__ void __launch_bounds__(256)
proj_norm_bwd_kernel(const bf16* __restrict__ x, const float* __restrict__ w, const GT* __restrict__ dy, bf16* __restrict__ dx,
                     float* __restrict__ dw, long long rows, int D, float eps, int rows_per_block) {
    __shared__ float red[40];
    const int c = threadIdx.x * 8;
    const bool on = c < D;
    float ww[8] = {0, 0, 0, 0, 0, 0, 0, 0}, dwacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (on) ld8_f32(w + c, ww);
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    for (long long row = r0; row < r0 + rows_per_block && row < rows; ++row) {
        float v[8] = {0, 0, 0, 0, 0, 0, 0, 0}, g[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        float s = 0.f;
        if (on) {
            ld8_bf16(x + row * D + c, v);
            if constexpr (sizeof(GT) == 2) ld8_bf16(reinterpret_cast<const bf16*>(dy) + row * D + c, g);
            else ld8_f32(reinterpret_cast<const float*>(dy) + row * D + c, g);
#pragma unroll
            for (int j = 0; j < 8; ++j) s += v[j] * v[j];
        }
        const float rstd = rsqrtf(block_sum(s, red) / D + eps);
        float dn[8], nh[8];
        float dot = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            nh[j] = bf16_round(v[j] * rstd);
            float dz = g[j];
            if (ACT) dz *= gelu_erf_grad(ww[j] * nh[j]);
            dwacc[j] += dz * nh[j];
            dn[j] = dz * ww[j];
            dot += dn[j] * v[j] * rstd;
        }
        dot = block_sum(dot, red) / D;
        if (on) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = rstd * (dn[j] - v[j] * rstd * dot);
            st8_bf16(dx + row * D + c, o);
        }
    }
    if (on && dw) {
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(dw + c + j, dwacc[j]);
    }
}

// =============================================================================================================
// embedding lookup + <audio> scatter (tiny_audio/asr_modeling.py:27-44, 497-515) -- index semantics are exact:
// the j-th <audio> position in row-major (b, s) order receives packed row j, where the packed rows are the first
// counts[i] rows of sample i (zero rows if counts[i] exceeds the projector output length)
// =============================================================================================================
// single block: src_row[p] = -1 (text token) | -2 (zero row) | flat row index into audio_embeds [B*n_a]
__global__ void __launch_bounds__(1024)
audio_index_kernel(const long long* __restrict__ ids, const long long* __restrict__ counts, int* __restrict__ src_row, int B,
                   int S, int n_a, long long audio_id) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    __shared__ long long s_cum[1025];
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    if (tid == 0) {
        s_carry = 0;
        long long c = 0;
        for (int i = 0; i < B && i < 1024; ++i) { s_cum[i] = c; c += counts ? counts[i] : 0; }
        s_cum[min(B, 1024)] = c;
    }
    __syncthreads();
    const int total = B * S;
    for (int base = 0; base < total; base += 1024) {
        const int p = base + tid;
        const int flag = (p < total && ids[p] == audio_id) ? 1 : 0;
        int incl = flag;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        if (lane == 31) s_warp[wp] = incl;
        __syncthreads();
        if (wp == 0) {
            int v = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += n;
            }
            s_warp[lane] = v;
        }
        __syncthreads();
        const int rank = s_carry + (wp ? s_warp[wp - 1] : 0) + incl - flag;   // exclusive global rank
        if (p < total) {
            int out = -1;
            if (flag) {
                out = -2;
                if (counts) {
                    // sample i with cum[i] <= rank < cum[i+1]
                    int lo = 0, hi = min(B, 1024);
                    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_cum[mid] <= rank) lo = mid; else hi = mid; }
                    if (rank < s_cum[min(B, 1024)]) {
                        const int r = rank - (int)s_cum[lo];
                        if (r < n_a) out = lo * n_a + r;
                    }
                } else {
                    // no explicit counts: sample b contributes exactly its own <audio> positions
                    out = -3;
                }
            }
            src_row[p] = out;
        }
        __syncthreads();
        if (tid == 1023) s_carry = s_carry + s_warp[31];
        __syncthreads();
    }
}

__global__ void embed_scatter_kernel(const long long* __restrict__ ids, const int* __restrict__ src_row,
                                     const float* __restrict__ table, const float* __restrict__ audio, float* __restrict__ out,
                                     long long n_tok, int D, long long vocab) {
    const int vecs = D / 4;
    const long long total = n_tok * vecs;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / vecs;
        const int c = (int)(i % vecs) * 4;
        const int sr = src_row[p];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (sr >= 0) v = *reinterpret_cast<const float4*>(audio + (long long)sr * D + c);
        else if (sr == -1) {
            long long id = ids[p];
            if (id < 0 || id >= vocab) id = 0;
            v = *reinterpret_cast<const float4*>(table + id * D + c);
        }
        *reinterpret_cast<float4*>(out + p * D + c) = v;
    }
}

// d_audio [n_rows_audio, D] (pre-zeroed) <- d_embeds rows at <audio> positions
__global__ void audio_grad_gather_kernel(const int* __restrict__ src_row, const float* __restrict__ d_emb,
                                         float* __restrict__ d_audio, long long n_tok, int D) {
    const int vecs = D / 4;
    const long long total = n_tok * vecs;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / vecs;
        const int c = (int)(i % vecs) * 4;
        const int sr = src_row[p];
        if (sr >= 0) *reinterpret_cast<float4*>(d_audio + (long long)sr * D + c) = *reinterpret_cast<const float4*>(d_emb + p * D + c);
    }
}

// =============================================================================================================
// *loss_sum += inv_items * sum(row_loss), one block, fixed summation order, double accumulation
__global__ void __launch_bounds__(1024) ce_loss_reduce_kernel(const float* __restrict__ row_loss, long long rows, float inv_items,
                                                              float* __restrict__ loss_sum) {
    __shared__ double part[1024];
    double a = 0.0;
    for (long long i = threadIdx.x; i < rows; i += blockDim.x) a += (double)row_loss[i];
    part[threadIdx.x] = a;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) part[threadIdx.x] += part[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) *loss_sum += (float)(part[0] * (double)inv_items);
}

// cross entropy on bf16 logits of the labelled rows (HF:loss/loss_utils.py:28-67): fp32 upcast, sum / num_items
// one block per row; writes d(logits) in place (bf16) and accumulates the loss
// =============================================================================================================
__global__ void __launch_bounds__(1024)
ce_fwd_bwd_kernel(bf16* __restrict__ logits, long long ld, const int* __restrict__ targets, int V, int Vpad,
                  float inv_items, float* __restrict__ loss_sum, float* __restrict__ row_loss, int write_grad) {
    __shared__ float red[40];
    const long long row = blockIdx.x;
    bf16* lr = logits + row * ld;
    const int tgt = targets[row];
    const int nvec = Vpad / 8;
    // one pass for max and sum (online softmax: rescale the running sum when the running max grows), a second pass for the gradient
    float m_run = -INFINITY, s_run = 0.f;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
        float v[8];
        ld8_bf16(lr + i * 8, v);
        float vm = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (i * 8 + j >= V) v[j] = -INFINITY;
            vm = fmaxf(vm, v[j]);
        }
        if (vm > -INFINITY) {
            const float m_new = fmaxf(m_run, vm);
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc += __expf(v[j] - m_new);      // exp(-inf) = 0 for the padding columns
            s_run = s_run * __expf(m_run - m_new) + acc;                   // exp(-inf - m_new) = 0 on the first vector
            m_run = m_new;
        }
    }
    const float mx = block_max(m_run, red);
    float se = block_sum((m_run > -INFINITY) ? s_run * __expf(m_run - mx) : 0.f, red);
    const float lse = mx + logf(se);
    if (threadIdx.x == 0) {
        const float lt = __bfloat162float(lr[tgt]);
        const float l = lse - lt;
        // with a row-loss buffer the batch loss is summed afterwards in a fixed order and in double precision (ce_loss_reduce_kernel):
        // reproducible and independent of how the batch is sharded.  A float atomicAdd per row drifts by ~1e-5 over 4 000 rows of ~12
        // (measured: the same 64 clips on 1 vs 2 GPUs differed by 1.6e-5)
        if (row_loss) row_loss[row] = l;
        else atomicAdd(loss_sum, l * inv_items);
    }
    if (!write_grad) return;
    __syncthreads();
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
        float v[8], o[8];
        ld8_bf16(lr + i * 8, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = i * 8 + j;
            float p = (col < V) ? __expf(v[j] - lse) : 0.f;
            if (col == tgt) p -= 1.0f;
            o[j] = p * inv_items;
        }
        st8_bf16(lr + i * 8, o);
    }
}

// =============================================================================================================
// misc: transpose, fp32 -> bf16 cast, gather of bf16 rows
// =============================================================================================================
__global__ void transpose_bf16_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int R, int C, long long ld_in,
                                      long long ld_out) {
    __shared__ bf16 tile[32][34];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < R && c < C) ? in[(long long)r * ld_in + c] : __float2bfloat16_rn(0.f);
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (c < C && r < R) out[(long long)c * ld_out + r] = tile[threadIdx.x][i];
    }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, long long n) {
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += (long long)gridDim.x * blockDim.x * 4) {
        if (i + 3 < n) {
            const float4 v = *reinterpret_cast<const float4*>(in + i);
            uint2 u;
            u.x = pack_bf16x2(v.x, v.y);
            u.y = pack_bf16x2(v.z, v.w);
            *reinterpret_cast<uint2*>(out + i) = u;
        } else {
            for (long long j = i; j < n; ++j) out[j] = __float2bfloat16_rn(in[j]);
        }
    }
}

// fp32 [rows, D] -> bf16 rows with leading dimension ld_out (D multiple of 4)
__global__ void cast_rows_f32_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, long long rows, int D, long long ld_out) {
    const int vecs = D / 4;
    const long long total = rows * vecs;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / vecs;
        const int c = (int)(i % vecs) * 4;
        const float4 v = *reinterpret_cast<const float4*>(in + r * D + c);
        uint2 u;
        u.x = pack_bf16x2(v.x, v.y);
        u.y = pack_bf16x2(v.z, v.w);
        *reinterpret_cast<uint2*>(out + r * ld_out + c) = u;
    }
}

// h = bf16(silu(gate)) * up from the interleaved (gate, up) stash [M, 2F] (64-blocks) -> [M, ld_h]  (LoRA wgrad recompute)
__global__ void swiglu_h_kernel(const bf16* __restrict__ gu, bf16* __restrict__ h, long long M, int F, long long ld_h) {
    const int chunks = F / 8;
    const long long total = M * chunks;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / chunks;
        const int j = (int)(i % chunks) * 8;
        const bf16* src = gu + r * (2LL * F) + (j / 64) * 128 + (j % 64);
        float g[8], u[8], o[8];
        ld8_bf16(src, g);
        ld8_bf16(src + 64, u);
#pragma unroll
        for (int t = 0; t < 8; ++t) o[t] = bf16_round(g[t] * sigmoidf_(g[t])) * u[t];
        st8_bf16(h + r * ld_h + j, o);
    }
}

// frame-stack copy for sequence lengths that are not a multiple of k (otherwise the stack is a free view):
// out [B, n, k*D] <- x [B, S, D] rows 0 .. n*k-1      (tiny_audio/projectors.py:79-87)
__global__ void frame_stack_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int B, int S, int n, int k, int D) {
    const int chunks = (k * D) / 8;
    const long long total = (long long)B * n * chunks;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % chunks);
        const long long bj = i / chunks;
        const int j = (int)(bj % n), b = (int)(bj / n);
        *reinterpret_cast<uint4*>(out + bj * (long long)(k * D) + c * 8) =
            *reinterpret_cast<const uint4*>(x + ((long long)b * S + (long long)j * k) * D + c * 8);
    }
}

// =============================================================================================================
// global-norm clip + AdamW (torch.optim.AdamW semantics; clip_grad_norm_ with error_if_nonfinite=False)
// =============================================================================================================
// audio-token dropout (tiny_audio/asr_modeling.py:458-479): whole encoder frames are zeroed by a {0,1} keep mask, no rescale.
// x bf16 [rows, D] in place, keep fp32 [rows] (the Bernoulli draw itself comes from torch's generator, as in the reference)
__global__ void frame_keep_mask_kernel(bf16* __restrict__ x, const float* __restrict__ keep, long long rows, int D) {
    const int vecs = D / 8;
    const long long total = rows * vecs;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / vecs;
        if (keep[r] == 0.0f) *reinterpret_cast<uint4*>(x + r * D + (i % vecs) * 8) = make_uint4(0u, 0u, 0u, 0u);
    }
}

// Label bookkeeping of the causal-LM loss (HF:loss/loss_utils.py:56-59) on the device: position p = (b, s) predicts labels[b, s + 1];
// the flat indices p of the non-ignored (!= -100) positions are compacted in ascending order into rows[], their targets into
// targets[], and the count goes to *count.  One block (B * S is a few 10^4): ordered, deterministic.
__global__ void __launch_bounds__(1024)
label_rows_kernel(const long long* __restrict__ labels, int B, int S, int* __restrict__ rows, int* __restrict__ targets,
                  int* __restrict__ count) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    const int total = B * S;
    for (int base = 0; base < total; base += 1024) {
        const int p = base + tid;
        long long tgt = -100;
        if (p < total && (p % S) != S - 1) tgt = labels[p + 1];
        const int flag = tgt != -100 ? 1 : 0;
        int incl = flag;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        if (lane == 31) s_warp[wp] = incl;
        __syncthreads();
        if (wp == 0) {
            int v = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += n;
            }
            s_warp[lane] = v;
        }
        __syncthreads();
        if (flag) {
            const int rank = s_carry + (wp ? s_warp[wp - 1] : 0) + incl - 1;
            rows[rank] = p;
            targets[rank] = (int)tgt;
        }
        __syncthreads();
        if (tid == 1023) s_carry += s_warp[31];
        __syncthreads();
    }
    if (tid == 0) *count = s_carry;
}

// GPU-side collation (scripts/train.py:324-348 builds these on the CPU workers): chat-template prompt assembly on the
// device.  Row b of the batch is   prefix | <audio> x counts[b] | middle | response_b | suffix | pad ...
// with labels = -100 except the response tokens and the first suffix token (<|im_end|>), attention_mask = 1 on the real tokens.
// One thread per (b, s) position.
__global__ void assemble_prompts_kernel(const long long* __restrict__ counts, const long long* __restrict__ resp, const long long* __restrict__ resp_off,
                                        const long long* __restrict__ prefix, int n_prefix, const long long* __restrict__ middle, int n_middle,
                                        const long long* __restrict__ suffix, int n_suffix, long long audio_id, long long pad_id, int B, int S,
                                        long long* __restrict__ ids, long long* __restrict__ labels, long long* __restrict__ mask) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * S) return;
    const int b = (int)(i / S);
    int s = (int)(i % S);
    const int n_a = (int)counts[b], n_r = (int)(resp_off[b + 1] - resp_off[b]);
    long long id = pad_id, lab = -100, m = 1;
    if (s < n_prefix) id = prefix[s];
    else if ((s -= n_prefix) < n_a) id = audio_id;
    else if ((s -= n_a) < n_middle) id = middle[s];
    else if ((s -= n_middle) < n_r) { id = resp[resp_off[b] + s]; lab = id; }
    else if ((s -= n_r) < n_suffix) { id = suffix[s]; if (s == 0) lab = id; }
    else m = 0;
    ids[i] = id;
    labels[i] = lab;
    mask[i] = m;
}

// *out += sum(g^2), DETERMINISTIC: block partials go to a scratch array and the last block to finish adds them up in index order (in
// double).  With a float atomicAdd per block the clip norm differed in its last bit between data-parallel replicas holding the same
// all-reduced gradient, and their parameters drifted apart by an ulp per step (bench.py parity.dp.param_max_abs_diff_across_ranks).
constexpr int SUMSQ_MAX_BLOCKS = 148 * 4;
__device__ float g_sumsq_partial[SUMSQ_MAX_BLOCKS];
__device__ unsigned int g_sumsq_done = 0;          // one grad-norm reduction at a time per device (they are issued on one stream)
__global__ void sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
    __shared__ float red[40];
    __shared__ bool last;
    float s = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s += g[i] * g[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) {
        g_sumsq_partial[blockIdx.x] = s;
        __threadfence();
        last = (atomicAdd(&g_sumsq_done, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        double t = 0.0;
        for (unsigned b = 0; b < gridDim.x; ++b) t += (double)reinterpret_cast<volatile float*>(g_sumsq_partial)[b];
        *out += (float)t;
        g_sumsq_done = 0;
    }
}

__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                             long long n, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt,
                             float max_norm, const float* __restrict__ gnorm_sq) {
    float coef = 1.0f;
    if (gnorm_sq && max_norm > 0.f) {
        const float tot = sqrtf(*gnorm_sq);
        coef = fminf(1.0f, max_norm / (tot + 1e-6f));
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gi = g[i] * coef;
        float pi = p[i] * (1.0f - lr * wd);
        const float mi = b1 * m[i] + (1.0f - b1) * gi;
        const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        pi -= (lr / bc1) * (mi / denom);
        p[i] = pi;
        m[i] = mi;
        v[i] = vi;
    }
}

inline int grid_for(long long n, int threads, int max_blocks = 148 * 16) {
    long long b = (n + threads - 1) / threads;
    if (b > max_blocks) b = max_blocks;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------
// internal launchers (kernels.cuh) -- also exported through thin C-ABI wrappers below for unit parity tests
// ------------------------------------------------------------------------------------------------------------
int k_im2col_k3(const bf16* x, bf16* out, int B, int T, int C, int stride, cudaStream_t st) {
    TA_REQUIRE(C % 8 == 0, "im2col: C must be a multiple of 8");
    const int T2 = (T + 2 - 3) / stride + 1;
    const long long total = (long long)B * T2 * 3 * (C / 8);
    im2col_k3_kernel<<<grid_for(total, 256), 256, 0, st>>>(x, out, B, T, T2, C, stride);
    TA_LAUNCH_CHECK();
    return 0;
}

int g_norm_wpb = 4;     // warps (= rows) per block of the LayerNorm / RMSNorm kernels (ta_debug_set(3, n): 1..8).  tools/sweep_norms.py at the
                        // production shapes: 4 is 2-5 % faster than 8 on all three kernels (profiles/r02_c31_sweep_norms.log)
int g_ln_reverse = 1;   // ta_layernorm_set_reverse: A/B switch for the row order of the encoder LayerNorm (L2 reuse, see the kernel)
int k_layernorm_bf16(const bf16* x, const float* w, const float* b, bf16* y, long long rows, int D, float eps, cudaStream_t st) {
    TA_REQUIRE(D % 256 == 0 && D <= 2048, "layernorm: D=%d must be a multiple of 256 and <= 2048", D);
    const int wpb = g_norm_wpb;
    const unsigned grid = (unsigned)((rows + wpb - 1) / wpb);
    if (D <= 1280) TA_KERNEL_LAUNCH(layernorm_bf16_kernel<5>, grid, wpb * 32, 0, st, x, w, b, y, rows, D, eps, g_ln_reverse);
    else TA_KERNEL_LAUNCH(layernorm_bf16_kernel<8>, grid, wpb * 32, 0, st, x, w, b, y, rows, D, eps, g_ln_reverse);
    return 0;
}

int k_rmsnorm_f32(const float* x, const float* w, bf16* y, const int* row_index, long long rows, int D, float eps, cudaStream_t st,
                  long long ldy) {
    if (ldy == 0) ldy = D;
    TA_REQUIRE(D % 256 == 0 && D <= 2048, "rmsnorm: D=%d must be a multiple of 256 and <= 2048", D);
    if (rows == 0) return 0;
    const int wpb = g_norm_wpb;
    const unsigned grid = (unsigned)((rows + wpb - 1) / wpb);
    if (D <= 1024) TA_KERNEL_LAUNCH(rmsnorm_f32_kernel<4>, grid, wpb * 32, 0, st, x, w, y, row_index, rows, D, eps, ldy);
    else TA_KERNEL_LAUNCH(rmsnorm_f32_kernel<8>, grid, wpb * 32, 0, st, x, w, y, row_index, rows, D, eps, ldy);
    return 0;
}

int k_rmsnorm_f32_bwd(const bf16* dy, const float* x, const float* w, float* dx, const int* row_index, long long rows, int D,
                      float eps, int accumulate, cudaStream_t st, bf16* dx_bf16, long long ld_b) {
    if (ld_b == 0) ld_b = D;
    TA_REQUIRE(D % 256 == 0 && D <= 2048, "rmsnorm bwd: D=%d must be a multiple of 256 and <= 2048", D);
    if (rows == 0) return 0;
    const int wpb = g_norm_wpb;
    const unsigned grid = (unsigned)((rows + wpb - 1) / wpb);
    if (D <= 1024)
        TA_KERNEL_LAUNCH(rmsnorm_f32_bwd_kernel<4>, grid, wpb * 32, 0, st, dy, x, w, dx, row_index, rows, D, eps, accumulate, dx_bf16, ld_b);
    else
        TA_KERNEL_LAUNCH(rmsnorm_f32_bwd_kernel<8>, grid, wpb * 32, 0, st, dy, x, w, dx, row_index, rows, D, eps, accumulate, dx_bf16, ld_b);
    return 0;
}

int k_enc_rope(bf16* qkv, const float* cosT, const float* sinT, long long rows, int S, int H, int hd, int rd, cudaStream_t st) {
    const long long total = rows * 2 * H * (rd / 2);
    enc_rope_kernel<<<grid_for(total, 256), 256, 0, st>>>(qkv, cosT, sinT, rows, S, H, hd, rd);
    TA_LAUNCH_CHECK();
    return 0;
}

int k_lm_qknorm_rope_fwd(const bf16* qkv, bf16* qk, const float* qw, const float* kw, const float* cosT, const float* sinT,
                         long long M, int S, int Hq, int Hkv, float eps, cudaStream_t st, const int* pos_ids) {
    const long long threads = M * (Hq + Hkv) * 8;      // 8 lanes per (row, head)
    if (threads == 0) return 0;
    TA_KERNEL_LAUNCH(lm_qknorm_rope_fwd_kernel, (unsigned)((threads + 255) / 256), 256, 0, st, qkv, qk, qw, kw, cosT, sinT, M, S, Hq, Hkv, eps,
                     pos_ids);
    return 0;
}

int k_lm_qknorm_rope_bwd(const bf16* qkv, const float* dq, const bf16* dk, const bf16* dv, bf16* dqkv, const float* qw,
                         const float* kw, const float* cosT, const float* sinT, long long M, int S, int Hq, int Hkv, float eps,
                         cudaStream_t st, long long ld_out) {
    if (ld_out == 0) ld_out = (long long)(Hq + 2 * Hkv) * 128;
    const long long threads = M * (Hq + 2 * Hkv) * 8;      // 8 lanes per (row, head)
    if (threads == 0) return 0;
    TA_KERNEL_LAUNCH(lm_qknorm_rope_bwd_kernel, (unsigned)((threads + 255) / 256), 256, 0, st, qkv, dq, dk, dv, dqkv, qw, kw, cosT, sinT, M, S,
                     Hq, Hkv, eps, ld_out);
    return 0;
}

int k_proj_norm_fwd(const bf16* x, const float* w, void* y, long long rows, int D, float eps, int gelu, cudaStream_t st) {
    TA_REQUIRE(D % 8 == 0 && D <= 2048, "projector norm: D=%d must be a multiple of 8 and <= 2048", D);
    if (rows == 0) return 0;
    if (gelu) proj_norm_fwd_kernel<1><<<(unsigned)rows, 256, 0, st>>>(x, w, y, D, eps);
    else proj_norm_fwd_kernel<0><<<(unsigned)rows, 256, 0, st>>>(x, w, y, D, eps);
    TA_LAUNCH_CHECK();
    return 0;
}

int k_proj_norm_bwd(const bf16* x, const float* w, const void* dy, int dy_is_f32, bf16* dx, float* dw, long long rows, int D,
                    float eps, int gelu, cudaStream_t st) {
    TA_REQUIRE(D % 8 == 0 && D <= 2048, "projector norm bwd: D=%d must be a multiple of 8 and <= 2048", D);
    if (rows == 0) return 0;
    const int rpb = 16;
    const unsigned grid = (unsigned)((rows + rpb - 1) / rpb);
    if (gelu && !dy_is_f32)
        proj_norm_bwd_kernel<1, bf16><<<grid, 256, 0, st>>>(x, w, (const bf16*)dy, dx, dw, rows, D, eps, rpb);
    else if (!gelu && dy_is_f32)
        proj_norm_bwd_kernel<0, float><<<grid, 256, 0, st>>>(x, w, (const float*)dy, dx, dw, rows, D, eps, rpb);
    else if (gelu && dy_is_f32)
        proj_norm_bwd_kernel<1, float><<<grid, 256, 0, st>>>(x, w, (const float*)dy, dx, dw, rows, D, eps, rpb);
    else
        proj_norm_bwd_kernel<0, bf16><<<grid, 256, 0, st>>>(x, w, (const bf16*)dy, dx, dw, rows, D, eps, rpb);
    TA_LAUNCH_CHECK();
    return 0;
}

int k_audio_index(const long long* ids, const long long* counts, int* src_row, int B, int S, int n_a, long long audio_id,
                  cudaStream_t st) {
    TA_REQUIRE(B <= 1024, "audio index: batch %d > 1024", B);
    audio_index_kernel<<<1, 1024, 0, st>>>(ids, counts, src_row, B, S, n_a, audio_id);
    TA_LAUNCH_CHECK();
    return 0;
}

int k_embed_scatter(const long long* ids, const int* src_row, const float* table, const float* audio, float* out,
                    long long n_tok, int D, long long vocab, cudaStream_t st) {
    TA_REQUIRE(D % 4 == 0, "embed: D must be a multiple of 4");
    if (n_tok == 0) return 0;
    embed_scatter_kernel<<<grid_for(n_tok * (D / 4), 256), 256, 0, st>>>(ids, src_row, table, audio, out, n_tok, D, vocab);
    TA_LAUNCH_CHECK();
    return 0;
}

int k_audio_grad_gather(const int* src_row, const float* d_emb, float* d_audio, long long n_tok, int D, cudaStream_t st) {
    if (n_tok == 0) return 0;
    audio_grad_gather_kernel<<<grid_for(n_tok * (D / 4), 256), 256, 0, st>>>(src_row, d_emb, d_audio, n_tok, D);
    TA_LAUNCH_CHECK();
    return 0;
}

int k_ce_fwd_bwd(bf16* logits, long long ld, const int* targets, long long rows, int V, int Vpad, float inv_items,
                 float* loss_sum, float* row_loss, int write_grad, cudaStream_t st) {
    TA_REQUIRE(Vpad % 8 == 0 && V <= Vpad, "cross entropy: bad vocab padding");
    if (rows == 0) return 0;
    ce_fwd_bwd_kernel<<<(unsigned)rows, 1024, 0, st>>>(logits, ld, targets, V, Vpad, inv_items, loss_sum, row_loss, write_grad);
    TA_LAUNCH_CHECK();
    if (row_loss) {
        ce_loss_reduce_kernel<<<1, 1024, 0, st>>>(row_loss, rows, inv_items, loss_sum);
        TA_LAUNCH_CHECK();
    }
    return 0;
}

int k_transpose_bf16(const bf16* in, bf16* out, int R, int C, long long ld_in, long long ld_out, cudaStream_t st) {
    dim3 grid((C + 31) / 32, (R + 31) / 32), block(32, 8);
    transpose_bf16_kernel<<<grid, block, 0, st>>>(in, out, R, C, ld_in, ld_out);
    TA_LAUNCH_CHECK();
    return 0;
}

int k_cast_f32_bf16(const float* in, bf16* out, long long n, cudaStream_t st) {
    if (n == 0) return 0;
    cast_f32_bf16_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, st>>>(in, out, n);
    TA_LAUNCH_CHECK();
    return 0;
}

int k_cast_rows_f32_bf16(const float* in, bf16* out, long long rows, int D, long long ld_out, cudaStream_t st) {
    TA_REQUIRE(D % 4 == 0 && ld_out % 4 == 0, "cast rows: D and ld must be multiples of 4");
    if (rows == 0) return 0;
    cast_rows_f32_bf16_kernel<<<grid_for(rows * (D / 4), 256), 256, 0, st>>>(in, out, rows, D, ld_out);
    TA_LAUNCH_CHECK();
    return 0;
}

int k_swiglu_h(const bf16* gu, bf16* h, long long M, int F, long long ld_h, cudaStream_t st) {
    TA_REQUIRE(F % 64 == 0, "swiglu_h: F must be a multiple of 64");
    if (M == 0) return 0;
    swiglu_h_kernel<<<grid_for(M * (F / 8), 256), 256, 0, st>>>(gu, h, M, F, ld_h);
    TA_LAUNCH_CHECK();
    return 0;
}

int k_frame_stack(const bf16* x, bf16* out, int B, int S, int n, int k, int D, cudaStream_t st) {
    TA_REQUIRE((k * D) % 8 == 0, "frame stack: k*D must be a multiple of 8");
    const long long total = (long long)B * n * ((k * D) / 8);
    if (total == 0) return 0;
    frame_stack_kernel<<<grid_for(total, 256), 256, 0, st>>>(x, out, B, S, n, k, D);
    TA_LAUNCH_CHECK();
    return 0;
}

int k_sumsq(const float* g, long long n, float* out, cudaStream_t st) {
    if (n == 0) return 0;
    sumsq_kernel<<<grid_for(n, 256, SUMSQ_MAX_BLOCKS), 256, 0, st>>>(g, n, out);
    TA_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------------
#define ST(s) reinterpret_cast<cudaStream_t>(s)

TA_API int ta_im2col_k3(const void* x, void* out, int B, int T, int C, int stride, void* stream) {
    return k_im2col_k3((const bf16*)x, (bf16*)out, B, T, C, stride, ST(stream));
}
TA_API int ta_layernorm_set_reverse(int on) {
    g_ln_reverse = on ? 1 : 0;
    return 0;
}
TA_API int ta_layernorm_bf16(const void* x, const float* w, const float* b, void* y, long long rows, int D, float eps, void* stream) {
    return k_layernorm_bf16((const bf16*)x, w, b, (bf16*)y, rows, D, eps, ST(stream));
}
TA_API int ta_rmsnorm_f32(const float* x, const float* w, void* y, const int* row_index, long long rows, int D, float eps, void* stream) {
    return k_rmsnorm_f32(x, w, (bf16*)y, row_index, rows, D, eps, ST(stream), 0);
}
TA_API int ta_rmsnorm_f32_bwd(const void* dy, const float* x, const float* w, float* dx, const int* row_index, long long rows, int D,
                              float eps, int accumulate, void* stream) {
    return k_rmsnorm_f32_bwd((const bf16*)dy, x, w, dx, row_index, rows, D, eps, accumulate, ST(stream), nullptr, 0);
}
TA_API int ta_enc_rope(void* qkv, const float* cosT, const float* sinT, long long rows, int S, int H, int hd, int rd, void* stream) {
    return k_enc_rope((bf16*)qkv, cosT, sinT, rows, S, H, hd, rd, ST(stream));
}
TA_API int ta_lm_qknorm_rope_fwd(const void* qkv, void* qk, const float* qw, const float* kw, const float* cosT, const float* sinT,
                                 long long M, int S, int Hq, int Hkv, float eps, void* stream) {
    return k_lm_qknorm_rope_fwd((const bf16*)qkv, (bf16*)qk, qw, kw, cosT, sinT, M, S, Hq, Hkv, eps, ST(stream), nullptr);
}
TA_API int ta_lm_qknorm_rope_bwd(const void* qkv, const float* dq, const void* dk, const void* dv, void* dqkv, const float* qw,
                                 const float* kw, const float* cosT, const float* sinT, long long M, int S, int Hq, int Hkv,
                                 float eps, void* stream) {
    return k_lm_qknorm_rope_bwd((const bf16*)qkv, dq, (const bf16*)dk, (const bf16*)dv, (bf16*)dqkv, qw, kw, cosT, sinT, M, S, Hq,
                                Hkv, eps, ST(stream), 0);
}
TA_API int ta_proj_norm_fwd(const void* x, const float* w, void* y, long long rows, int D, float eps, int gelu, void* stream) {
    return k_proj_norm_fwd((const bf16*)x, w, y, rows, D, eps, gelu, ST(stream));
}
TA_API int ta_proj_norm_bwd(const void* x, const float* w, const void* dy, int dy_is_f32, void* dx, float* dw, long long rows, int D,
                            float eps, int gelu, void* stream) {
    return k_proj_norm_bwd((const bf16*)x, w, dy, dy_is_f32, (bf16*)dx, dw, rows, D, eps, gelu, ST(stream));
}
TA_API int ta_audio_index(const long long* ids, const long long* counts, int* src_row, int B, int S, int n_a, long long audio_id,
                          void* stream) {
    return k_audio_index(ids, counts, src_row, B, S, n_a, audio_id, ST(stream));
}
TA_API int ta_embed_scatter(const long long* ids, const int* src_row, const float* table, const float* audio, float* out,
                            long long n_tok, int D, long long vocab, void* stream) {
    return k_embed_scatter(ids, src_row, table, audio, out, n_tok, D, vocab, ST(stream));
}
TA_API int ta_audio_grad_gather(const int* src_row, const float* d_emb, float* d_audio, long long n_tok, int D, void* stream) {
    return k_audio_grad_gather(src_row, d_emb, d_audio, n_tok, D, ST(stream));
}
TA_API int ta_ce_fwd_bwd(void* logits, long long ld, const int* targets, long long rows, int V, int Vpad, float inv_items,
                         float* loss_sum, float* row_loss, int write_grad, void* stream) {
    return k_ce_fwd_bwd((bf16*)logits, ld, targets, rows, V, Vpad, inv_items, loss_sum, row_loss, write_grad, ST(stream));
}
TA_API int ta_transpose_bf16(const void* in, void* out, int R, int C, long long ld_in, long long ld_out, void* stream) {
    return k_transpose_bf16((const bf16*)in, (bf16*)out, R, C, ld_in, ld_out, ST(stream));
}
TA_API int ta_cast_f32_bf16(const float* in, void* out, long long n, void* stream) {
    return k_cast_f32_bf16(in, (bf16*)out, n, ST(stream));
}
TA_API int ta_frame_stack(const void* x, void* out, int B, int S, int n, int k, int D, void* stream) {
    return k_frame_stack((const bf16*)x, (bf16*)out, B, S, n, k, D, ST(stream));
}

// encoder frames [rows, D] bf16 *= keep[rows] in {0, 1}  (audio_token_dropout, asr_modeling.py:458-479)
TA_API int ta_frame_keep_mask(void* x, const float* keep, long long rows, int D, void* stream) {
    TA_REQUIRE(x && keep, "ta_frame_keep_mask: null pointer");
    TA_REQUIRE(D % 8 == 0, "ta_frame_keep_mask: D=%d must be a multiple of 8", D);
    if (rows == 0) return 0;
    frame_keep_mask_kernel<<<grid_for(rows * (D / 8), 256), 256, 0, ST(stream)>>>((bf16*)x, keep, rows, D);
    TA_LAUNCH_CHECK();
    return 0;
}

// labels int64 [B, S] (device) -> ascending flat positions p with labels[p + 1] != -100 (same row), their targets, and the count
TA_API int ta_label_rows(const long long* labels, int B, int S, int* rows, int* targets, int* count, void* stream) {
    TA_REQUIRE(labels && rows && targets && count, "ta_label_rows: null pointer");
    TA_REQUIRE(B >= 0 && S >= 1 && (long long)B * S < (1LL << 31), "ta_label_rows: bad shape B=%d S=%d", B, S);
    label_rows_kernel<<<1, 1024, 0, ST(stream)>>>(labels, B, S, rows, targets, count);
    TA_LAUNCH_CHECK();
    return 0;
}

// GPU-side prompt assembly: counts [B] (audio tokens per clip), resp (packed response ids) + resp_off [B+1], template pieces as device
// arrays -> input_ids / labels / attention_mask [B, S] (all int64).  S is chosen by the caller (>= the longest row; longer rows are an error
// the caller rules out on the host, where the lengths come from).
TA_API int ta_assemble_prompts(const long long* counts, const long long* resp, const long long* resp_off, const long long* prefix, int n_prefix,
                               const long long* middle, int n_middle, const long long* suffix, int n_suffix, long long audio_id, long long pad_id,
                               int B, int S, long long* ids, long long* labels, long long* mask, void* stream) {
    TA_REQUIRE(counts && resp_off && ids && labels && mask, "ta_assemble_prompts: null pointer");
    TA_REQUIRE((n_prefix == 0 || prefix) && (n_middle == 0 || middle) && (n_suffix == 0 || suffix), "ta_assemble_prompts: template piece missing");
    if (B == 0 || S == 0) return 0;
    assemble_prompts_kernel<<<grid_for((long long)B * S, 256), 256, 0, ST(stream)>>>(counts, resp, resp_off, prefix, n_prefix, middle, n_middle,
                                                                                     suffix, n_suffix, audio_id, pad_id, B, S, ids, labels, mask);
    TA_LAUNCH_CHECK();
    return 0;
}

// grads -> global 2-norm^2 accumulated into *gnorm_sq (caller zeroes it once per step, calls this per tensor)
TA_API int ta_grad_sumsq(const float* g, long long n, float* gnorm_sq, void* stream) { return k_sumsq(g, n, gnorm_sq, ST(stream)); }

// one tensor of the fused clip + AdamW step; `step` is the 1-based step count
TA_API int ta_adamw_clip_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                              float eps, float weight_decay, int step, float max_grad_norm, const float* gnorm_sq, void* stream) {
    TA_REQUIRE(p && g && m && v, "ta_adamw_clip_step: null pointer");
    TA_REQUIRE(step >= 1, "ta_adamw_clip_step: step must be >= 1");
    if (n == 0) return 0;
    const float bc1 = 1.0f - powf(beta1, (float)step);
    const float bc2s = sqrtf(1.0f - powf(beta2, (float)step));
    adamw_kernel<<<grid_for(n, 256), 256, 0, ST(stream)>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2s,
                                                             max_grad_norm, gnorm_sq);
    TA_LAUNCH_CHECK();
    return 0;
}
