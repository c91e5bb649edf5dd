// The QFormer projector's glue between its tcgen05 linears (tiny_audio/projectors.py:359-475, arithmetic of
// HF:models/blip_2/modeling_blip_2.py:537-1042): BertSelfOutput-style  LayerNorm(dropout(dense(x)) + residual)  forward and backward
// as ONE kernel each, exact-erf GELU forward / backward, and the column sums that are the bias gradients.  They replace the ATen
// layer_norm / add / copy / GammaBetaBackward / gelu / sum kernels of the round-1 trace (profiles/r02_c01_trace_qformer_v2.txt).
// All fp32 math; rows are a few thousand x 1280, so everything here is latency / launch bound rather than HBM bound.
#include "common.cuh"
#include "kernels.cuh"
#include "tinyaudio_b200.h"

namespace {

constexpr int ALN_WARPS = 8;

// z = bf16(o) * mask + resid[row % resid_rows]   (each term optional) for this lane's float4 vectors of one row
template <int MAXV>
__device__ __forceinline__ void aln_load_z(float4 (&z)[MAXV], int nv, int lane, long long row, int H, const bf16* o, const float* mask,
                                           const float* resid, long long rrow /* = row % resid_rows, hoisted: a 64-bit modulo per vector
                                           cost more than the loads */) {
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        if (i >= nv) break;
        const int c = (i * 32 + lane) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (o) {
            const uint2 u = *reinterpret_cast<const uint2*>(o + row * H + c);
            const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
            v = make_float4(a.x, a.y, b.x, b.y);
            if (mask) {
                const float4 m = *reinterpret_cast<const float4*>(mask + row * H + c);
                v.x *= m.x; v.y *= m.y; v.z *= m.z; v.w *= m.w;
            }
        }
        if (resid) {
            const float4 r = *reinterpret_cast<const float4*>(resid + rrow * H + c);
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        }
        z[i] = v;
    }
}

template <int MAXV>
__global__ void __launch_bounds__(ALN_WARPS * 32)
add_layernorm_fwd_kernel(const bf16* __restrict__ o, const float* __restrict__ mask, const float* __restrict__ resid, long long resid_rows,
                         const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ post_mask, long long post_rows,
                         float* __restrict__ y32, bf16* __restrict__ y16, float* __restrict__ stats, long long R, int H, float eps) {
    const int lane = threadIdx.x & 31;
    const int nv = H / 128;
    const long long warp0 = (long long)blockIdx.x * ALN_WARPS + (threadIdx.x >> 5), nwarps = (long long)gridDim.x * ALN_WARPS;
    for (long long row = warp0; row < R; row += nwarps) {
        float4 z[MAXV];
        const long long rrow = (resid && resid_rows != R) ? row % resid_rows : row;
        const long long prow = (post_mask && post_rows != R) ? row % post_rows : row;
        aln_load_z<MAXV>(z, nv, lane, row, H, o, mask, resid, rrow);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i)
            if (i < nv) s += (z[i].x + z[i].y) + (z[i].z + z[i].w);
        const float mean = warp_sum(s) / H;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i)
            if (i < nv) {
                const float a = z[i].x - mean, bb = z[i].y - mean, c = z[i].z - mean, d = z[i].w - mean;
                q += (a * a + bb * bb) + (c * c + d * d);
            }
        const float rstd = rsqrtf(warp_sum(q) / H + eps);
        if (lane == 0) {
            stats[2 * row] = mean;
            stats[2 * row + 1] = rstd;
        }
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            if (i >= nv) break;
            const int c = (i * 32 + lane) * 4;
            const float4 ww = *reinterpret_cast<const float4*>(w + c), bv = *reinterpret_cast<const float4*>(b + c);
            float4 y = make_float4((z[i].x - mean) * rstd * ww.x + bv.x, (z[i].y - mean) * rstd * ww.y + bv.y,
                                   (z[i].z - mean) * rstd * ww.z + bv.z, (z[i].w - mean) * rstd * ww.w + bv.w);
            if (post_mask) {
                const float4 m = *reinterpret_cast<const float4*>(post_mask + prow * H + c);
                y.x *= m.x; y.y *= m.y; y.z *= m.z; y.w *= m.w;
            }
            *reinterpret_cast<float4*>(y32 + row * H + c) = y;
            if (y16) {
                uint2 u;
                u.x = pack_bf16x2(y.x, y.y);
                u.y = pack_bf16x2(y.z, y.w);
                *reinterpret_cast<uint2*>(y16 + row * H + c) = u;
            }
        }
    }
}

// dy = (g32 + g16) * post_mask ;  xhat = (z - mean) rstd ;  g = dy w ;  dz = rstd (g - mean(g) - xhat mean(g xhat))
// d_o = bf16(dz * mask), d_resid (+)= dz, dw += sum_rows dy xhat, db += sum_rows dy
template <int MAXV>
__global__ void __launch_bounds__(ALN_WARPS * 32)
add_layernorm_bwd_kernel(const float* __restrict__ g32, const bf16* __restrict__ g16, const bf16* __restrict__ o,
                         const float* __restrict__ mask, const float* __restrict__ resid, long long resid_rows, const float* __restrict__ w,
                         const float* __restrict__ post_mask, long long post_rows, const float* __restrict__ stats, bf16* __restrict__ d_o,
                         float* __restrict__ d_resid, int d_resid_atomic, float* __restrict__ partial /* [gridDim.x][2 H] */, long long R, int H) {
    // [warps][2][H]: every warp accumulates its rows' dw / db contributions in ITS OWN shared-memory slice (each lane owns fixed columns,
    // so plain read-modify-write, no atomics).  In registers the two accumulators cost 80 of 198 registers and held the kernel to one
    // 8-warp block per SM (172 us per launch at 9 600 x 1 280, profiles/r02_c20_trace_qformer.txt)
    extern __shared__ float s_acc[];
    const int lane = threadIdx.x & 31;
    const int nv = H / 128;
    for (int i = threadIdx.x; i < ALN_WARPS * 2 * H; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();
    float* sw = s_acc + (threadIdx.x >> 5) * 2 * H;
    const long long warp0 = (long long)blockIdx.x * ALN_WARPS + (threadIdx.x >> 5), nwarps = (long long)gridDim.x * ALN_WARPS;
    for (long long row = warp0; row < R; row += nwarps) {
        float4 z[MAXV], dy[MAXV];
        const long long rrow = (resid && resid_rows != R) ? row % resid_rows : row;
        const long long prow = (post_mask && post_rows != R) ? row % post_rows : row;
        aln_load_z<MAXV>(z, nv, lane, row, H, o, mask, resid, rrow);
        const float mean = stats[2 * row], rstd = stats[2 * row + 1];
        float c1 = 0.f, c2 = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            if (i >= nv) break;
            const int c = (i * 32 + lane) * 4;
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (g32) g = *reinterpret_cast<const float4*>(g32 + row * H + c);
            if (g16) {
                const uint2 u = *reinterpret_cast<const uint2*>(g16 + row * H + c);
                const float2 a = unpack_bf16x2(u.x), bb = unpack_bf16x2(u.y);
                g.x += a.x; g.y += a.y; g.z += bb.x; g.w += bb.y;
            }
            if (post_mask) {
                const float4 m = *reinterpret_cast<const float4*>(post_mask + prow * H + c);
                g.x *= m.x; g.y *= m.y; g.z *= m.z; g.w *= m.w;
            }
            dy[i] = g;
            z[i] = make_float4((z[i].x - mean) * rstd, (z[i].y - mean) * rstd, (z[i].z - mean) * rstd, (z[i].w - mean) * rstd);   // xhat
            {
                float4 a = *reinterpret_cast<float4*>(sw + c), bsum = *reinterpret_cast<float4*>(sw + H + c);
                a.x += g.x * z[i].x; a.y += g.y * z[i].y; a.z += g.z * z[i].z; a.w += g.w * z[i].w;
                bsum.x += g.x; bsum.y += g.y; bsum.z += g.z; bsum.w += g.w;
                *reinterpret_cast<float4*>(sw + c) = a;
                *reinterpret_cast<float4*>(sw + H + c) = bsum;
            }
            const float4 ww = *reinterpret_cast<const float4*>(w + c);
            dy[i] = make_float4(g.x * ww.x, g.y * ww.y, g.z * ww.z, g.w * ww.w);
            c1 += (dy[i].x + dy[i].y) + (dy[i].z + dy[i].w);
            c2 += (dy[i].x * z[i].x + dy[i].y * z[i].y) + (dy[i].z * z[i].z + dy[i].w * z[i].w);
        }
        c1 = warp_sum(c1) / H;
        c2 = warp_sum(c2) / H;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            if (i >= nv) break;
            const int c = (i * 32 + lane) * 4;
            const float4 dz = make_float4(rstd * (dy[i].x - c1 - z[i].x * c2), rstd * (dy[i].y - c1 - z[i].y * c2),
                                          rstd * (dy[i].z - c1 - z[i].z * c2), rstd * (dy[i].w - c1 - z[i].w * c2));
            if (d_o) {
                float4 t = dz;
                if (mask) {
                    const float4 m = *reinterpret_cast<const float4*>(mask + row * H + c);
                    t.x *= m.x; t.y *= m.y; t.z *= m.z; t.w *= m.w;
                }
                uint2 u;
                u.x = pack_bf16x2(t.x, t.y);
                u.y = pack_bf16x2(t.z, t.w);
                *reinterpret_cast<uint2*>(d_o + row * H + c) = u;
            }
            if (d_resid) {
                float* p = d_resid + rrow * H + c;
                if (d_resid_atomic) {
                    atomicAdd(p, dz.x); atomicAdd(p + 1, dz.y); atomicAdd(p + 2, dz.z); atomicAdd(p + 3, dz.w);
                } else {
                    *reinterpret_cast<float4*>(p) = dz;
                }
            }
        }
    }
    __syncthreads();
    // block partials (the warps' slices summed in a fixed order), then aln_reduce_partials_kernel: deterministic dw / db, no atomics
    for (int i = threadIdx.x; i < 2 * H; i += blockDim.x) {
        float a = 0.f;
#pragma unroll
        for (int wv = 0; wv < ALN_WARPS; ++wv) a += s_acc[wv * 2 * H + i];
        partial[(long long)blockIdx.x * 2 * H + i] = a;
    }
}

__global__ void aln_reduce_partials_kernel(const float* __restrict__ partial, int n_blocks, int H, float* __restrict__ dw, float* __restrict__ db) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= 2 * H) return;
    float a = 0.f;
    for (int b = 0; b < n_blocks; ++b) a += partial[(long long)b * 2 * H + j];
    if (j < H) dw[j] = a;
    else db[j - H] = a;
}

// exact GELU (hidden_act = "gelu"): y = 0.5 x (1 + erf(x / sqrt 2)) on bf16 storage, fp32 math
__global__ void gelu_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, long long n8) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const uint4 u = reinterpret_cast<const uint4*>(x)[i];
        const uint32_t in[4] = {u.x, u.y, u.z, u.w};
        uint32_t out[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float2 v = unpack_bf16x2(in[t]);
            out[t] = pack_bf16x2(0.5f * v.x * (1.0f + erff(v.x * 0.70710678118654752f)), 0.5f * v.y * (1.0f + erff(v.y * 0.70710678118654752f)));
        }
        reinterpret_cast<uint4*>(y)[i] = make_uint4(out[0], out[1], out[2], out[3]);
    }
}
// dx = dy (Phi(x) + x phi(x))
__global__ void gelu_bwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, bf16* __restrict__ dx, long long n8) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const uint4 u = reinterpret_cast<const uint4*>(x)[i], g = reinterpret_cast<const uint4*>(dy)[i];
        const uint32_t in[4] = {u.x, u.y, u.z, u.w}, gin[4] = {g.x, g.y, g.z, g.w};
        uint32_t out[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float2 v = unpack_bf16x2(in[t]), d = unpack_bf16x2(gin[t]);
            const float a = 0.5f * (1.0f + erff(v.x * 0.70710678118654752f)) + v.x * 0.3989422804014327f * __expf(-0.5f * v.x * v.x);
            const float b = 0.5f * (1.0f + erff(v.y * 0.70710678118654752f)) + v.y * 0.3989422804014327f * __expf(-0.5f * v.y * v.y);
            out[t] = pack_bf16x2(d.x * a, d.y * b);
        }
        reinterpret_cast<uint4*>(dx)[i] = make_uint4(out[0], out[1], out[2], out[3]);
    }
}

// out[j] += sum over this block's rows of x[r, j]: 128 threads x 2 columns, blockIdx.y = chunk of 128 rows
__global__ void __launch_bounds__(128) colsum_bf16_kernel(const bf16* __restrict__ x, long long ld, float* __restrict__ out, long long R, int C) {
    const int c = (blockIdx.x * 128 + threadIdx.x) * 2;
    if (c >= C) return;
    const long long r0 = (long long)blockIdx.y * 128, r1 = min(R, r0 + 128);
    float a = 0.f, b = 0.f;
    for (long long r = r0; r < r1; ++r) {
        const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x + r * ld + c));
        a += v.x;
        b += v.y;
    }
    atomicAdd(&out[c], a);
    atomicAdd(&out[c + 1], b);
}

constexpr int ALN_BWD_MAX_BLOCKS = 148 * 2;
int aln_grid(long long R, int cap) {
    const long long blocks = (R + ALN_WARPS - 1) / ALN_WARPS;
    return (int)(blocks < cap ? blocks : cap);
}

}  // namespace

TA_API int ta_add_layernorm_fwd(const void* o, const float* mask, const float* resid, long long resid_rows, const float* w, const float* b,
                                const float* post_mask, long long post_rows, float* y32, void* y16, float* stats, long long rows, int H,
                                float eps, void* stream) {
    TA_REQUIRE(w && b && y32 && stats && (o || resid), "ta_add_layernorm_fwd: null pointer");
    TA_REQUIRE(H % 128 == 0 && H <= 2048, "ta_add_layernorm_fwd: H=%d must be a multiple of 128 and <= 2048", H);
    TA_REQUIRE(!resid || resid_rows > 0, "ta_add_layernorm_fwd: resid_rows must be positive");
    TA_REQUIRE(!post_mask || post_rows > 0, "ta_add_layernorm_fwd: post_rows must be positive");
    if (rows == 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bf16* ob = reinterpret_cast<const bf16*>(o);
    bf16* yb = reinterpret_cast<bf16*>(y16);
    if (H <= 1280)
        TA_KERNEL_LAUNCH(add_layernorm_fwd_kernel<10>, aln_grid(rows, 1 << 20), ALN_WARPS * 32, 0, st, ob, mask, resid, resid_rows, w, b, post_mask,
                         post_rows, y32, yb, stats, rows, H, eps);
    else
        TA_KERNEL_LAUNCH(add_layernorm_fwd_kernel<16>, aln_grid(rows, 1 << 20), ALN_WARPS * 32, 0, st, ob, mask, resid, resid_rows, w, b, post_mask,
                         post_rows, y32, yb, stats, rows, H, eps);
    return 0;
}

TA_API int ta_add_layernorm_bwd(const float* g32, const void* g16, const void* o, const float* mask, const float* resid, long long resid_rows,
                                const float* w, const float* post_mask, long long post_rows, const float* stats, void* d_o, float* d_resid,
                                float* dw, float* db, float* partial, long long rows, int H, void* stream) {
    TA_REQUIRE(w && stats && dw && db && partial && (g32 || g16) && (o || resid), "ta_add_layernorm_bwd: null pointer");
    TA_REQUIRE(H % 128 == 0 && H <= 2048, "ta_add_layernorm_bwd: H=%d must be a multiple of 128 and <= 2048", H);
    TA_REQUIRE(!d_o || o, "ta_add_layernorm_bwd: d_o without o");
    TA_REQUIRE(!d_resid || (resid && resid_rows > 0), "ta_add_layernorm_bwd: d_resid without resid");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int atomic = (d_resid && resid_rows < rows) ? 1 : 0;
    if (atomic) TA_CHECK_CUDA(cudaMemsetAsync(d_resid, 0, sizeof(float) * resid_rows * H, st));
    if (rows == 0) {
        TA_CHECK_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * H, st));
        TA_CHECK_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * H, st));
        return 0;
    }
    const int grid = aln_grid(rows, ALN_BWD_MAX_BLOCKS);
    const bf16* ob = reinterpret_cast<const bf16*>(o);
    const bf16* gb = reinterpret_cast<const bf16*>(g16);
    bf16* dob = reinterpret_cast<bf16*>(d_o);
    const size_t smem = sizeof(float) * ALN_WARPS * 2 * H;
    static bool attr_done = false;
    if (!attr_done) {
        TA_CHECK_CUDA(cudaFuncSetAttribute(add_layernorm_bwd_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float) * ALN_WARPS * 2 * 1280)));
        TA_CHECK_CUDA(cudaFuncSetAttribute(add_layernorm_bwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float) * ALN_WARPS * 2 * 2048)));
        attr_done = true;
    }
    if (H <= 1280)
        TA_KERNEL_LAUNCH(add_layernorm_bwd_kernel<10>, grid, ALN_WARPS * 32, smem, st, g32, gb, ob, mask, resid, resid_rows, w, post_mask,
                         post_rows, stats, dob, d_resid, atomic, partial, rows, H);
    else
        TA_KERNEL_LAUNCH(add_layernorm_bwd_kernel<16>, grid, ALN_WARPS * 32, smem, st, g32, gb, ob, mask, resid, resid_rows, w, post_mask,
                         post_rows, stats, dob, d_resid, atomic, partial, rows, H);
    TA_KERNEL_LAUNCH(aln_reduce_partials_kernel, (2 * H + 255) / 256, 256, 0, st, partial, grid, H, dw, db);
    return 0;
}

TA_API long long ta_add_layernorm_bwd_partial_floats(int H) { return (long long)ALN_BWD_MAX_BLOCKS * 2 * H; }

TA_API int ta_gelu_fwd_bf16(const void* x, void* y, long long n, void* stream) {
    TA_REQUIRE(x && y, "ta_gelu_fwd_bf16: null pointer");
    TA_REQUIRE(n % 8 == 0, "ta_gelu_fwd_bf16: n=%lld must be a multiple of 8", n);
    if (n == 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long n8 = n / 8;
    const unsigned grid = (unsigned)((n8 + 255) / 256 < 148 * 8 ? (n8 + 255) / 256 : 148 * 8);
    TA_KERNEL_LAUNCH(gelu_fwd_kernel, grid, 256, 0, st, reinterpret_cast<const bf16*>(x), reinterpret_cast<bf16*>(y), n8);
    return 0;
}

TA_API int ta_gelu_bwd_bf16(const void* x, const void* dy, void* dx, long long n, void* stream) {
    TA_REQUIRE(x && dy && dx, "ta_gelu_bwd_bf16: null pointer");
    TA_REQUIRE(n % 8 == 0, "ta_gelu_bwd_bf16: n=%lld must be a multiple of 8", n);
    if (n == 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long n8 = n / 8;
    const unsigned grid = (unsigned)((n8 + 255) / 256 < 148 * 8 ? (n8 + 255) / 256 : 148 * 8);
    TA_KERNEL_LAUNCH(gelu_bwd_kernel, grid, 256, 0, st, reinterpret_cast<const bf16*>(x), reinterpret_cast<const bf16*>(dy),
                     reinterpret_cast<bf16*>(dx), n8);
    return 0;
}

TA_API int ta_colsum_bf16(const void* x, long long ld, float* out, long long rows, int cols, void* stream) {
    TA_REQUIRE(x && out, "ta_colsum_bf16: null pointer");
    TA_REQUIRE(cols % 2 == 0 && ld % 2 == 0, "ta_colsum_bf16: cols=%d and ld=%lld must be even", cols, ld);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    TA_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * cols, st));
    if (rows == 0 || cols == 0) return 0;
    dim3 grid((unsigned)((cols / 2 + 127) / 128), (unsigned)((rows + 127) / 128));
    TA_KERNEL_LAUNCH(colsum_bf16_kernel, grid, 128, 0, st, reinterpret_cast<const bf16*>(x), ld, out, rows, cols);
    return 0;
}
