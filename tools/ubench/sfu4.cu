// Which pipe do the fp32 -> bf16x2 conversions use, and how does one warp's instruction stream interleave with MUFU.EX2?
//   MODE 0: ex2 only (ILP 16)                       MODE 1: F2FP (cvt.rn.bf16x2.f32) only
//   MODE 2: ex2 + F2FP 2:1 (the softmax ratio)      MODE 3: ex2 + integer pack (2 adds + prmt per pair) 2:1
//   MODE 4: ex2 + 3 independent FFMA per ex2        MODE 5: 3 FFMA per "slot" only (no ex2)
// For each mode: 1, 2 and 4 warps per SM sub-partition.  Reports warp-instructions of the MAIN op per clock and sub-partition,
// using the SM clock (clock64), not wall time.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o sfu4 sfu4.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t cvt2(float a, float b) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; }
__device__ __forceinline__ uint32_t ipack(float lo, float hi) {
    const uint32_t a = __float_as_uint(lo) + 0x8000u, b = __float_as_uint(hi) + 0x8000u;
    return __byte_perm(a, b, 0x7632);
}
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters, float seed) {
    float v[16], w[16];
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) { v[i] = seed * (threadIdx.x + i) * 1e-6f - 1.0f; w[i] = v[i] * 0.5f; }
    const float sc = seed * 0.999f, nm = -seed * 1e-3f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            if (MODE == 0) { v[i] = ex2(v[i]) - 1.5f; v[i + 1] = ex2(v[i + 1]) - 1.5f; }
            if (MODE == 1) { acc += cvt2(v[i], v[i + 1]); v[i] += 1.0f; }
            if (MODE == 2) { const float a = ex2(v[i]), b = ex2(v[i + 1]); acc += cvt2(a, b); v[i] = a - 1.5f; v[i + 1] = b - 1.5f; }
            if (MODE == 3) { const float a = ex2(v[i]), b = ex2(v[i + 1]); acc += ipack(a, b); v[i] = a - 1.5f; v[i + 1] = b - 1.5f; }
            if (MODE == 4) {
                v[i] = ex2(v[i]) - 1.5f; v[i + 1] = ex2(v[i + 1]) - 1.5f;
                w[i] = fmaf(w[i], sc, nm); w[i + 1] = fmaf(w[i + 1], sc, nm); w[i] = fmaf(w[i], sc, nm); w[i + 1] = fmaf(w[i + 1], sc, nm);
                w[i] = fmaf(w[i], sc, nm); w[i + 1] = fmaf(w[i + 1], sc, nm);
            }
            if (MODE == 5) {
                w[i] = fmaf(w[i], sc, nm); w[i + 1] = fmaf(w[i + 1], sc, nm); w[i] = fmaf(w[i], sc, nm); w[i + 1] = fmaf(w[i + 1], sc, nm);
                w[i] = fmaf(w[i], sc, nm); w[i + 1] = fmaf(w[i + 1], sc, nm);
            }
        }
    }
    const long long t1 = clock64();
    float s = 0; for (int i = 0; i < 16; ++i) s += v[i] + w[i];
    if (s == 12345.678f || acc == 0x12345u) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int MODE> void run(const char* name, float* d, long long* c, int threads) {
    const int iters = 4096, blocks = 148;
    k<MODE><<<blocks, threads>>>(d, c, 16, 1.f);
    k<MODE><<<blocks, threads>>>(d, c, iters, 1.f);
    long long cyc; cudaMemcpy(&cyc, c, 8, cudaMemcpyDeviceToHost);
    const double main_ops = (double)iters * 16 * (MODE == 1 ? 0.5 : 1.0) * (threads / 128);   // warp-instructions of the main op per SMSP
    printf("%-34s warps/SMSP=%d  %9lld clk  main op: %.3f warp-instr/clk/SMSP = one per %.2f clk\n", name, threads / 128, cyc, main_ops / cyc, cyc / main_ops);
}
int main() {
    float* d; long long* c; cudaMalloc(&d, 4); cudaMalloc(&c, 8);
    for (int t : {128, 256, 512}) run<0>("ex2 only", d, c, t);
    for (int t : {128, 256, 512}) run<1>("cvt.rn.bf16x2 only (per cvt)", d, c, t);
    for (int t : {128, 256, 512}) run<2>("ex2 + cvt 2:1 (per ex2)", d, c, t);
    for (int t : {128, 256, 512}) run<3>("ex2 + integer pack 2:1 (per ex2)", d, c, t);
    for (int t : {128, 256, 512}) run<4>("ex2 + 3 FFMA (per ex2)", d, c, t);
    for (int t : {128, 256, 512}) run<5>("3 FFMA only (per slot)", d, c, t);
    return 0;
}
