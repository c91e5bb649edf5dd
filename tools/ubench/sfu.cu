// Micro-benchmark: MUFU.EX2 vs F2FP (bf16x2 pack) vs a polynomial exp2 on the FMA pipe -- which pipes do they share on sm_100a?
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o sfu sfu.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float a, float b) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; }
// exp2 for x <= 0 (softmax domain), degree-3 minimax on [0,1) after range reduction with the magic-number trick
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -126.0f);
    const float t = x + 12582912.0f;              // 1.5 * 2^23: low mantissa bits hold round(x)
    const float f = x - (t - 12582912.0f);        // f in [-0.5, 0.5]
    float p = fmaf(f, 0.0555041086648216f, 0.2402264923172690f);
    p = fmaf(p, f, 0.6931471805599453f);
    p = fmaf(p, f, 1.0f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
template <int MODE>
__global__ void k(float* out, int iters, float seed) {
    float a[8]; uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed * (threadIdx.x + i) * 1e-6f - 1.0f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = ex2(a[i]) - 1.5f;
            if (MODE == 1) { acc ^= pack(a[i], a[(i + 1) & 7]); a[i] += 1.0f; }
            if (MODE == 2) { a[i] = ex2(a[i]) - 1.5f; if (i & 1) acc ^= pack(a[i], a[i - 1]); }
            if (MODE == 3) a[i] = ex2_poly(a[i]) - 1.5f;
            if (MODE == 4) a[i] = ((i & 3) == 3 ? ex2_poly(a[i]) : ex2(a[i])) - 1.5f;      // 25 % emulated
            if (MODE == 5) a[i] = ((i & 1) ? ex2_poly(a[i]) : ex2(a[i])) - 1.5f;            // 50 % emulated
            if (MODE == 6) { a[i] = ((i & 3) == 3 ? ex2_poly(a[i]) : ex2(a[i])) - 1.5f; if (i & 1) acc ^= pack(a[i], a[i - 1]); }
        }
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 12345.678f || acc == 0x12345u) out[0] = s + acc;
}
template <int MODE> void run(const char* name, float* d, double per_iter) {
    const int iters = 4096, blocks = 148 * 8, threads = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(d, 64, 1.f);
    cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(d, iters, 1.f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double n = (double)blocks * threads * iters * per_iter;
    printf("%-28s %8.3f ms  %8.2f Gelem/s  = %6.2f elem/clk/SM @1.9GHz (%.2f @1.6)\n", name, ms, n / ms * 1e-6, n / (ms * 1e-3) / 148 / 1.9e9, n / (ms * 1e-3) / 148 / 1.6e9);
}
int main() {
    float* d; cudaMalloc(&d, 4);
    run<0>("ex2 only", d, 8); run<1>("pack only (per pack)", d, 8); run<2>("ex2 + pack/2 (per ex2)", d, 8); run<3>("poly only", d, 8);
    run<4>("75% ex2 + 25% poly", d, 8); run<5>("50% ex2 + 50% poly", d, 8); run<6>("75/25 + pack/2", d, 8);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0); printf("nominal clock %d kHz\n", clk);
    return 0;
}
