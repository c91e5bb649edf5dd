// How many warps per SM sub-partition does it take to saturate MUFU.EX2 with a softmax-like instruction mix?
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o sfu2 sfu2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float a, float b) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; }
// MODE 0: pure ex2 chains (ILP 8).  MODE 1: softmax mix on 64 register-resident inputs: fma, ex2, add, pack, 16-byte smem store
template <int MODE>
__global__ void k(float* out, int iters, float seed) {
    extern __shared__ uint4 sm[];
    float v[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = seed * (threadIdx.x + i) * 1e-6f - 1.0f;
    float l4[4] = {0, 0, 0, 0};
    const float sc = seed * 0.5f, nm = -seed;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 64; ++i) v[i] = ex2(v[i]) - 1.5f;
        } else {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                float e[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) { e[t] = ex2(fmaf(v[8 * g + t], sc, nm)); l4[t & 3] += e[t]; }
                uint4 u; u.x = pack(e[0], e[1]); u.y = pack(e[2], e[3]); u.z = pack(e[4], e[5]); u.w = pack(e[6], e[7]);
                sm[threadIdx.x * 8 + (g ^ (threadIdx.x & 7))] = u;
            }
            v[it & 63] += 1e-3f;
        }
    }
    float s = l4[0] + l4[1] + l4[2] + l4[3]; for (int i = 0; i < 64; ++i) s += v[i];
    if (s == 12345.678f) out[0] = s + sm[5].x;
}
template <int MODE> void run(const char* name, float* d, int threads) {
    const int iters = 2048, blocks = 148;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads, threads * 128>>>(d, 16, 1.f);
    cudaEventRecord(e0); k<MODE><<<blocks, threads, threads * 128>>>(d, iters, 1.f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double n = (double)blocks * threads * iters * 64;
    printf("%-14s warps/SMSP=%2d  %8.3f ms  %6.2f exp/clk/SM @1.9GHz\n", name, threads / 128, ms, n / (ms * 1e-3) / 148 / 1.9e9);
}
int main() {
    float* d; cudaMalloc(&d, 4);
    for (int t : {128, 256, 512, 1024}) run<0>("ex2 only", d, t);
    for (int t : {128, 256, 512, 1024}) run<1>("softmax mix", d, t);
    return 0;
}
