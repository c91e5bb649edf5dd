// Common device/host helpers for the tiny-audio B200 hot path (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#ifndef TA_API
#define TA_API extern "C" __attribute__((visibility("default")))
#endif

// ----------------------------------------------------------------------------------------------
// error plumbing: every C-ABI entry returns 0 or a negative code; the message is kept thread-local
// ----------------------------------------------------------------------------------------------
void ta_set_error(const char* fmt, ...);
#define TA_CHECK_CUDA(expr)                                                                  \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            ta_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return -(int)_e - 1000;                                                          \
        }                                                                                    \
    } while (0)
#define TA_REQUIRE(cond, ...)                                                                \
    do {                                                                                     \
        if (!(cond)) {                                                                       \
            ta_set_error(__VA_ARGS__);                                                       \
            return -1;                                                                       \
        }                                                                                    \
    } while (0)
extern unsigned long long g_ta_launches;   // kernels launched by this library (bench.py reports it as gpu_launches)
#define TA_LAUNCH_CHECK()                          \
    do {                                           \
        ++g_ta_launches;                           \
        TA_CHECK_CUDA(cudaGetLastError());         \
    } while (0)

typedef __nv_bfloat16 bf16;

// ----------------------------------------------------------------------------------------------
// programmatic dependent launch (PDL): a kernel launched with launch_pdl() may start while its predecessor in the stream
// is still running; it must call pdl_wait() before touching anything the predecessor writes (or reads, if it overwrites
// it).  Everything before pdl_wait() -- index setup, prefetch of FROZEN weights -- overlaps the predecessor's tail.
// ----------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
#endif

// PDL for the TRAINING towers is compile-gated (make PDL=1 -> -DTA_PDL): in a default build TA_PDL_ENTRY() is empty and
// TA_KERNEL_LAUNCH is the plain <<< >>> launch, i.e. the binary is the one the parity tests and profiles were run on.  In a PDL
// build every hot-loop kernel executes griddepcontrol.wait + launch_dependents at TA_PDL_ENTRY() (before its first global-memory
// access; wait first, so at most ONE dependent grid is ever parked behind a running one) and is launched with the
// programmatic-stream-serialization attribute while ta_set_pdl(1) is in effect (runtime A/B inside the PDL build).
// Status: compiles, NOT yet run on hardware.
#ifdef TA_PDL
extern int g_ta_pdl;
#define TA_PDL_ENTRY()              \
    do {                            \
        pdl_wait();                 \
        pdl_launch_dependents();    \
    } while (0)
#define TA_KERNEL_LAUNCH(kern, grid, block, smem, st, ...)                                  \
    do {                                                                                    \
        if (g_ta_pdl) {                                                                     \
            TA_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(block), smem, st, __VA_ARGS__)); \
            ++g_ta_launches;                                                                \
        } else {                                                                            \
            kern<<<grid, block, smem, st>>>(__VA_ARGS__);                                   \
            TA_LAUNCH_CHECK();                                                              \
        }                                                                                   \
    } while (0)
#else
#define TA_PDL_ENTRY() \
    do {               \
    } while (0)
#define TA_KERNEL_LAUNCH(kern, grid, block, smem, st, ...) \
    do {                                                   \
        kern<<<grid, block, smem, st>>>(__VA_ARGS__);      \
        TA_LAUNCH_CHECK();                                 \
    } while (0)
#endif

__host__ __device__ inline int64_t ceil_div_i64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ----------------------------------------------------------------------------------------------
// small math
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
    return cdf + x * pdf;
}
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
// exact-erf GELU for GEMM epilogues in 15 instructions (2 MUFU): Abramowitz-Stegun 7.1.26, |erf error| <= 1.5e-7.
//   a = |x| sqrt(log2(e)/2);  t = 1/(1 + p' a);  q = 0.5 (1 - erf(|x|/sqrt2)) = t P(t) 2^(-a^2)   (0.5 folded into P)
//   gelu(x) = max(x, 0) - |x| q            (both tails keep their relative accuracy)
__device__ __forceinline__ float gelu_erf_fast(float x) {
    const float a = fabsf(x) * 0.84932180028801904272f;                 // sqrt(log2(e) / 2)
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(a, 0.2727374809f, 1.0f)));   // p / sqrt(log2 e)
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-a * a));
    float p = fmaf(0.5307027145f, t, -0.7265760135f);
    p = fmaf(p, t, 0.7107068705f);
    p = fmaf(p, t, -0.142248368f);
    p = fmaf(p, t, 0.127414796f);
    const float q = p * t * e;
    return fmaxf(x, 0.0f) - fabsf(x) * q;
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi);
// Round two floats to bf16 precision through ONE packed conversion (F2FP.BF16.F32.PACK_AB) + two integer ops.  A scalar
// __float2bfloat16_rn is an F2F.BF16.F32, which issues on the XU pipe next to MUFU (16 / clk / SM): the SwiGLU-backward epilogue spent
// as much XU time on its two roundings per element as on the sigmoid's ex2 + rcp (ncu: XU pipe 50 % of the kernel, F2F on the hot list).
__device__ __forceinline__ void bf16_round_pair(float& a, float& b) {
    const uint32_t p = pack_bf16x2(a, b);
    a = __uint_as_float(p << 16);
    b = __uint_as_float(p & 0xffff0000u);
}
template <int N>
__device__ __forceinline__ void bf16_round_all(float (&v)[N]) {
    static_assert(N % 2 == 0, "pairs");
#pragma unroll
    for (int i = 0; i < N; i += 2) bf16_round_pair(v[i], v[i + 1]);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(v);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// block-wide sum for blockDim.x <= 1024 (result broadcast to all threads)
__device__ __forceinline__ float block_sum(float v, float* red /* >= 33 floats */) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
    if (w == 0) {
        t = warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}
__device__ __forceinline__ float block_max(float v, float* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float t = (threadIdx.x < nw) ? red[threadIdx.x] : -INFINITY;
    if (w == 0) {
        t = warp_max(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

// ----------------------------------------------------------------------------------------------
// PTX: shared-address conversion, mbarrier, TMA, tcgen05
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Waiting on an mbarrier.  TA_MBAR_WAIT_MODE (compile time; A/B builds through `make VARIANT=... EXTRA=-DTA_MBAR_WAIT_MODE=n`):
//   0: try_wait with a suspend-time hint -- the waiting thread is parked by the hardware instead of spinning through SYNCS / BRA,
//      which would steal issue slots from the compute warps of the same SM sub-partition (ncu on the encoder attention, round 1:
//      30 % of all executed warp instructions were such spins without the hint);
//   1: try_wait without a hint (the system-dependent default time limit, then the loop re-issues it);
//   2: test_wait in a tight loop (never parks).
// Round 2's kernel timeline (tools/attn_trace.py) measured ~1 000 clk between an arrive and the parked waiter's next instruction
// in mode 0 -- on the critical path of every (softmax -> MMA -> softmax) hand-over.
#ifndef TA_MBAR_SUSPEND_HINT
#define TA_MBAR_SUSPEND_HINT 0x989680
#endif
#ifndef TA_MBAR_WAIT_MODE
#define TA_MBAR_WAIT_MODE 0
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#if TA_MBAR_WAIT_MODE == 0
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(TA_MBAR_SUSPEND_HINT)
        : "memory");
#elif TA_MBAR_WAIT_MODE == 1
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
#else
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
#endif
}

__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// 2D tiled load, completes on an mbarrier (transaction bytes)
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// warm L2 with the box at {c0, c1} (no shared-memory destination, nothing to wait on)
__device__ __forceinline__ void tma_prefetch_l2_2d(const void* tmap, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(tmap), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void* tmap, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
// same box, but the shared-memory tile is ADDED to global memory (fp32 tensor map): split-K partial sums reduced in L2
__device__ __forceinline__ void tma_reduce_add_2d(const void* tmap, const void* smem_src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(NCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]   (kind::f16 covers bf16/fp16 inputs with fp32 accumulate)
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 32 columns of fp32: thread i of the warp receives lane (base+i), columns [col, col+32)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled operand tile (rows of 64 bf16 = 128 B; 8-row groups 1024 B apart).
//   bits [0,14)  start address >> 4        bits [16,30) leading byte offset >> 4 (unused for SW128 K-major)
//   bits [32,46) stride byte offset >> 4   bits [46,48) version = 1 (Blackwell)
//   bits [61,64) layout type: 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc_sw128_kmajor(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor, kind::f16: D=f32 (bits 4-5 = 1), A=B=bf16 (bits 7-9 / 10-12 = 1), both K-major,
// N>>3 at bit 17, M>>4 at bit 24
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n"
        ".reg .pred P;\n"
        "elect.sync _|P, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, P;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// cp.async (16 B) with zero fill when pred is false
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool pred) {
    const int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
