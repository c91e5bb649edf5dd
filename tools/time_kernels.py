"""CUDA-event timings of the dominant kernels at production shapes, A/B over the kernel variants."""
import sys
import torch
sys.path.insert(0, ".")
from tiny_audio_b200 import lib as L

BF16, F32 = torch.bfloat16, torch.float32
lib = L.load()
dev = "cuda"


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


shapes = [("enc fc1 gelu", 48000, 5120, 1280, L.EPI_BF16_GELU), ("enc qkv", 48000, 3840, 1280, L.EPI_BF16),
          ("enc fc2 resid", 48000, 1280, 5120, L.EPI_BF16_RESID), ("enc o resid", 48000, 1280, 1280, L.EPI_BF16_RESID),
          ("lm gate_up swiglu", 14848, 6144, 1024, L.EPI_SWIGLU), ("lm down f32resid", 14848, 1024, 3072, L.EPI_F32_RESID),
          ("lm qkv", 14848, 4096, 1024, L.EPI_BF16), ("lm head", 2080, 151936, 1024, L.EPI_BF16),
          ("lm dhead", 2080, 1024, 151936, L.EPI_BF16)]
for name, M, N, K, epi in shapes:
    a = torch.randn(M, K, device=dev, dtype=BF16)
    w = torch.randn(N, K, device=dev, dtype=BF16) * 0.03
    bias = torch.zeros(N, device=dev, dtype=F32) if epi in (L.EPI_BF16, L.EPI_BF16_GELU, L.EPI_BF16_RESID) else None
    kw = {}
    if epi == L.EPI_BF16_RESID:
        kw["resid"] = torch.zeros(M, N, device=dev, dtype=BF16)
    if epi == L.EPI_F32_RESID:
        kw["resid"] = torch.zeros(M, N, device=dev, dtype=F32)
    if epi == L.EPI_SWIGLU:
        kw["out2"] = torch.empty(M, N, device=dev, dtype=BF16)
    out = L.gemm(a, w, epi=epi, bias=bias, **kw)
    res = []
    for pair in (0, 1):
        for bn in (128, 256):
            if N % bn:
                continue
            lib.ta_gemm_set_cta_pair(pair)
            lib.ta_gemm_set_tile_n(bn)
            t = timeit(lambda: L.gemm(a, w, epi=epi, bias=bias, out=out, **kw))
            res.append(f"pair{pair}/bn{bn}: {t:.3f} ms {2.0 * M * N * K / t / 1e9:.0f} TF/s")
    tt = timeit(lambda: torch.matmul(a, w.t()))
    print(f"{name:18s} M={M} N={N} K={K} | " + " | ".join(res) + f" | cuBLAS {tt:.3f} ms {2.0 * M * N * K / tt / 1e9:.0f} TF/s", flush=True)
    del a, w, out
lib.ta_gemm_set_tile_n(0)
lib.ta_gemm_set_cta_pair(1)

B, S, H, hd = 32, 1500, 20, 64
qkv = torch.randn(B, S, 3 * H * hd, device=dev, dtype=BF16)
o = torch.empty(B, S, H * hd, device=dev, dtype=BF16)
for tc in (0, 1, 2, 3, 4, 5):
    lib.ta_attn_set_tc(tc)
    t = timeit(lambda: L.check(lib.ta_attn_fwd(L.ptr(qkv), L.ptr(qkv[:, :, H * hd:]), L.ptr(qkv[:, :, 2 * H * hd:]), L.ptr(o), None, B, S, H,
                                               H, hd, 3 * H * hd, 3 * H * hd, 3 * H * hd, H * hd, 0, hd ** -0.5, L.stream_ptr())), reps=5)
    print(f"enc attention fwd tc={tc}: {t:.3f} ms  {4.0 * B * H * S * S * hd / t / 1e9:.0f} TF/s", flush=True)
lib.ta_attn_set_tc(2)

# decoder attention fwd + bwd (causal GQA, hd 128)
B, S, Hq, Hkv, hd = 32, 464, 16, 8, 128
q = torch.randn(B, S, Hq * hd, device=dev, dtype=BF16)
k = torch.randn(B, S, Hkv * hd, device=dev, dtype=BF16)
v = torch.randn(B, S, Hkv * hd, device=dev, dtype=BF16)
do = torch.randn(B, S, Hq * hd, device=dev, dtype=BF16)
o = torch.empty_like(q)
lse = torch.empty(B, Hq, S, device=dev, dtype=F32)
dsum = torch.empty_like(lse)
dq = torch.empty(B, S, Hq * hd, device=dev, dtype=F32)
dk, dv = torch.empty_like(k), torch.empty_like(v)
for tc in (0, 1):
    lib.ta_attn_set_tc(tc)
    tf = timeit(lambda: L.check(lib.ta_attn_fwd(L.ptr(q), L.ptr(k), L.ptr(v), L.ptr(o), L.ptr(lse), B, S, Hq, Hkv, hd, Hq * hd, Hkv * hd,
                                                Hkv * hd, Hq * hd, 1, hd ** -0.5, L.stream_ptr())), reps=5)
    tb = timeit(lambda: L.check(lib.ta_attn_bwd(L.ptr(q), L.ptr(k), L.ptr(v), L.ptr(o), L.ptr(do), L.ptr(lse), L.ptr(dsum), L.ptr(dq),
                                                L.ptr(dk), L.ptr(dv), B, S, Hq, Hkv, hd, Hq * hd, Hkv * hd, Hkv * hd, Hq * hd, Hq * hd,
                                                Hq * hd, Hkv * hd, Hkv * hd, 1, hd ** -0.5, L.stream_ptr())), reps=5)
    print(f"lm attention tc={tc}: fwd {tf:.3f} ms  bwd (prep+memset+main) {tb:.3f} ms", flush=True)
lib.ta_attn_set_tc(2)

lib.ta_debug_set(1, 1)
tb = timeit(lambda: L.check(lib.ta_attn_bwd(L.ptr(q), L.ptr(k), L.ptr(v), L.ptr(o), L.ptr(do), L.ptr(lse), L.ptr(dsum), L.ptr(dq),
                                            L.ptr(dk), L.ptr(dv), B, S, Hq, Hkv, hd, Hq * hd, Hkv * hd, Hkv * hd, Hq * hd, Hq * hd,
                                            Hq * hd, Hkv * hd, Hkv * hd, 1, hd ** -0.5, L.stream_ptr())), reps=5)
print(f"lm attention bwd WITHOUT dQ atomics (experiment): {tb:.3f} ms", flush=True)
lib.ta_debug_set(1, 0)

# log-mel front end: 32 x 30 s clips
import ctypes as C
B, Ls = 32, 480000
wave = 0.1 * torch.randn(B, Ls, device=dev)
n = C.c_longlong()
L.check(lib.ta_logmel_workspace_floats(B, Ls, C.byref(n)))
ws = torch.empty(n.value, device=dev, dtype=F32)
im2 = torch.empty(B * (Ls // 160), 384, device=dev, dtype=BF16)
t = timeit(lambda: L.check(lib.ta_logmel_fwd(L.ptr(wave), wave.stride(0), B, Ls, L.ptr(ws), None, L.ptr(im2), L.stream_ptr())), reps=10)
print(f"log-mel 32 x 30 s (power + finalize + im2col): {t:.3f} ms  ({B * (4 * Ls + 4 * 128 * (Ls // 160)) / t / 1e6:.0f} GB/s algorithmic)", flush=True)
