"""Launch the dominant kernels of the step once each at their production shapes (for `ncu --set full` captures)."""
import sys
import torch

sys.path.insert(0, ".")
from tiny_audio_b200 import lib as L

BF16, F32 = torch.bfloat16, torch.float32
which = sys.argv[1:] or ["gemm", "attn_enc", "attn_lm"]
dev = "cuda"
lib = L.load()
torch.manual_seed(0)
if "gemm" in which:
    M, N, K = 48000, 5120, 1280
    a = torch.randn(M, K, device=dev, dtype=BF16)
    w = torch.randn(N, K, device=dev, dtype=BF16) * 0.03
    bias = torch.zeros(N, device=dev, dtype=F32)
    out = torch.empty(M, N, device=dev, dtype=BF16)
    for _ in range(3):
        L.gemm(a, w, epi=L.EPI_BF16_GELU, bias=bias, out=out)
    # LM shape: gate/up with SwiGLU epilogue
    M, N, K = 14848, 6144, 1024
    a2 = torch.randn(M, K, device=dev, dtype=BF16)
    w2 = torch.randn(N, K, device=dev, dtype=BF16) * 0.03
    gu = torch.empty(M, N, device=dev, dtype=BF16)
    for _ in range(3):
        L.gemm(a2, w2, epi=L.EPI_SWIGLU, out2=gu)
if "attn_enc" in which:
    B, S, H, hd = 32, 1500, 20, 64
    qkv = torch.randn(B, S, 3 * H * hd, device=dev, dtype=BF16)
    o = torch.empty(B, S, H * hd, device=dev, dtype=BF16)
    for _ in range(2):
        L.check(lib.ta_attn_fwd(L.ptr(qkv), L.ptr(qkv[:, :, H * hd:]), L.ptr(qkv[:, :, 2 * H * hd:]), L.ptr(o), None, B, S, H, H, hd,
                                3 * H * hd, 3 * H * hd, 3 * H * hd, H * hd, 0, hd ** -0.5, L.stream_ptr()))
if "attn_lm" in which:
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
    for _ in range(2):
        L.check(lib.ta_attn_fwd(L.ptr(q), L.ptr(k), L.ptr(v), L.ptr(o), L.ptr(lse), B, S, Hq, Hkv, hd, Hq * hd, Hkv * hd, Hkv * hd,
                                Hq * hd, 1, hd ** -0.5, L.stream_ptr()))
        L.check(lib.ta_attn_bwd(L.ptr(q), L.ptr(k), L.ptr(v), L.ptr(o), L.ptr(do), L.ptr(lse), L.ptr(dsum), L.ptr(dq), L.ptr(dk),
                                L.ptr(dv), B, S, Hq, Hkv, hd, Hq * hd, Hkv * hd, Hkv * hd, Hq * hd, Hq * hd, Hq * hd, Hkv * hd,
                                Hkv * hd, 1, hd ** -0.5, L.stream_ptr()))
torch.cuda.synchronize()
print("done")
