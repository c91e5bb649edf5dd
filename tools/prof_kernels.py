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
if "down" in which:
    M, D, Fd = 14848, 1024, 3072
    hh = torch.randn(M, Fd, device=dev, dtype=BF16)
    wd = torch.randn(D, Fd, device=dev, dtype=BF16) * 0.03
    resid = torch.zeros(M, D, device=dev, dtype=F32)
    yy = torch.empty(M, D, device=dev, dtype=F32)
    for _ in range(3):
        L.gemm(hh, wd, epi=L.EPI_F32_RESID, resid=resid, out=yy)
    torch.cuda.synchronize()
if "swiglu_bwd" in which:
    M, D, Fd = 14848, 1024, 3072
    gu = torch.randn(M, 2 * Fd, device=dev, dtype=BF16)
    dy = torch.randn(M, D, device=dev, dtype=BF16)
    wd_t = torch.randn(Fd, D, device=dev, dtype=BF16) * 0.03
    dgu = torch.empty(M, 2 * Fd, device=dev, dtype=BF16)
    for _ in range(3):
        L.gemm(dy, wd_t, epi=L.EPI_SWIGLU_BWD, aux=gu, out=dgu)
    torch.cuda.synchronize()
if "gemm_cmp" in which:
    # this library's pair GEMM next to cuBLAS on the encoder fc1 shape and on 8192^3 (what tile / cluster shape does the library pick?)
    for (M, N, K) in ((48000, 5120, 1280), (8192, 8192, 8192)):
        a = torch.randn(M, K, device=dev, dtype=BF16)
        w = torch.randn(N, K, device=dev, dtype=BF16) * 0.03
        out = torch.empty(M, N, device=dev, dtype=BF16)
        for _ in range(2):
            L.gemm(a, w, epi=L.EPI_BF16, out=out)
            torch.matmul(a, w.t(), out=out)
        torch.cuda.synchronize()
        del a, w, out
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

# ---- HBM-bound kernels at production shapes (run with:  python tools/prof_kernels.py hbm) ----
if "hbm" in which:
    import ctypes as C
    from tiny_audio_b200.engine import FusedClipAdamW
    torch.manual_seed(0)
    # log-mel: 32 x 30 s clips
    B, Ls = 32, 480000
    wave = 0.1 * torch.randn(B, Ls, device=dev)
    n = C.c_longlong()
    L.check(lib.ta_logmel_workspace_floats(B, Ls, C.byref(n)))
    ws = torch.empty(n.value, device=dev, dtype=F32)
    im2 = torch.empty(B * (Ls // 160), 384, device=dev, dtype=BF16)
    for _ in range(2):
        L.check(lib.ta_logmel_fwd(L.ptr(wave), wave.stride(0), B, Ls, L.ptr(ws), None, L.ptr(im2), L.stream_ptr()))
    # encoder LayerNorm: 48000 x 1280 bf16
    rows, D = 48000, 1280
    x = torch.randn(rows, D, device=dev, dtype=BF16)
    w = torch.ones(D, device=dev, dtype=F32)
    y = torch.empty_like(x)
    for _ in range(2):
        L.check(lib.ta_layernorm_bf16(L.ptr(x), L.ptr(w), L.ptr(w), L.ptr(y), rows, D, 1e-5, L.stream_ptr()))
    # decoder RMSNorm fwd / bwd: 14848 x 1024 fp32 residual stream
    rows, D = 14848, 1024
    xf = torch.randn(rows, D, device=dev, dtype=F32)
    w1 = torch.ones(D, device=dev, dtype=F32)
    yb = torch.empty(rows, D, device=dev, dtype=BF16)
    dx = torch.zeros(rows, D, device=dev, dtype=F32)
    dw = torch.zeros(D, device=dev, dtype=F32)
    for _ in range(2):
        L.check(lib.ta_rmsnorm_f32(L.ptr(xf), L.ptr(w1), L.ptr(yb), None, rows, D, 1e-6, L.stream_ptr()))
        L.check(lib.ta_rmsnorm_f32_bwd(L.ptr(yb), L.ptr(xf), L.ptr(w1), L.ptr(dx), None, rows, D, 1e-6, 1, L.stream_ptr()))
        L.check(lib.ta_rmsnorm_dw(L.ptr(yb), L.ptr(xf), None, rows, D, 1e-6, L.ptr(dw), L.stream_ptr()))
    # CE on the labelled rows: 2080 x 152064 bf16 logits
    R, V, Vp = 2080, 151936, 152064
    lg = torch.randn(R, Vp, device=dev, dtype=BF16)
    tg = torch.randint(0, V, (R,), device=dev, dtype=torch.int32)
    loss = torch.zeros(1, device=dev, dtype=F32)
    for _ in range(2):
        L.check(lib.ta_ce_fwd_bwd(L.ptr(lg), Vp, L.ptr(tg), R, V, Vp, 1.0 / R, L.ptr(loss), None, 1, L.stream_ptr()))
    # clip + AdamW over the 12.6 M projector parameters (H = 2048)
    ps = [torch.randn(2048, 5120, device=dev), torch.randn(1024, 2048, device=dev)]
    opt = FusedClipAdamW(ps, lr=1e-3)
    gs = [torch.randn_like(p) for p in ps]
    for _ in range(2):
        opt.step(gs)
if "decode" in which:
    # decode-path kernels: tied lm_head for 32 sequences (HBM-bound weight stream), gate/up SwiGLU, split-K o_proj, attention over 433 cached keys
    M = 32
    xn = torch.randn(M, 1024, device=dev, dtype=BF16)
    emb = torch.randn(152064, 1024, device=dev, dtype=BF16) * 0.03
    logits = torch.empty(M, 152064, device=dev, dtype=BF16)
    wgu = torch.randn(6144, 1024, device=dev, dtype=BF16) * 0.03
    h = torch.empty(M, 3072, device=dev, dtype=BF16)
    wo = torch.randn(1024, 2048, device=dev, dtype=BF16) * 0.03
    att = torch.randn(M, 2048, device=dev, dtype=BF16)
    part = torch.empty(4, M, 1024, device=dev, dtype=F32)
    Hq, Hkv, hd, n_keys = 16, 8, 128, 433
    q = torch.randn(M, Hq * hd, device=dev, dtype=BF16)
    kc = torch.randn(M, 512, Hkv * hd, device=dev, dtype=BF16)
    vc = torch.randn(M, 512, Hkv * hd, device=dev, dtype=BF16)
    o = torch.empty(M, Hq * hd, device=dev, dtype=BF16)
    pos = torch.tensor([n_keys - 1], device=dev, dtype=torch.int32)
    for _ in range(2):
        L.check(lib.ta_skinny_gemm_bf16(L.ptr(xn), 1024, L.ptr(emb), 1024, M, 152064, 1024, L.SKINNY_BF16, L.ptr(logits), 152064, None, 1, L.stream_ptr()))
        L.check(lib.ta_skinny_gemm_bf16(L.ptr(xn), 1024, L.ptr(wgu), 1024, M, 6144, 1024, L.SKINNY_SWIGLU, L.ptr(h), 3072, None, 1, L.stream_ptr()))
        L.check(lib.ta_skinny_gemm_bf16(L.ptr(att), 2048, L.ptr(wo), 2048, M, 1024, 2048, L.SKINNY_PARTIAL, L.ptr(part), 1024, None, 4, L.stream_ptr()))
        L.check(lib.ta_decode_attn(L.ptr(q), L.ptr(kc), L.ptr(vc), L.ptr(o), Hq * hd, L.ptr(pos), M, Hq, Hkv, 512, hd ** -0.5, L.stream_ptr()))
if "wgrad" in which:
    # weight-gradient GEMM (TN form): d(W_gate_up) [6144, 1024] = dY^T X over 14848 tokens
    dy = torch.randn(14848, 6144, device=dev, dtype=BF16)
    xx = torch.randn(14848, 1024, device=dev, dtype=BF16)
    out = torch.empty(6144, 1024, device=dev, dtype=F32)
    for _ in range(3):
        L.gemm_tn(dy, xx, out=out)
if "lm_gemm" in which:
    # the decoder GEMMs that run furthest below the tensor roofline in situ: SwiGLU-backward epilogue (502 TFLOP/s)
    # and the N = 1024 products with the fp32-residual epilogue (818 TFLOP/s; 232 tiles = 3.14 waves on 74 CTA pairs)
    M, D, Fd = 14688, 1024, 3072
    dy = torch.randn(M, D, device=dev, dtype=BF16)
    wd_t = torch.randn(Fd, D, device=dev, dtype=BF16) * 0.03
    gu = torch.randn(M, 2 * Fd, device=dev, dtype=BF16)
    dgu = torch.empty(M, 2 * Fd, device=dev, dtype=BF16)
    h = torch.randn(M, Fd, device=dev, dtype=BF16)
    wd = torch.randn(D, Fd, device=dev, dtype=BF16) * 0.03
    resid = torch.zeros(M, D, device=dev, dtype=F32)
    y = torch.empty(M, D, device=dev, dtype=F32)
    for _ in range(3):
        L.gemm(dy, wd_t, epi=L.EPI_SWIGLU_BWD, aux=gu, out=dgu)
        L.gemm(h, wd, epi=L.EPI_F32_RESID, resid=resid, out=y)
if "window" in which:
    # QFormer window attention at batch 32 x 30 s: 3200 windows, 16 heads x 80, cross-attention (15 keys), both formulations
    from tiny_audio_b200.projectors import _WindowAttnFn
    Wn, heads, hd = 3200, 16, 80
    q = torch.randn(Wn, 3, heads * hd, device=dev, dtype=BF16, requires_grad=True)
    k = torch.randn(Wn, 15, heads * hd, device=dev, dtype=BF16, requires_grad=True)
    v = torch.randn(Wn, 15, heads * hd, device=dev, dtype=BF16, requires_grad=True)
    g = torch.randn(Wn, 3, heads * hd, device=dev, dtype=BF16)
    for variant in (1, 2):
        prev = lib.ta_window_attn_set_variant(variant)
        for _ in range(2):
            _WindowAttnFn.apply(q, k, v, None, heads).backward(g)
        lib.ta_window_attn_set_variant(prev)
if "lora_wgrad" in which:
    # rank-8 (padded to 128) LoRA gradient products: dBs [6144, 128] = dy^T t and dA [128, 1024] = u^T x over 14688 tokens,
    # unsplit and with the split-K form (ta_gemm_set_tn_splitk)
    M = 14688
    dy = torch.randn(M, 6144, device=dev, dtype=BF16)
    t = torch.randn(M, 128, device=dev, dtype=BF16)
    x = torch.randn(M, 1024, device=dev, dtype=BF16)
    for on in (0, 1):
        L.check(lib.ta_gemm_set_tn_splitk(on))
        for _ in range(3):
            L.gemm_tn(dy, t)
            L.gemm_tn(t, x)
    L.check(lib.ta_gemm_set_tn_splitk(0))
torch.cuda.synchronize()
print("done")
