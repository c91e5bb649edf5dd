"""Qwen3 FFN GEMMs at the production shape (M = 32 x 464 tokens): epilogue variants A/B (what does the SwiGLU epilogue / the stash cost?)."""
import sys
import torch
sys.path.insert(0, ".")
from tiny_audio_b200 import lib as L

BF16, F32 = torch.bfloat16, torch.float32
lib = L.load()
dev = "cuda"


def timeit(fn, reps=20, warm=3):
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


M, D, Fd = 14848, 1024, 3072
x = torch.randn(M, D, device=dev, dtype=BF16)
wgu = torch.randn(2 * Fd, D, device=dev, dtype=BF16) * 0.03
h = torch.empty(M, Fd, device=dev, dtype=BF16)
gu = torch.empty(M, 2 * Fd, device=dev, dtype=BF16)
plain = torch.empty(M, 2 * Fd, device=dev, dtype=BF16)
wd = torch.randn(D, Fd, device=dev, dtype=BF16) * 0.03
wd_t = wd.t().contiguous()
resid = torch.zeros(M, D, device=dev, dtype=F32)
y = torch.empty(M, D, device=dev, dtype=F32)
dy = torch.randn(M, D, device=dev, dtype=BF16)
dgu = torch.empty(M, 2 * Fd, device=dev, dtype=BF16)
fl_gu = 2.0 * M * 2 * Fd * D
fl_d = 2.0 * M * D * Fd
rows = [("gate_up plain bf16 out [M,6144]", lambda: L.gemm(x, wgu, epi=L.EPI_BF16, out=plain), fl_gu),
        ("gate_up SwiGLU, no stash", lambda: L.gemm(x, wgu, epi=L.EPI_SWIGLU, out=h), fl_gu),
        ("gate_up SwiGLU + (g,u) stash (training)", lambda: L.gemm(x, wgu, epi=L.EPI_SWIGLU, out=h, out2=gu), fl_gu),
        ("down + fp32 residual", lambda: L.gemm(h, wd, epi=L.EPI_F32_RESID, resid=resid, out=y), fl_d),
        ("d(h) SwiGLU-backward -> (d gate, d up)", lambda: L.gemm(dy, wd_t, epi=L.EPI_SWIGLU_BWD, aux=gu, out=dgu), fl_d),
        ("cuBLAS gate_up", lambda: torch.matmul(x, wgu.t()), fl_gu), ("cuBLAS down", lambda: torch.matmul(h, wd.t()), fl_d)]
def report(rows):
    for name, fn, fl in rows:
        t = timeit(fn)
        print(f"{name:42s} {t * 1e3:7.1f} us  {fl / t / 1e9:6.0f} TFLOP/s", flush=True)


report(rows)
# A/B switches
lib.ta_gemm_set_swiglu_bwd_tma(0)
report([("d(h) SwiGLU-backward, per-thread stash loads", rows[4][1], fl_d)])
lib.ta_gemm_set_swiglu_bwd_tma(1)
lib.ta_gemm_set_tail_split(1)
report([("down + fp32 residual, tail split 256x128", rows[3][1], fl_d)])
lib.ta_gemm_set_tail_split(0)
lib.ta_gemm_set_tile_n(128)
report([("down + fp32 residual, 256x128 tiles", rows[3][1], fl_d)])
lib.ta_gemm_set_tile_n(0)

# encoder bf16-residual GEMMs: per-thread residual loads vs the in-place TMA epilogue
Me = 48000
for name, N_, K_ in (("enc o + residual 48000x1280x1280", 1280, 1280), ("enc fc2 + residual 48000x1280x5120", 1280, 5120)):
    xe = torch.randn(Me, K_, device=dev, dtype=BF16)
    we = torch.randn(N_, K_, device=dev, dtype=BF16) * 0.03
    be = torch.zeros(N_, device=dev, dtype=F32)
    re_ = torch.zeros(Me, N_, device=dev, dtype=BF16)
    for mode in (0, 1):
        lib.ta_gemm_set_resid_tma(mode)
        report([(f"{name} [resid_tma={mode}]", lambda: L.gemm(xe, we, epi=L.EPI_BF16_RESID, bias=be, resid=re_, out=re_), 2.0 * Me * N_ * K_)])
lib.ta_gemm_set_resid_tma(0)
