"""LayerNorm / RMSNorm kernels at production shapes over warps-per-block (ta_debug_set key 3) and row order: GB/s of algorithmic traffic."""
import sys
import torch
sys.path.insert(0, ".")
from tiny_audio_b200 import lib as L

lib = L.load()
BF16, F32 = torch.bfloat16, torch.float32
dev = "cuda"


def timeit(fn, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


rows, D = 48000, 1280
x = torch.randn(rows, D, device=dev, dtype=BF16)
w = torch.ones(D, device=dev, dtype=F32)
y = torch.empty_like(x)
big = torch.empty(256 << 20, device=dev, dtype=torch.uint8)       # L2 flush between launches would hide nothing here: tensors are 123 MB
rl, Dl = 14848, 1024
xf = torch.randn(rl, Dl, device=dev, dtype=F32)
w1 = torch.ones(Dl, device=dev, dtype=F32)
yb = torch.empty(rl, Dl, device=dev, dtype=BF16)
dx = torch.zeros(rl, Dl, device=dev, dtype=F32)
dxb = torch.empty(rl, Dl, device=dev, dtype=BF16)
for wpb in (8, 4, 2, 1):
    lib.ta_debug_set(3, wpb)
    for rev in (1, 0):
        lib.ta_layernorm_set_reverse(rev)
        t = timeit(lambda: L.check(lib.ta_layernorm_bf16(L.ptr(x), L.ptr(w), L.ptr(w), L.ptr(y), rows, D, 1e-5, L.stream_ptr())))
        print(f"wpb {wpb} reverse {rev}: LayerNorm 48000x1280   {t:6.1f} us  {2 * rows * D * 2 / t / 1e3:6.0f} GB/s")
    t = timeit(lambda: L.check(lib.ta_rmsnorm_f32(L.ptr(xf), L.ptr(w1), L.ptr(yb), None, rl, Dl, 1e-6, L.stream_ptr())))
    print(f"wpb {wpb}: RMSNorm fwd 14848x1024 f32 -> bf16   {t:6.1f} us  {rl * Dl * 6 / t / 1e3:6.0f} GB/s")
    t = timeit(lambda: L.check(lib.ta_rmsnorm_f32_bwd(L.ptr(yb), L.ptr(xf), L.ptr(w1), L.ptr(dx), None, rl, Dl, 1e-6, 1, L.stream_ptr())))
    print(f"wpb {wpb}: RMSNorm bwd (accumulate)              {t:6.1f} us  {rl * Dl * 14 / t / 1e3:6.0f} GB/s")
lib.ta_debug_set(3, 8)
lib.ta_layernorm_set_reverse(1)
