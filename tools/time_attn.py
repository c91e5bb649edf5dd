"""Time the encoder-shape attention forward variants (ta_attn_set_tc modes) at the production shape."""
import sys
import torch
sys.path.insert(0, ".")
from tiny_audio_b200 import lib as L

lib = L.load()
BF16 = torch.bfloat16
B, S, H, hd = 32, 1500, 20, 64
torch.manual_seed(0)
qkv = torch.randn(B, S, 3 * H * hd, device="cuda", dtype=BF16)
o = torch.empty(B, S, H * hd, device="cuda", dtype=BF16)
modes = [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4]


def run():
    L.check(lib.ta_attn_fwd(L.ptr(qkv), L.ptr(qkv[:, :, H * hd:]), L.ptr(qkv[:, :, 2 * H * hd:]), L.ptr(o), None, B, S, H, H, hd,
                            3 * H * hd, 3 * H * hd, 3 * H * hd, H * hd, 0, hd ** -0.5, L.stream_ptr()))


for tc in modes:
    lib.ta_attn_set_tc(tc)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 20
    print(f"enc attention fwd mode={tc}: {t:.3f} ms  {4.0 * B * H * S * S * hd / t / 1e9:.0f} TF/s  "
          f"{B * H * S * S / t / 1e6 / 148:.2f} Gexp/s/SM", flush=True)
lib.ta_attn_set_tc(14)
