"""Timeline of the persistent encoder-attention kernel (ta_attn_set_tc modes 6+): SM-clock stamps of CTA 0's pipeline events
(ta_attn_set_trace) -> per-phase durations in cycles, steady-state medians.

  slots 0-7 = softmax warps (group A = 0-3, group B = 4-7; warp q and q+4 share SM sub-partition q), events per kv-tile step:
     0 before the s_full wait | 1 S product arrived | 2 S in registers | 3 row max done | 4 pv_done (P buffer free) | 5 exp2 phase starts
     6 P written (before the p_full arrive)
  slots 8 / 9 = MMA threads of group A / B: 0 loop top | 1 S(t+1) issued | 3 p_full seen | 4 V tile + O free seen | 2 PV(t) issued
usage: python tools/attn_trace.py [mode ...]"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from tiny_audio_b200 import lib as L

lib = L.load()
B, S, H, hd = 32, 1500, 20, 64
qkv = torch.randn(B, S, 3 * H * hd, device="cuda", dtype=torch.bfloat16)
o = torch.empty(B, S, H * hd, device="cuda", dtype=torch.bfloat16)
STEPS = 26 * 12


def run():
    L.check(lib.ta_attn_fwd(L.ptr(qkv), L.ptr(qkv[:, :, H * hd:]), L.ptr(qkv[:, :, 2 * H * hd:]), L.ptr(o), None, B, S, H, H, hd,
                            3 * H * hd, 3 * H * hd, 3 * H * hd, H * hd, 0, hd ** -0.5, L.stream_ptr()))


med = np.median
for mode in [int(a) for a in sys.argv[1:]] or [8]:
    lib.ta_attn_set_tc(mode)
    for _ in range(2):
        run()
    buf = torch.zeros(10, STEPS, 8, device="cuda", dtype=torch.int64)
    L.check(lib.ta_attn_set_trace(L.ptr(buf), STEPS))
    run()
    torch.cuda.synchronize()
    L.check(lib.ta_attn_set_trace(None, 0))
    t = buf.cpu().numpy().astype(np.float64)
    lo, hi = 24, STEPS - 24                                     # steady state
    print(f"== mode {mode}: step period median {med(np.diff(t[0, lo:hi, 0])):.0f} clk, total {t[0, -1, 6] - t[0, 0, 0]:.0f} clk for {STEPS} steps")
    names = ["wait S", "TMEM ld", "max", "wait PV", "wait turn", "exp2+P", "to next top"]
    for w in range(8):
        x = t[w, lo:hi]
        d = [med(x[:, i + 1] - x[:, i]) for i in range(6)] + [med(t[w, lo + 1:hi + 1, 0] - t[w, lo:hi, 6])]
        print(f"  warp {w} ({'AB'[w // 4]}{w % 4}): " + " | ".join(f"{n} {v:.0f}" for n, v in zip(names, d)))
    for g in range(2):
        ws = slice(4 * g, 4 * g + 4)
        ld_done = t[ws, lo:hi, 2]          # S in registers
        p_done = t[ws, lo:hi, 6]
        m = t[8 + g, lo:hi]
        print(f"  group {'AB'[g]}: skew between its 4 warps at 'S in registers' {med(ld_done.max(0) - ld_done.min(0)):.0f} clk, at 'P written' "
              f"{med(p_done.max(0) - p_done.min(0)):.0f} clk")
        print(f"     MMA thread: S(t+1) issued {med(m[:, 1] - ld_done.max(0)):.0f} clk after the LAST warp had S(t) in registers; p_full seen "
              f"{med(m[:, 3] - p_done.max(0)):.0f} clk after the LAST warp wrote P; V/O-free wait {med(m[:, 4] - m[:, 3]):.0f}; PV issue {med(m[:, 2] - m[:, 4]):.0f}")
        nxt_pv = t[ws, lo + 1:hi + 1, 4]   # pv_done seen by the softmax warps in the next step
        print(f"     pv_done seen by the softmax warps {med(nxt_pv.min(0) - m[:, 2]):.0f} .. {med(nxt_pv.max(0) - m[:, 2]):.0f} clk after the PV issue")
lib.ta_attn_set_tc(14)
