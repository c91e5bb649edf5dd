"""Timeline of the persistent encoder-attention kernel (ta_attn_set_tc modes 6+): SM-clock stamps of CTA 0's pipeline events
(ta_attn_set_trace) -> per-phase durations in cycles, steady-state medians.

  slots 0 / 1 = softmax warp 0 of group A / B, events per kv-tile step:
     0 before the s_full wait | 1 S product arrived | 2 S in registers | 3 row max done | 4 pv_done (P buffer free) | 5 exp2 phase starts
     6 P written (before p_full arrive)
  slot 2 = group A's MMA thread, per step: 0 loop top | 1 S_A(t+1) issued | 2 PV_A(t) issued
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


for mode in [int(a) for a in sys.argv[1:]] or [8]:
    lib.ta_attn_set_tc(mode)
    for _ in range(2):
        run()
    buf = torch.zeros(3, STEPS, 8, device="cuda", dtype=torch.int64)
    L.check(lib.ta_attn_set_trace(L.ptr(buf), STEPS))
    run()
    torch.cuda.synchronize()
    L.check(lib.ta_attn_set_trace(None, 0))
    t = buf.cpu().numpy().astype(np.float64)
    lo, hi = 24, STEPS - 24                                     # steady state
    a, b, m = t[0, lo:hi], t[1, lo:hi], t[2, lo:hi]
    print(f"== mode {mode}: step period (softmax A) median {np.median(np.diff(t[0, lo:hi, 0])):.0f} clk, total {t[0, -1, 6] - t[0, 0, 0]:.0f} clk for {STEPS} steps")
    names = ["wait S product", "TMEM load of S", "mask + row max", "wait PV (P free)", "wait turn", "exp2 + P write", "arrive -> next step top"]
    for g, x, full in (("A", a, t[0]), ("B", b, t[1])):
        d = [np.median(x[:, i + 1] - x[:, i]) for i in range(6)]
        d.append(np.median(full[lo + 1:hi + 1, 0] - full[lo:hi, 6]))
        print(f"  softmax {g}: " + " | ".join(f"{n} {v:.0f}" for n, v in zip(names, d)))
    mn = ["issue S_A(t+1) (waits K, s_empty A)", "issue PV_A(t) (waits p_full A, V)", "loop"]
    d = [np.median(m[:, i + 1] - m[:, i]) for i in range(2)] + [np.median(t[2, lo + 1:hi + 1, 0] - t[2, lo:hi, 2])]
    print("  MMA thread: " + " | ".join(f"{n} {v:.0f}" for n, v in zip(mn, d)))
    # alignment of the chains: when does the S product for step t arrive relative to when softmax A asks for it?
    # MMA issue time of S_A(t+1) is event 1 of MMA step t
    ask = t[0, lo + 1:hi + 1, 0]
    got = t[0, lo + 1:hi + 1, 1]
    issued = t[2, lo:hi, 1]
    print(f"  S_A(t+1): issued {np.median(ask - issued):.0f} clk BEFORE softmax A asks for it (negative = late); arrives {np.median(got - issued):.0f} clk after issue")
    pfull = t[0, lo:hi, 6]
    pv_issued = t[2, lo:hi, 2]
    print(f"  PV_A(t): issued {np.median(pv_issued - pfull):.0f} clk after softmax A finished writing P; P free again {np.median(t[0, lo + 1:hi + 1, 4] - pv_issued):.0f} clk after issue")
lib.ta_attn_set_tc(2)
