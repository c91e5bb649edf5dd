"""Decoder (Qwen3, head_dim 128, causal, GQA 16/8) attention forward + backward at the production shape, with the backward's dQ
atomics switched off as an A/B (ta_debug_set key 1 bit 0): how much of the backward is the fp32 red.global.add traffic?"""
import sys
import torch
sys.path.insert(0, ".")
from tiny_audio_b200 import lib as L

lib = L.load()
BF16, F32 = torch.bfloat16, torch.float32
B, S, Hq, Hkv, hd = 32, 464, 16, 8, 128
dev = "cuda"
torch.manual_seed(0)
q = torch.randn(B, S, Hq * hd, device=dev, dtype=BF16)
k = torch.randn(B, S, Hkv * hd, device=dev, dtype=BF16)
v = torch.randn(B, S, Hkv * hd, device=dev, dtype=BF16)
do = torch.randn(B, S, Hq * hd, device=dev, dtype=BF16)
o = torch.empty_like(q)
lse = torch.empty(B, Hq, S, device=dev, dtype=F32)
dsum = torch.empty_like(lse)
dq = torch.empty(B, S, Hq * hd, device=dev, dtype=F32)
dk, dv = torch.empty_like(k), torch.empty_like(v)


def fwd():
    L.check(lib.ta_attn_fwd(L.ptr(q), L.ptr(k), L.ptr(v), L.ptr(o), L.ptr(lse), B, S, Hq, Hkv, hd, Hq * hd, Hkv * hd, Hkv * hd, Hq * hd, 1,
                            hd ** -0.5, L.stream_ptr()))


def bwd():
    L.check(lib.ta_attn_bwd(L.ptr(q), L.ptr(k), L.ptr(v), L.ptr(o), L.ptr(do), L.ptr(lse), L.ptr(dsum), L.ptr(dq), L.ptr(dk), L.ptr(dv), B, S,
                            Hq, Hkv, hd, Hq * hd, Hkv * hd, Hkv * hd, Hq * hd, Hq * hd, Hq * hd, Hkv * hd, Hkv * hd, 1, hd ** -0.5, L.stream_ptr()))


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


fl = 2.0 * B * Hq * S * S * hd          # causal: half of 4 S^2 hd
for variant, name in ((0, "128-key tiles, 1 CTA/SM"), (3, "64-key tiles, 2 CTAs/SM, 3-slot ring"), (4, "64-key tiles, 4-slot ring"), (6, "64-key tiles, 1 CTA/SM, 6-slot ring"), (8, "64-key tiles, 1 CTA/SM, 8-slot ring"), (1, "default")):
    lib.ta_attn_set_tc_lm(variant)
    t = timeit(fwd)
    print(f"forward [{name:38s}] {t:7.1f} us  ({fl / t / 1e6:.0f} TFLOP/s)  ring slots {lib.ta_attn_tc_lm_ring_slots()}")
for variant, name in ((1, "128-query serial kernel"), (2, "64-query pipelined kernel (default)")):
    lib.ta_attn_set_bwd_variant(variant)
    t = timeit(bwd)
    print(f"backward [{name:36s}] (prep + memset + main) {t:7.1f} us  ({2.5 * fl / t / 1e6:.0f} TFLOP/s)")
    lib.ta_debug_set(1, 1)
    print(f"backward [{name:36s}] dQ atomics OFF         {timeit(bwd):7.1f} us")
    lib.ta_debug_set(1, 0)
lib.ta_attn_set_bwd_variant(2)
