"""Joules per launch of the step's kernels at production shapes: the step is power-capped (nvidia-smi: sw_power_cap, SM clock ~1.55 of
1.965 GHz), so its duration is (energy per step) / (board power limit) and a kernel's cost is its ENERGY, not its isolated run time.

Each kernel runs back to back for `--seconds` while a sampler thread reads board power (NVML instantaneous field when the driver
offers it, else the averaged reading) and the SM clock; the first 40 % of the window is discarded (power ramps, NVML averages).
Reported per kernel: time per launch (CUDA events over the whole window), mean power, mean SM clock, J per launch, and J per step =
J per launch x launches per step (the 32-layer encoder / 28-layer decoder counts of the headline workload).

  python tools/energy_profile.py [--seconds 1.5] > profiles/rNN_energy_profile.txt
"""
import argparse
import sys
import threading
import time

import torch

sys.path.insert(0, ".")
from tiny_audio_b200 import lib as L  # noqa: E402

BF16, F32 = torch.bfloat16, torch.float32
dev = "cuda"


class Sampler:
    def __init__(self):
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        self.h = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
        self.limit_w = pynvml.nvmlDeviceGetEnforcedPowerLimit(self.h) / 1000.0
        self.instant = False
        try:
            v = pynvml.nvmlDeviceGetFieldValues(self.h, [pynvml.NVML_FI_DEV_POWER_INSTANT])
            self.instant = v[0].nvmlReturn == 0
        except Exception:
            self.instant = False
        self.samples = []
        self.stop = False

    def read(self):
        nv = self.nv
        if self.instant:
            v = nv.nvmlDeviceGetFieldValues(self.h, [nv.NVML_FI_DEV_POWER_INSTANT])
            p = v[0].value.uiVal / 1000.0
        else:
            p = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
        c = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
        return p, c

    def run(self):
        while not self.stop:
            self.samples.append((time.perf_counter(),) + self.read())
            time.sleep(0.01)

    def measure(self, fn, seconds):
        """-> (ms per launch, mean W, mean MHz, n launches)"""
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        self.samples, self.stop = [], False
        th = threading.Thread(target=self.run)
        th.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        n = 0
        e0.record()
        while time.perf_counter() - t0 < seconds:
            for _ in range(20):
                fn()
            n += 20
            if n % 200 == 0:
                torch.cuda.synchronize()          # keep the launch queue bounded
        e1.record()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        self.stop = True
        th.join()
        cut = t0 + 0.4 * (t1 - t0)
        tail = [(p, c) for (t, p, c) in self.samples if cut <= t <= t1]
        pw = sum(p for p, _ in tail) / max(len(tail), 1)
        ck = sum(c for _, c in tail) / max(len(tail), 1)
        return e0.elapsed_time(e1) / n, pw, ck, n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=1.5)
    a = ap.parse_args()
    lib = L.load()
    sm = Sampler()
    print(f"# board power limit {sm.limit_w:.0f} W, power reading: {'instantaneous' if sm.instant else 'driver-averaged'}, "
          f"{a.seconds:.1f} s per kernel (first 40 % discarded)")
    time.sleep(1.0)
    p_idle, c_idle = sm.read()
    print(f"# idle: {p_idle:.0f} W at {c_idle} MHz")
    rows = []          # (name, fn, launches per step, flop per launch)

    def gemm_row(name, M, N, K, epi, per_step, **extra):
        x = torch.randn(M, K, device=dev, dtype=BF16)
        w = torch.randn(N, K, device=dev, dtype=BF16) * 0.03
        kw = {}
        if epi in (L.EPI_BF16_GELU, L.EPI_BF16_RESID, L.EPI_BF16_ROPE):
            kw["bias"] = torch.zeros(N, device=dev, dtype=F32)
        if epi == L.EPI_BF16_RESID:
            kw["resid"] = torch.zeros(M, N, device=dev, dtype=BF16)
        if epi == L.EPI_F32_RESID:
            kw["resid"] = torch.zeros(M, N, device=dev, dtype=F32)
        if epi == L.EPI_SWIGLU:
            kw["out2"] = torch.empty(M, N, device=dev, dtype=BF16)
        if epi == L.EPI_SWIGLU_BWD:
            kw["aux"] = torch.randn(M, 2 * N, device=dev, dtype=BF16)
        out = L.gemm(x, w, epi=epi, **kw)
        rows.append((name, lambda: L.gemm(x, w, epi=epi, out=out, **kw), per_step, 2.0 * M * N * K))

    Me, Md = 48000, 14848
    gemm_row("enc fc1 + GELU        48000x5120x1280", Me, 5120, 1280, L.EPI_BF16_GELU, 32)
    gemm_row("enc fc2 + residual    48000x1280x5120", Me, 1280, 5120, L.EPI_BF16_RESID, 32)
    gemm_row("enc qkv               48000x3840x1280", Me, 3840, 1280, L.EPI_BF16, 32)
    gemm_row("enc o + residual      48000x1280x1280", Me, 1280, 1280, L.EPI_BF16_RESID, 32)
    gemm_row("dec gate_up SwiGLU+stash 14848x6144x1024", Md, 6144, 1024, L.EPI_SWIGLU, 28)
    gemm_row("dec down + f32 resid  14848x1024x3072", Md, 1024, 3072, L.EPI_F32_RESID, 28)
    gemm_row("dec d(h) SwiGLU-bwd   14848x3072x1024", Md, 3072, 1024, L.EPI_SWIGLU_BWD, 28)
    gemm_row("dec d(xn) from gu     14848x1024x6144", Md, 1024, 6144, L.EPI_BF16, 28)
    gemm_row("dec qkv               14848x4096x1024", Md, 4096, 1024, L.EPI_BF16, 28)
    gemm_row("dec d(xn) from qkv    14848x1024x4096", Md, 1024, 4096, L.EPI_BF16, 28)
    gemm_row("dec o + f32 resid     14848x1024x2048", Md, 1024, 2048, L.EPI_F32_RESID, 28)
    gemm_row("dec d(att)            14848x2048x1024", Md, 2048, 1024, L.EPI_BF16, 28)

    # encoder attention
    B, S, H, hd = 32, 1500, 20, 64
    qkv = torch.randn(B, S, 3 * H * hd, device=dev, dtype=BF16)
    o = torch.empty(B, S, H * hd, device=dev, dtype=BF16)
    rows.append(("enc attention fwd     32x20x1500x1500x64",
                 lambda: L.check(lib.ta_attn_fwd(L.ptr(qkv), L.ptr(qkv[:, :, H * hd:]), L.ptr(qkv[:, :, 2 * H * hd:]), L.ptr(o), None, B, S, H, H, hd,
                                                 3 * H * hd, 3 * H * hd, 3 * H * hd, H * hd, 0, hd ** -0.5, L.stream_ptr())), 32,
                 4.0 * B * H * S * S * hd))
    # decoder attention
    Bd, Sd, Hq, Hkv, hdd = 32, 464, 16, 8, 128
    q = torch.randn(Bd, Sd, Hq * hdd, device=dev, dtype=BF16)
    k = torch.randn(Bd, Sd, Hkv * hdd, device=dev, dtype=BF16)
    v = torch.randn(Bd, Sd, Hkv * hdd, device=dev, dtype=BF16)
    do = torch.randn(Bd, Sd, Hq * hdd, device=dev, dtype=BF16)
    od = torch.empty_like(q)
    lse = torch.empty(Bd, Hq, Sd, device=dev, dtype=F32)
    dsum = torch.empty_like(lse)
    dq = torch.empty(Bd, Sd, Hq * hdd, device=dev, dtype=F32)
    dk, dv = torch.empty_like(k), torch.empty_like(v)
    fl_att = 4.0 * Bd * Hq * Sd * Sd * hdd / 2
    rows.append(("dec attention fwd     causal GQA 464x128",
                 lambda: L.check(lib.ta_attn_fwd(L.ptr(q), L.ptr(k), L.ptr(v), L.ptr(od), L.ptr(lse), Bd, Sd, Hq, Hkv, hdd, Hq * hdd, Hkv * hdd,
                                                 Hkv * hdd, Hq * hdd, 1, hdd ** -0.5, L.stream_ptr())), 28, fl_att))
    rows.append(("dec attention bwd     (prep + memset + main)",
                 lambda: L.check(lib.ta_attn_bwd(L.ptr(q), L.ptr(k), L.ptr(v), L.ptr(od), L.ptr(do), L.ptr(lse), L.ptr(dsum), L.ptr(dq), L.ptr(dk),
                                                 L.ptr(dv), Bd, Sd, Hq, Hkv, hdd, Hq * hdd, Hkv * hdd, Hkv * hdd, Hq * hdd, Hq * hdd, Hq * hdd,
                                                 Hkv * hdd, Hkv * hdd, 1, hdd ** -0.5, L.stream_ptr())), 28, 2.5 * fl_att))
    # the bandwidth-bound glue, represented by the largest one: encoder LayerNorm
    xe = torch.randn(Me, 1280, device=dev, dtype=BF16)
    ye = torch.empty_like(xe)
    wl, bl = torch.ones(1280, device=dev, dtype=F32), torch.zeros(1280, device=dev, dtype=F32)
    rows.append(("enc LayerNorm         48000x1280 bf16",
                 lambda: L.check(lib.ta_layernorm_bf16(L.ptr(xe), L.ptr(wl), L.ptr(bl), L.ptr(ye), Me, 1280, 1e-5, L.stream_ptr())), 65, 0.0))
    # library reference point
    a8 = torch.randn(8192, 8192, device=dev, dtype=BF16)
    b8 = torch.randn(8192, 8192, device=dev, dtype=BF16)
    rows.append(("cuBLAS bf16 8192^3 (reference point)", lambda: torch.matmul(a8, b8), 0, 2.0 * 8192 ** 3))

    print(f"{'kernel':44s} {'us':>8s} {'TFLOP/s':>8s} {'W':>6s} {'MHz':>6s} {'mJ/launch':>10s} {'pJ/FLOP':>8s} {'x/step':>6s} {'J/step':>7s} {'ms/step':>8s}")
    tot_j = tot_ms = 0.0
    for name, fn, per_step, flop in rows:
        ms, pw, ck, n = sm.measure(fn, a.seconds)
        mj = pw * ms
        tot_j += mj * per_step / 1e3
        tot_ms += ms * per_step
        print(f"{name:44s} {ms * 1e3:8.1f} {flop / ms / 1e9 if flop else 0.0:8.0f} {pw:6.0f} {ck:6.0f} {mj:10.2f} "
              f"{(mj * 1e9 / flop) if flop else 0.0:8.3f} {per_step:6d} {mj * per_step / 1e3:7.2f} {ms * per_step:8.2f}", flush=True)
        time.sleep(0.3)
    print(f"# listed kernels: {tot_j:.1f} J and {tot_ms:.1f} ms per step when each runs alone at the power cap "
          f"(step measured by bench.py: ~125 ms => ~{0.125 * sm.limit_w:.0f} J at the {sm.limit_w:.0f} W limit)")


if __name__ == "__main__":
    main()
