"""Projector plugins (same contract as the reference's `tiny_audio/projectors.py`):

    PROJECTOR_CLASSES[name](config)  ->  nn.Module with .forward(x[B,T,D]) -> [B,T',llm_dim],
                                         .get_output_length(int | Tensor), optional .get_aux_loss()

`MLPAudioProjector` keeps the reference's parameter names (`linear_1.weight`, `norm.weight`,
`linear_2.weight`, `norm_2.weight`, all bias-free; projectors.py:23-71) so checkpoints interchange, but
its arithmetic runs in libtinyaudio_b200 (frame-stack as a view, tcgen05 GEMMs, fused RMSNorm(+GELU)
kernels, hand-written backward).  There is no PyTorch fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.nn as nn

from . import lib as L


def frame_stack_length(seq_len, k: int):
    """`(L - k) // k + 1` -- GLM-ASR merge formula (projectors.py:52-55)."""
    return (seq_len - k) // k + 1


class _Gain(nn.Module):
    """Holder for an RMSNorm gain so that parameter names match LlamaRMSNorm (`<name>.weight`)."""

    def __init__(self, dim: int, eps: float):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.variance_epsilon = eps


class _BiasFreeLinear(nn.Module):
    """Parameter holder with nn.Linear's default init (kaiming-uniform, a=sqrt(5)); no bias."""

    def __init__(self, in_features: int, out_features: int):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))


class _MLPProjectorFn(torch.autograd.Function):
    """frame-stack -> linear_1 -> RMSNorm -> GELU -> linear_2 -> RMSNorm, forward and backward in CUDA."""

    @staticmethod
    def forward(ctx, x, w1, n1, w2, n2, k, eps):
        from .engine import BF16, F32
        lib = L.load()
        L.require_cuda(x, w1, n1, w2, n2)
        B, S, D = x.shape
        n = frame_stack_length(S, k)
        H, O = w1.shape[0], w2.shape[0]
        dev = x.device
        st = L.stream_ptr()
        xb = x.detach().to(BF16).contiguous()
        if n * k == S:
            xs = xb.view(B * n, k * D)
        else:
            xs = torch.empty(B * n, k * D, device=dev, dtype=BF16)
            L.check(lib.ta_frame_stack(L.ptr(xb), L.ptr(xs), B, S, n, k, D, st))
        w1b = torch.empty(H, k * D, device=dev, dtype=BF16)
        w2b = torch.empty(O, H, device=dev, dtype=BF16)
        w2t = torch.empty(H, O, device=dev, dtype=BF16)
        w1f, w2f = w1.detach().float().contiguous(), w2.detach().float().contiguous()
        L.check(lib.ta_cast_f32_bf16(L.ptr(w1f), L.ptr(w1b), w1f.numel(), st))
        L.check(lib.ta_cast_f32_bf16(L.ptr(w2f), L.ptr(w2b), w2f.numel(), st))
        L.check(lib.ta_transpose_bf16(L.ptr(w2b), L.ptr(w2t), O, H, H, O, st))
        n1f, n2f = n1.detach().float().contiguous(), n2.detach().float().contiguous()
        pw = L.MlpProjectorWeights(k * D, H, O, eps, L.ptr(w1b), L.ptr(n1f), L.ptr(w2b), L.ptr(w2t), L.ptr(n2f))
        M = B * n
        y1 = torch.empty(M, H, device=dev, dtype=BF16)
        a1 = torch.empty(M, H, device=dev, dtype=BF16)
        y2 = torch.empty(M, O, device=dev, dtype=BF16)
        out = torch.empty(M, O, device=dev, dtype=F32)
        L.check(lib.ta_mlp_projector_forward(C.byref(pw), L.ptr(xs), M, L.ptr(y1), L.ptr(a1), L.ptr(y2), L.ptr(out), st))
        ctx.keep = (xs, w1b, w2b, w2t, n1f, n2f, y1, a1, y2)
        ctx.dims = (M, k * D, H, O, eps)
        ctx.param_dtypes = (w1.dtype, n1.dtype, w2.dtype, n2.dtype)
        return out.view(B, n, O)

    @staticmethod
    def backward(ctx, d_out):
        from .engine import F32
        lib = L.load()
        xs, w1b, w2b, w2t, n1f, n2f, y1, a1, y2 = ctx.keep
        M, I, H, O, eps = ctx.dims
        dev = xs.device
        st = L.stream_ptr()
        pw = L.MlpProjectorWeights(I, H, O, eps, L.ptr(w1b), L.ptr(n1f), L.ptr(w2b), L.ptr(w2t), L.ptr(n2f))
        nbytes = C.c_longlong()
        L.check(lib.ta_mlp_projector_backward_workspace_bytes(C.byref(pw), M, C.byref(nbytes)))
        ws = torch.empty(nbytes.value, device=dev, dtype=torch.uint8)
        g = d_out.detach().to(F32).contiguous().view(M, O)
        dw1 = torch.empty(H, I, device=dev, dtype=F32)
        dw2 = torch.empty(O, H, device=dev, dtype=F32)
        dn1 = torch.zeros(H, device=dev, dtype=F32)
        dn2 = torch.zeros(O, device=dev, dtype=F32)
        L.check(lib.ta_mlp_projector_backward(C.byref(pw), L.ptr(xs), M, L.ptr(y1), L.ptr(a1), L.ptr(y2), L.ptr(g), L.ptr(ws),
                                              nbytes.value, L.ptr(dw1), L.ptr(dn1), L.ptr(dw2), L.ptr(dn2), st))
        t1, tn1, t2, tn2 = ctx.param_dtypes
        # the encoder output is produced under no_grad in the path (asr_modeling.py:448-450): no d(x)
        return None, dw1.to(t1), dn1.to(tn1), dw2.to(t2), dn2.to(tn2), None, None


class MLPAudioProjector(nn.Module):
    """2-layer MLP with frame-stacking downsampling (reference: projectors.py:23-71)."""

    def __init__(self, config):
        super().__init__()
        encoder_dim = getattr(config, "encoder_dim", 768)
        llm_dim = getattr(config, "llm_dim", 2048)
        self.k = getattr(config, "projector_pool_stride", 4)
        hidden_dim = getattr(config, "projector_hidden_dim", None) or llm_dim
        self.linear_1 = _BiasFreeLinear(encoder_dim * self.k, hidden_dim)
        self.norm = _Gain(hidden_dim, 1e-6)
        self.linear_2 = _BiasFreeLinear(hidden_dim, llm_dim)
        self.norm_2 = _Gain(llm_dim, 1e-6)

    def get_output_length(self, input_length):
        return frame_stack_length(input_length, self.k)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise L.TinyAudioB200Error("MLPAudioProjector runs only on CUDA (tiny_audio_b200 has no CPU fallback)")
        return _MLPProjectorFn.apply(x, self.linear_1.weight, self.norm.weight, self.linear_2.weight, self.norm_2.weight,
                                     self.k, self.norm.variance_epsilon)


class _NotOnThePath(nn.Module):
    """mosa / moe / qformer are registered names in the reference (projectors.py:482-487).  They are outside
    this round's hot-path scope (configs 1-3, 5 use `mlp`; `qformer` is a 'next' row) and
    fail loudly instead of silently running a different implementation."""

    kind = "?"

    def __init__(self, config):
        super().__init__()
        raise NotImplementedError(
            f"projector_type={self.kind!r} is not implemented in tiny_audio_b200 yet (hot-path scope: 'mlp'); "
            "see DESIGN.md 'out of scope / next'.")


def _stub(kind):
    return type(f"{kind.upper()}ProjectorUnavailable", (_NotOnThePath,), {"kind": kind})


PROJECTOR_CLASSES = {
    "mlp": MLPAudioProjector,
    "mosa": _stub("mosa"),
    "moe": _stub("moe"),
    "qformer": _stub("qformer"),
}
