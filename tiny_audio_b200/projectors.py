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


class _TcLinearFn(torch.autograd.Function):
    """y = x W^T + b on the tcgen05 GEMM (bf16 operands, fp32 accumulate), with dgrad and wgrad on the same kernel:
    dx = dy W,  dW = dy^T x (fp32 out, operands transposed by ta_transpose_bf16),  db = column sums of dy."""

    @staticmethod
    def forward(ctx, x, w, b):
        from .engine import BF16
        L.require_cuda(x, w)
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1]).to(BF16).contiguous()
        wb = w.detach().to(BF16).contiguous()
        bias = b.detach().float().contiguous() if b is not None else None
        out = L.gemm(x2, wb, epi=L.EPI_BF16, bias=bias)
        ctx.save_for_backward(x2, wb)
        ctx.meta = (lead, w.dtype, b.dtype if b is not None else None, x.dtype)
        return out.view(*lead, w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        from .engine import BF16, F32
        lib = L.load()
        x2, wb = ctx.saved_tensors
        lead, wdt, bdt, xdt = ctx.meta
        N, K = wb.shape
        M = x2.shape[0]
        dy2 = dy.reshape(M, N).to(BF16).contiguous()
        st = L.stream_ptr()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            wt = torch.empty(K, N, device=wb.device, dtype=BF16)
            L.check(lib.ta_transpose_bf16(L.ptr(wb), L.ptr(wt), N, K, K, N, st))
            # the GEMM needs an output width that is a multiple of 128: all widths on this path are (1280, 5120, ...)
            dx = L.gemm(dy2, wt, epi=L.EPI_BF16).view(*lead, K).to(xdt)
        if ctx.needs_input_grad[1]:
            Mp = (M + 7) // 8 * 8
            dyt = torch.zeros(N, Mp, device=wb.device, dtype=BF16)
            xt = torch.zeros(K, Mp, device=wb.device, dtype=BF16)
            L.check(lib.ta_transpose_bf16(L.ptr(dy2), L.ptr(dyt), M, N, N, Mp, st))
            L.check(lib.ta_transpose_bf16(L.ptr(x2), L.ptr(xt), M, K, K, Mp, st))
            dw = L.gemm(dyt, xt, epi=L.EPI_F32, k=M).to(wdt)
        if bdt is not None and ctx.needs_input_grad[2]:
            db = dy2.float().sum(0).to(bdt)
        return dx, dw, db


def tc_linear(x, weight, bias=None):
    return _TcLinearFn.apply(x, weight, bias)


class QFormerAudioProjector(nn.Module):
    """BLIP-2 QFormer projector with learnable queries (reference: tiny_audio/projectors.py:359-475; arithmetic of
    HF:models/blip_2/modeling_blip_2.py:537-1042).  Parameter names and initialisation are the reference's (the HF
    `Blip2QFormerModel` is instantiated as the owner of the weights, exactly as the reference does), so checkpoints
    interchange.  Every linear -- q/k/v/o of self- and cross-attention, the FFN, the final projection; forward, dgrad and
    wgrad -- runs on the tcgen05 GEMM; LayerNorm, the 3x3 / 3x15 softmax attention and GELU are PyTorch glue (< 1 % of the
    projector's FLOPs; dedicated kernels are a 'next' item in DESIGN.md)."""

    def __init__(self, config):
        super().__init__()
        from transformers import AutoModel, Blip2QFormerConfig
        encoder_dim, llm_dim = config.encoder_dim, config.llm_dim
        self.window_size = getattr(config, "qformer_window_size", 15)
        self.downsample_rate = getattr(config, "downsample_rate", 5)
        self.num_queries = self.window_size // self.downsample_rate
        hidden = getattr(config, "qformer_hidden_size", None) or encoder_dim
        layers = getattr(config, "qformer_num_layers", 2)
        heads = getattr(config, "qformer_num_heads", 16)
        inter = getattr(config, "qformer_intermediate_size", None) or hidden * 4
        self.num_heads = heads
        self.query = nn.Parameter(torch.zeros(1, self.num_queries, hidden))
        self.query.data.normal_(mean=0.0, std=1.0)
        self.encoder_proj = nn.Linear(encoder_dim, hidden, bias=False) if encoder_dim != hidden else None
        qcfg = Blip2QFormerConfig(hidden_size=hidden, num_hidden_layers=layers, num_attention_heads=heads, intermediate_size=inter,
                                  encoder_hidden_size=hidden, cross_attention_frequency=1, hidden_act="gelu",
                                  attention_probs_dropout_prob=0.1, hidden_dropout_prob=0.1, layer_norm_eps=1e-12,
                                  initializer_range=0.02)
        self.qformer = AutoModel.from_config(qcfg)      # weight owner only; its forward is never called
        self.linear = nn.Linear(hidden, llm_dim)
        self.p_hidden, self.p_attn, self.ln_eps = 0.1, 0.1, 1e-12

    def get_output_length(self, input_length):
        nblocks = (input_length + self.window_size - 1) // self.window_size
        return nblocks * self.num_queries

    def _attend(self, att, x, kv_src):
        """Blip2QFormerMultiHeadAttention + SelfOutput: x [W, q, H] queries, kv_src [W, n, H] keys/values."""
        F_ = torch.nn.functional
        Wn, nq, H = x.shape
        hd = H // self.num_heads
        q = tc_linear(x, att.attention.query.weight, att.attention.query.bias)
        k = tc_linear(kv_src, att.attention.key.weight, att.attention.key.bias)
        v = tc_linear(kv_src, att.attention.value.weight, att.attention.value.bias)
        q = q.view(Wn, nq, self.num_heads, hd).transpose(1, 2).float()
        k = k.view(Wn, -1, self.num_heads, hd).transpose(1, 2).float()
        v = v.view(Wn, -1, self.num_heads, hd).transpose(1, 2).float()
        probs = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(hd), dim=-1)
        probs = F_.dropout(probs, self.p_attn, self.training)
        ctx = (probs @ v).transpose(1, 2).reshape(Wn, nq, H)
        o = tc_linear(ctx, att.output.dense.weight, att.output.dense.bias).float()
        o = F_.dropout(o, self.p_hidden, self.training)
        return F_.layer_norm(o + x.float(), (H,), att.output.LayerNorm.weight.float(), att.output.LayerNorm.bias.float(), self.ln_eps)

    def forward(self, hidden_states: torch.Tensor) -> torch.Tensor:
        if not hidden_states.is_cuda:
            raise L.TinyAudioB200Error("QFormerAudioProjector runs only on CUDA (tiny_audio_b200 has no CPU fallback)")
        F_ = torch.nn.functional
        B, S, _ = hidden_states.shape
        x_enc = hidden_states
        if self.encoder_proj is not None:
            x_enc = tc_linear(x_enc, self.encoder_proj.weight)
        nblocks = math.ceil(S / self.window_size)
        pad = nblocks * self.window_size - S
        if pad > 0:
            x_enc = F_.pad(x_enc, (0, 0, 0, pad))
        Wn = B * nblocks
        x_enc = x_enc.reshape(Wn, self.window_size, -1)
        qf = self.qformer
        H = self.query.shape[-1]
        x = F_.layer_norm(self.query.float(), (H,), qf.layernorm.weight.float(), qf.layernorm.bias.float(), self.ln_eps)
        x = F_.dropout(x, self.p_hidden, self.training).expand(Wn, -1, -1)
        for layer in qf.encoder.layer:
            x = self._attend(layer.attention, x, x)
            x = self._attend(layer.crossattention, x, x_enc)
            h = tc_linear(x, layer.intermediate_query.dense.weight, layer.intermediate_query.dense.bias)
            h = F_.gelu(h)
            f = tc_linear(h, layer.output_query.dense.weight, layer.output_query.dense.bias).float()
            f = F_.dropout(f, self.p_hidden, self.training)
            x = F_.layer_norm(f + x, (H,), layer.output_query.LayerNorm.weight.float(), layer.output_query.LayerNorm.bias.float(),
                              self.ln_eps)
        out = tc_linear(x.reshape(B, nblocks * self.num_queries, H), self.linear.weight, self.linear.bias)
        return out


class _NotOnThePath(nn.Module):
    """mosa / moe are registered names in the reference (projectors.py:482-487) but no BASELINE config uses them
    (out of scope); they fail loudly instead of silently running a different implementation."""

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
    "qformer": QFormerAudioProjector,
}
