"""Projector plugins (same contract as the reference's `tiny_audio/projectors.py`):

    PROJECTOR_CLASSES[name](config)  ->  nn.Module with .forward(x[B,T,D]) -> [B,T',llm_dim],
                                         .get_output_length(int | Tensor), optional .get_aux_loss()

`MLPAudioProjector` keeps the reference's parameter names (`linear_1.weight`, `norm.weight`,
`linear_2.weight`, `norm_2.weight`, all bias-free; projectors.py:23-71) so checkpoints interchange, but
its arithmetic runs in libtinyaudio_b200 (frame-stack as a view, tcgen05 GEMMs, fused RMSNorm(+GELU)
kernels, hand-written backward).  `QFormerAudioProjector`, `MOSAProjector` and `MoEAudioProjector` (the other three
registered names, projectors.py:482-487) run every wide linear -- forward, dgrad and wgrad -- on the same tcgen05 GEMM
through `tc_linear`; the mixture projectors fold all their adapters into two GEMMs (`folded_adapter_mixture`).
There is no PyTorch fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
import math
import os

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


def _bf16_weight(w: torch.Tensor) -> torch.Tensor:
    """bf16 GEMM operand of a (usually fp32 master) weight: the library's cast kernel, no ATen copy."""
    w = w.detach()
    if w.dtype == torch.bfloat16:
        return w.contiguous()
    if w.dtype == torch.float32 and w.is_contiguous() and w.is_cuda:
        out = torch.empty(w.shape, device=w.device, dtype=torch.bfloat16)
        L.check(L.load().ta_cast_f32_bf16(L.ptr(w), L.ptr(out), w.numel(), L.stream_ptr()))
        return out
    return w.to(torch.bfloat16).contiguous()


class _TcLinearFn(torch.autograd.Function):
    """y = x W^T + b on the tcgen05 GEMM (bf16 operands, fp32 accumulate), with dgrad and wgrad on the same kernel:
    dx = dy W,  dW = dy^T x (fp32 out; the TN form of the GEMM reads both row-major activations as MN-major operands, no
    transposed copies),  db = column sums of dy."""

    @staticmethod
    def forward(ctx, x, w, b):
        from .engine import BF16
        L.require_cuda(x, w)
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1]).to(BF16).contiguous()
        wb = _bf16_weight(w)
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
            # fp32 activations in -> fp32 gradient out: the mixture projectors' gate gradients are sums over E * hidden products of
            # this tensor with strong cancellation, a bf16 round trip here costs them several % (measured 6.8e-2 -> see tests)
            dx = L.gemm(dy2, wt, epi=L.EPI_F32 if xdt == torch.float32 else L.EPI_BF16).view(*lead, K).to(xdt)
        if ctx.needs_input_grad[1]:
            if K % 128 == 0:
                # weight-gradient form of the GEMM: dW = dy^T x straight from the row-major activations (both operands MN-major)
                dw = L.gemm_tn(dy2, x2).to(wdt)
            else:
                Mp = (M + 7) // 8 * 8
                dyt = torch.zeros(N, Mp, device=wb.device, dtype=BF16)
                xt = torch.zeros(K, Mp, device=wb.device, dtype=BF16)
                L.check(lib.ta_transpose_bf16(L.ptr(dy2), L.ptr(dyt), M, N, N, Mp, st))
                L.check(lib.ta_transpose_bf16(L.ptr(x2), L.ptr(xt), M, K, K, Mp, st))
                dw = L.gemm(dyt, xt, epi=L.EPI_F32, k=M).to(wdt)
        if bdt is not None and ctx.needs_input_grad[2]:
            db32 = torch.empty(N, device=dy2.device, dtype=F32)
            L.check(lib.ta_colsum_bf16(L.ptr(dy2), N, L.ptr(db32), M, N, st))
            db = db32.to(bdt)
        return dx, dw, db


def tc_linear(x, weight, bias=None):
    return _TcLinearFn.apply(x, weight, bias)


class _WindowAttnFn(torch.autograd.Function):
    """dropout(softmax(q k^T / sqrt(hd))) v per (window, head) for the QFormer's tiny attention problems (3 queries x 3 or 15
    keys x 80): one warp per (window, head) in csrc/window_attn.cu, bf16 in / out, probabilities recomputed in the backward."""

    @staticmethod
    def forward(ctx, q, k, v, drop_mask, heads):
        from .engine import BF16
        lib = L.load()
        L.require_cuda(q, k, v)
        Wn, nq, H = q.shape
        nk = k.shape[1]
        hd = H // heads
        qb, kb, vb = (t.detach().to(BF16).contiguous() for t in (q, k, v))
        out = torch.empty_like(qb)
        scale = 1.0 / math.sqrt(hd)
        L.check(lib.ta_window_attn_fwd(L.ptr(qb), L.ptr(kb), L.ptr(vb), L.ptr(drop_mask), L.ptr(out), Wn, nq, nk, heads, hd, scale,
                                       L.stream_ptr()))
        ctx.save_for_backward(qb, kb, vb, drop_mask)
        ctx.meta = (heads, scale, q.dtype, k.dtype, v.dtype)
        return out

    @staticmethod
    def backward(ctx, g):
        from .engine import BF16
        lib = L.load()
        qb, kb, vb, drop_mask = ctx.saved_tensors
        heads, scale, qdt, kdt, vdt = ctx.meta
        Wn, nq, H = qb.shape
        nk = kb.shape[1]
        gb = g.detach().to(BF16).contiguous()
        dq, dk, dv = torch.empty_like(qb), torch.empty_like(kb), torch.empty_like(vb)
        L.check(lib.ta_window_attn_bwd(L.ptr(qb), L.ptr(kb), L.ptr(vb), L.ptr(drop_mask), L.ptr(gb), L.ptr(dq), L.ptr(dk), L.ptr(dv),
                                       Wn, nq, nk, heads, H // heads, scale, L.stream_ptr()))
        return dq.to(qdt), dk.to(kdt), dv.to(vdt), None, None


def window_attention(q, k, v, heads: int, dropout_p: float = 0.0, training: bool = False):
    """q [W, nq, H], k / v [W, nk, H] -> [W, nq, H] bf16.  Dropout on the probabilities uses torch's RNG (F.dropout on a tensor of
    ones gives the 0 or 1/(1-p) multipliers), so seeding behaves like the reference's nn.Dropout."""
    mask = None
    if training and dropout_p > 0.0:
        mask = torch.nn.functional.dropout(torch.ones(q.shape[0], heads, q.shape[1], k.shape[1], device=q.device,
                                                      dtype=torch.float32), dropout_p, True).contiguous()
    return _WindowAttnFn.apply(q, k, v, mask, heads)


class _AddLayerNormFn(torch.autograd.Function):
    """(y fp32, y bf16) = LayerNorm(bf16(o) * mask + resid[row % resid_rows]) * post_mask   in one kernel, and its backward in one
    kernel (csrc/qformer_glue.cu): Blip2QFormerSelfOutput / Blip2QFormerOutput (HF:models/blip_2/modeling_blip_2.py:637-648,693-704)
    and the query LayerNorm + dropout (:985-986).  `mask` / `post_mask` are dropout multipliers (0 or 1/(1-p)) or None.  The bf16 copy
    feeds the next tcgen05 linear without a cast kernel; its gradient arrives as a second (bf16) cotangent."""

    @staticmethod
    def forward(ctx, o, mask, resid, w, b, post_mask, rows, eps):
        from .engine import BF16, F32
        lib = L.load()
        L.require_cuda(resid, w, b)
        ctx.set_materialize_grads(False)       # an unused output (the last block's fp32 copy) arrives as None, not as a zero tensor
        H = w.shape[0]
        dev = w.device
        o2 = o.reshape(rows, H).contiguous() if o is not None else None
        r2 = resid.reshape(-1, H).float().contiguous()
        w32, b32 = w.detach().float().contiguous(), b.detach().float().contiguous()
        m2 = mask.reshape(rows, H).contiguous() if mask is not None else None
        pm2 = post_mask.reshape(-1, H).contiguous() if post_mask is not None else None
        y32 = torch.empty(rows, H, device=dev, dtype=F32)
        y16 = torch.empty(rows, H, device=dev, dtype=BF16)
        stats = torch.empty(rows, 2, device=dev, dtype=F32)
        L.check(lib.ta_add_layernorm_fwd(L.ptr(o2), L.ptr(m2), L.ptr(r2), r2.shape[0], L.ptr(w32), L.ptr(b32), L.ptr(pm2),
                                         pm2.shape[0] if pm2 is not None else 0, L.ptr(y32), L.ptr(y16), L.ptr(stats), rows, H, eps,
                                         L.stream_ptr()))
        ctx.save_for_backward(o2, m2, r2, w32, pm2, stats)
        ctx.meta = (rows, H, o.shape if o is not None else None, resid.shape, w.dtype, b.dtype, resid.dtype)
        return y32, y16

    @staticmethod
    def backward(ctx, g32, g16):
        from .engine import BF16, F32
        lib = L.load()
        if g32 is None and g16 is None:
            return (None,) * 8
        o2, m2, r2, w32, pm2, stats = ctx.saved_tensors
        rows, H, o_shape, r_shape, wdt, bdt, rdt = ctx.meta
        dev = w32.device
        g32c = g32.reshape(rows, H).float().contiguous() if g32 is not None else None
        g16c = g16.reshape(rows, H).contiguous() if g16 is not None else None
        d_o = torch.empty(rows, H, device=dev, dtype=BF16) if (o2 is not None and ctx.needs_input_grad[0]) else None
        d_r = torch.empty(r2.shape, device=dev, dtype=F32) if ctx.needs_input_grad[2] else None
        dw = torch.empty(H, device=dev, dtype=F32)
        db = torch.empty(H, device=dev, dtype=F32)
        scratch = torch.empty(lib.ta_add_layernorm_bwd_partial_floats(H), device=dev, dtype=F32)
        L.check(lib.ta_add_layernorm_bwd(L.ptr(g32c), L.ptr(g16c), L.ptr(o2), L.ptr(m2), L.ptr(r2), r2.shape[0], L.ptr(w32), L.ptr(pm2),
                                         pm2.shape[0] if pm2 is not None else 0, L.ptr(stats), L.ptr(d_o), L.ptr(d_r), L.ptr(dw), L.ptr(db),
                                         L.ptr(scratch), rows, H, L.stream_ptr()))
        return (d_o.view(o_shape) if d_o is not None else None, None, d_r.view(r_shape).to(rdt) if d_r is not None else None, dw.to(wdt),
                db.to(bdt), None, None, None)


class _GeluFn(torch.autograd.Function):
    """Exact-erf GELU on bf16 storage (Blip2QFormerIntermediate, hidden_act = "gelu"; HF:models/blip_2/modeling_blip_2.py:677-689)."""

    @staticmethod
    def forward(ctx, x):
        L.require_cuda(x)
        xc = x.contiguous()
        y = torch.empty_like(xc)
        L.check(L.load().ta_gelu_fwd_bf16(L.ptr(xc), L.ptr(y), xc.numel(), L.stream_ptr()))
        ctx.save_for_backward(xc)
        return y

    @staticmethod
    def backward(ctx, g):
        (xc,) = ctx.saved_tensors
        gc = g.to(torch.bfloat16).contiguous()
        dx = torch.empty_like(xc)
        L.check(L.load().ta_gelu_bwd_bf16(L.ptr(xc), L.ptr(gc), L.ptr(dx), xc.numel(), L.stream_ptr()))
        return dx


class QFormerAudioProjector(nn.Module):
    """BLIP-2 QFormer projector with learnable queries (reference: tiny_audio/projectors.py:359-475; arithmetic of
    HF:models/blip_2/modeling_blip_2.py:537-1042).  Parameter names and initialisation are the reference's (the HF
    `Blip2QFormerModel` is instantiated as the owner of the weights, exactly as the reference does), so checkpoints
    interchange.  Every linear -- q/k/v/o of self- and cross-attention, the FFN, the final projection; forward, dgrad and
    wgrad -- runs on the tcgen05 GEMM; the 3x3 / 3x15 softmax attention is one warp per (window, head) in csrc/window_attn.cu
    (forward and backward); dropout + residual + LayerNorm is one kernel each way and the exact GELU another (csrc/qformer_glue.cu);
    PyTorch supplies only the dropout masks (its Philox stream, so seeding behaves like nn.Dropout)."""

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
        # window attention kernel (csrc/window_attn.cu) instead of batched fp32 matmuls + softmax + transposes in PyTorch
        self.fused_attention = os.environ.get("TA_QFORMER_FUSED_ATTN", "1") != "0"

    def get_output_length(self, input_length):
        nblocks = (input_length + self.window_size - 1) // self.window_size
        return nblocks * self.num_queries

    def _drop_mask(self, rows: int, H: int, p: float, device):
        """[rows, H] fp32 multipliers 0 or 1/(1-p) from torch's RNG, or None when dropout is off."""
        if not self.training or p <= 0.0:
            return None
        return torch.nn.functional.dropout(torch.ones(rows, H, device=device, dtype=torch.float32), p, True)

    def _attend(self, att, x32, x16, kv16):
        """Blip2QFormerMultiHeadAttention + SelfOutput: x [W, q, H] queries (fp32 residual + its bf16 copy), kv16 [W, n, H] keys/values."""
        Wn, nq, H = x16.shape
        hd = H // self.num_heads
        q = tc_linear(x16, att.attention.query.weight, att.attention.query.bias)
        k = tc_linear(kv16, att.attention.key.weight, att.attention.key.bias)
        v = tc_linear(kv16, att.attention.value.weight, att.attention.value.bias)
        if self.fused_attention:
            ctx = window_attention(q, k, v, self.num_heads, self.p_attn, self.training)
        else:       # PyTorch glue (A/B reference for the kernel)
            F_ = torch.nn.functional
            q = q.view(Wn, nq, self.num_heads, hd).transpose(1, 2).float()
            k = k.view(Wn, -1, self.num_heads, hd).transpose(1, 2).float()
            v = v.view(Wn, -1, self.num_heads, hd).transpose(1, 2).float()
            probs = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(hd), dim=-1)
            probs = F_.dropout(probs, self.p_attn, self.training)
            ctx = (probs @ v).transpose(1, 2).reshape(Wn, nq, H).to(torch.bfloat16)
        o = tc_linear(ctx, att.output.dense.weight, att.output.dense.bias)
        y32, y16 = _AddLayerNormFn.apply(o, self._drop_mask(Wn * nq, H, self.p_hidden, o.device), x32, att.output.LayerNorm.weight,
                                         att.output.LayerNorm.bias, None, Wn * nq, self.ln_eps)
        return y32.view(Wn, nq, H), y16.view(Wn, nq, H)

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
        x_enc = x_enc.reshape(Wn, self.window_size, -1).to(torch.bfloat16)
        qf = self.qformer
        nq, H = self.num_queries, self.query.shape[-1]
        # the reference expands the query to every window BEFORE layernorm + dropout (projectors.py:461, HF :985-986): one LayerNorm row per
        # (window, query) with its own dropout mask; the kernel reads the 3 query rows through the row-broadcast residual slot
        x32, x16 = _AddLayerNormFn.apply(None, None, self.query.reshape(nq, H), qf.layernorm.weight, qf.layernorm.bias,
                                         self._drop_mask(Wn * nq, H, self.p_hidden, x_enc.device), Wn * nq, self.ln_eps)
        x32, x16 = x32.view(Wn, nq, H), x16.view(Wn, nq, H)
        for layer in qf.encoder.layer:
            x32, x16 = self._attend(layer.attention, x32, x16, x16)
            x32, x16 = self._attend(layer.crossattention, x32, x16, x_enc)
            h = tc_linear(x16, layer.intermediate_query.dense.weight, layer.intermediate_query.dense.bias)
            h = _GeluFn.apply(h)
            f = tc_linear(h, layer.output_query.dense.weight, layer.output_query.dense.bias)
            y32, y16 = _AddLayerNormFn.apply(f, self._drop_mask(Wn * nq, H, self.p_hidden, f.device), x32, layer.output_query.LayerNorm.weight,
                                             layer.output_query.LayerNorm.bias, None, Wn * nq, self.ln_eps)
            x32, x16 = y32.view(Wn, nq, H), y16.view(Wn, nq, H)
        out = tc_linear(x16.reshape(B, nblocks * nq, H), self.linear.weight, self.linear.bias)
        return out


class SimpleAdapter(nn.Module):
    """Parameter holder with the reference's names (`fc1`, `fc2`, both with bias; projectors.py:90-100).  The arithmetic does
    not run here: the mixture projectors below fold all adapters into two wide tcgen05 GEMMs."""

    def __init__(self, input_dim: int, hidden_dim: int, output_dim: int):
        super().__init__()
        self.fc1 = nn.Linear(input_dim, hidden_dim)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden_dim, output_dim)


def folded_adapter_mixture(x, adapters, gates):
    """sum_e gates[:, e] * adapter_e(x)  for 2-layer GELU adapters, as TWO GEMMs instead of 2 * E:

        h   = x @ [W1_0; W1_1; ...]^T + [b1_0, b1_1, ...]                    one GEMM, N = E * hidden
        out = [g_0 * gelu(h_0) | g_1 * gelu(h_1) | ...] @ [W2_0 | W2_1 | ...]^T  one GEMM, K = E * hidden
              + gates @ [b2_0; b2_1; ...]

    so the mixture is summed by the tensor core's fp32 accumulator (TMEM) rather than by E elementwise passes over the
    outputs.  A gate of exactly 0 reproduces a token that did not select the expert (sparse top-k routing, evaluated densely:
    no host sync on the routing decision, no ragged launches).  x [M, I] (any float dtype), gates [M, E] fp32.  Autograd
    flows through `tc_linear` (dgrad + wgrad on the same GEMM kernel), the concatenations and the gate product."""
    E = len(adapters)
    hid = adapters[0].fc1.out_features
    w1 = torch.cat([a.fc1.weight for a in adapters], dim=0)
    b1 = torch.cat([a.fc1.bias for a in adapters], dim=0)
    w2 = torch.cat([a.fc2.weight for a in adapters], dim=1)
    b2 = torch.stack([a.fc2.bias for a in adapters], dim=0).float()
    h = tc_linear(x, w1, b1)
    act = torch.nn.functional.gelu(h.float()).view(-1, E, hid) * gates.unsqueeze(-1)
    return tc_linear(act.view(-1, E * hid), w2).float() + gates @ b2


class MOSAProjector(nn.Module):
    """MOSA-Base projector (reference: tiny_audio/projectors.py:103-177): two k=3 / stride-2 convolutions with GELU (4x
    downsampling), a 2-layer ReLU router, and a dense softmax mixture of 4 two-layer GELU adapters.  Same parameter names and
    default initialisation as the reference (`downsampler.{0,2}`, `router.{0,2}`, `experts.{i}.fc{1,2}`), so checkpoints
    interchange.  On the B200 path each convolution is an im2col view + one tcgen05 GEMM (K = 3 * C_in), the router's wide layer
    is a GEMM, and the four adapters are folded into two GEMMs (`folded_adapter_mixture`); the 512 -> 4 router head, softmax
    and GELU are PyTorch glue (< 1 % of the projector's FLOPs)."""

    ADAPTER_HIDDEN_DIM = 4096
    ROUTER_HIDDEN_DIM = 512
    CONV_KERNEL = 3
    CONV_STRIDE = 2
    CONV_PADDING = 1

    def __init__(self, config):
        super().__init__()
        self.encoder_dim = getattr(config, "encoder_dim", None) or 1280
        self.llm_dim = getattr(config, "llm_dim", None) or 2048
        self.num_experts = getattr(config, "num_experts", None) or 4
        conv = dict(kernel_size=self.CONV_KERNEL, stride=self.CONV_STRIDE, padding=self.CONV_PADDING)
        self.downsampler = nn.Sequential(nn.Conv1d(self.encoder_dim, self.encoder_dim, **conv), nn.GELU(),
                                         nn.Conv1d(self.encoder_dim, self.llm_dim, **conv), nn.GELU())
        self.router = nn.Sequential(nn.Linear(self.llm_dim, self.ROUTER_HIDDEN_DIM), nn.ReLU(),
                                    nn.Linear(self.ROUTER_HIDDEN_DIM, self.num_experts))
        self.experts = nn.ModuleList([SimpleAdapter(self.llm_dim, self.ADAPTER_HIDDEN_DIM, self.llm_dim)
                                      for _ in range(self.num_experts)])

    def get_output_length(self, input_length):
        length = input_length
        for _ in range(2):
            length = (length + 2 * self.CONV_PADDING - self.CONV_KERNEL) // self.CONV_STRIDE + 1
        return length

    def _conv_gelu(self, x, conv):
        """Conv1d(k=3, s=2, p=1) + GELU on token-major activations: x [B, S, C] -> [B, S', C_out].  The im2col operand row for
        output step t is (c, tap) -> x[2t - 1 + tap, c], which is exactly `weight.view(C_out, C * 3)`'s column order."""
        F_ = torch.nn.functional
        B = x.shape[0]
        cols = F_.pad(x, (0, 0, self.CONV_PADDING, self.CONV_PADDING)).unfold(1, self.CONV_KERNEL, self.CONV_STRIDE)
        y = tc_linear(cols.reshape(B, cols.shape[1], -1), conv.weight.reshape(conv.weight.shape[0], -1), conv.bias)
        return F_.gelu(y.float())

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise L.TinyAudioB200Error("MOSAProjector runs only on CUDA (tiny_audio_b200 has no CPU fallback)")
        return self._mixture(x)

    def _mixture(self, x):
        F_ = torch.nn.functional
        x = self._conv_gelu(self._conv_gelu(x, self.downsampler[0]), self.downsampler[2])
        B, n, D = x.shape
        flat = x.reshape(B * n, D)
        hidden = F_.relu(tc_linear(flat, self.router[0].weight, self.router[0].bias).float())
        mix = torch.softmax(F_.linear(hidden, self.router[2].weight.float(), self.router[2].bias.float()), dim=-1)
        return folded_adapter_mixture(flat, list(self.experts), mix).view(B, n, -1)


class MoEAudioProjector(nn.Module):
    """Shared + sparse mixture-of-experts projector (reference: tiny_audio/projectors.py:185-351): frame-stack -> RMSNorm ->
    shared adapter + top-k of `num_experts` adapters, with the load-balance + z-loss auxiliary term exposed through
    `get_aux_loss()` (added to the LM loss by ASRModel.forward, asr_modeling.py:528-531).  Parameter names, defaults and
    `_init_weights` follow the reference.  The router (in_dim -> 4 logits, fp32 softmax, top-k, renormalisation with +1e-6,
    multiplicative jitter in training mode) is PyTorch glue on the device -- no `.any()` / `torch.where` host syncs as in the
    reference's dispatch loop; the shared adapter and the experts run as ONE folded mixture (gate 1 for the shared adapter,
    the renormalised top-k weight or exactly 0 for each expert): two tcgen05 GEMMs forward, four backward."""

    def __init__(self, config):
        super().__init__()
        self.k = getattr(config, "projector_pool_stride", 4)
        self.aux_coef = getattr(config, "router_aux_loss_coef", 0.01)
        self.router_z_loss_coef = getattr(config, "router_z_loss_coef", 1e-4)
        self.router_jitter_noise = getattr(config, "router_jitter_noise", 0.01)
        in_dim = config.encoder_dim * self.k
        out_dim = config.llm_dim
        hidden_dim = getattr(config, "projector_hidden_dim", None) or out_dim
        self.num_experts = getattr(config, "num_experts", 4)
        self.top_k = getattr(config, "num_experts_per_tok", 2)
        self.norm = _Gain(in_dim, 1e-6)
        self.router = nn.Linear(in_dim, self.num_experts, bias=False)
        self.experts = nn.ModuleList([SimpleAdapter(in_dim, hidden_dim, out_dim) for _ in range(self.num_experts)])
        self.shared_expert = SimpleAdapter(in_dim, hidden_dim, out_dim)
        with torch.no_grad():           # reference _init_weights (:242-251)
            nn.init.normal_(self.router.weight, mean=0.0, std=0.02)
            for adapter in [self.shared_expert, *self.experts]:
                nn.init.xavier_uniform_(adapter.fc1.weight)
                nn.init.normal_(adapter.fc2.weight, mean=0.0, std=0.01)
        self.last_aux_loss = torch.tensor(0.0)

    def get_output_length(self, input_length):
        return frame_stack_length(input_length, self.k)

    def get_aux_loss(self) -> torch.Tensor:
        return self.last_aux_loss

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise L.TinyAudioB200Error("MoEAudioProjector runs only on CUDA (tiny_audio_b200 has no CPU fallback)")
        return self._mixture(x)

    def _mixture(self, x):
        B, S, D = x.shape
        n = frame_stack_length(S, self.k)
        flat = x[:, : n * self.k].reshape(B * n, self.k * D).float()
        flat = self.norm.weight.float() * (flat * torch.rsqrt(flat.pow(2).mean(-1, keepdim=True) + self.norm.variance_epsilon))
        logits = torch.nn.functional.linear(flat, self.router.weight.float())
        if self.training and self.router_jitter_noise > 0:
            logits = logits * torch.empty_like(logits).uniform_(1.0 - self.router_jitter_noise, 1.0 + self.router_jitter_noise)
        probs = torch.softmax(logits, dim=-1, dtype=torch.float32)
        top_w, top_i = torch.topk(probs, self.top_k, dim=-1)
        top_w = top_w / (top_w.sum(dim=-1, keepdim=True) + 1e-6)
        if self.training:
            balance = self.aux_coef * ((probs.mean(0) - 1.0 / self.num_experts) ** 2).mean() * self.num_experts
            self.last_aux_loss = balance + self.router_z_loss_coef * torch.logsumexp(logits, dim=-1).pow(2).mean()
        else:
            self.last_aux_loss = torch.zeros((), device=x.device)
        # gates [M, 1 + E]: the shared adapter always on, each expert with its renormalised top-k weight or exactly 0
        gates = torch.cat([torch.ones_like(probs[:, :1]), torch.zeros_like(probs).scatter(1, top_i, top_w)], dim=1)
        out = folded_adapter_mixture(flat, [self.shared_expert, *self.experts], gates)
        return out.view(B, n, -1)


PROJECTOR_CLASSES = {
    "mlp": MLPAudioProjector,
    "mosa": MOSAProjector,
    "moe": MoEAudioProjector,
    "qformer": QFormerAudioProjector,
}
