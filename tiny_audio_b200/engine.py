"""Host-side driver of the B200 hot path: packs the frozen towers' weights into the layouts the CUDA
kernels want, owns the workspaces, and runs one fused forward+backward of

    waveform / mel -> GLM-ASR encoder -> MLP projector -> <audio> scatter -> Qwen3 -> CE -> d(projector params)

through libtinyaudio_b200.so.  PyTorch supplies device memory and the stream, nothing else.

Numerics recipe = the reference's production recipe (fp32 master weights + bf16 autocast,
configs/training/production.yaml:49): bf16 GEMM operands with fp32 accumulation, fp32 LayerNorm /
RMSNorm / softmax / CE, fp32 residual stream in the decoder, bf16 residual stream in the encoder
(conv output is bf16 under autocast), fp32 embedding lookup.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import lib as L

BF16 = torch.bfloat16
F32 = torch.float32


def _round_up(a: int, b: int) -> int:
    return (a + b - 1) // b * b


@dataclass
class PathDims:
    """Dimensions of the path (mirrors oracle.path_oracle.PathConfig; kept separate so the product
    never imports the oracle)."""
    n_mels: int = 128
    hop: int = 160
    enc_dim: int = 1280
    enc_ffn: int = 5120
    enc_layers: int = 32
    enc_heads: int = 20
    enc_rope_theta: float = 10000.0
    enc_partial_rotary: float = 0.5
    enc_ln_eps: float = 1e-5
    enc_max_pos: int = 1500
    proj_k: int = 4
    proj_hidden: int = 1024
    proj_eps: float = 1e-6
    lm_dim: int = 1024
    lm_ffn: int = 3072
    lm_layers: int = 28
    lm_heads: int = 16
    lm_kv_heads: int = 8
    lm_head_dim: int = 128
    lm_rope_theta: float = 1e6
    lm_eps: float = 1e-6
    lm_max_pos: int = 4096
    vocab: int = 151936
    audio_token_id: int = 151669

    @classmethod
    def from_any(cls, cfg) -> "PathDims":
        d = cfg if isinstance(cfg, dict) else {k: getattr(cfg, k) for k in cls.__dataclass_fields__ if hasattr(cfg, k)}
        return cls(**{k: v for k, v in d.items() if k in cls.__dataclass_fields__})


def _rope_tables(n_pos: int, dim: int, theta: float, device, round_bf16: bool):
    inv = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.int64).float() / dim))
    fr = torch.arange(n_pos).float()[:, None] * inv[None, :]
    cos, sin = fr.cos(), fr.sin()
    if round_bf16:       # the encoder's cos/sin are cast to the activation dtype (bf16 under autocast)
        cos, sin = cos.to(BF16).float(), sin.to(BF16).float()
    return cos.contiguous().to(device), sin.contiguous().to(device)


class PackedEncoder:
    """GLM-ASR encoder weights in kernel layout (state_dict names = HF GlmAsrEncoder)."""

    def __init__(self, sd: Dict[str, torch.Tensor], dims: PathDims, device):
        d = dims
        D = d.enc_dim
        self.dims = d
        self.keep = []

        def dev(t, dtype):
            t = t.detach().to(device=device, dtype=dtype).contiguous()
            self.keep.append(t)
            return t

        self.conv1_w = dev(sd["conv1.weight"].permute(0, 2, 1).reshape(D, 3 * d.n_mels), BF16)
        self.conv1_b = dev(sd["conv1.bias"], F32)
        self.conv2_w = dev(sd["conv2.weight"].permute(0, 2, 1).reshape(D, 3 * D), BF16)
        self.conv2_b = dev(sd["conv2.bias"], F32)
        self.lnf_w, self.lnf_b = dev(sd["norm.weight"], F32), dev(sd["norm.bias"], F32)
        rot = int((D // d.enc_heads) * d.enc_partial_rotary)
        self.rope_cos, self.rope_sin = _rope_tables(d.enc_max_pos, rot, d.enc_rope_theta, device, True)
        ptrs = []
        for i in range(d.enc_layers):
            p = f"layers.{i}."
            wq, wk, wv = (sd[p + f"self_attn.{n}_proj.weight"] for n in "qkv")
            bq, bv = sd[p + "self_attn.q_proj.bias"], sd[p + "self_attn.v_proj.bias"]
            layer = [None] * L.ENC_PTRS_PER_LAYER
            layer[L.ENC_LN1_W] = dev(sd[p + "input_layernorm.weight"], F32)
            layer[L.ENC_LN1_B] = dev(sd[p + "input_layernorm.bias"], F32)
            layer[L.ENC_WQKV] = dev(torch.cat([wq, wk, wv], 0), BF16)
            layer[L.ENC_BQKV] = dev(torch.cat([bq.float().cpu(), torch.zeros(D), bv.float().cpu()], 0), F32)
            layer[L.ENC_WO] = dev(sd[p + "self_attn.o_proj.weight"], BF16)
            layer[L.ENC_BO] = dev(sd[p + "self_attn.o_proj.bias"], F32)
            layer[L.ENC_LN2_W] = dev(sd[p + "post_attention_layernorm.weight"], F32)
            layer[L.ENC_LN2_B] = dev(sd[p + "post_attention_layernorm.bias"], F32)
            layer[L.ENC_W1] = dev(sd[p + "mlp.fc1.weight"], BF16)
            layer[L.ENC_B1] = dev(sd[p + "mlp.fc1.bias"], F32)
            layer[L.ENC_W2] = dev(sd[p + "mlp.fc2.weight"], BF16)
            layer[L.ENC_B2] = dev(sd[p + "mlp.fc2.bias"], F32)
            ptrs.extend(layer)
        self.table = L.pointer_table(ptrs)
        self.c = L.EncoderWeights(d.enc_layers, D, d.enc_ffn, d.enc_heads, D // d.enc_heads, rot, d.n_mels, d.enc_max_pos,
                                  d.enc_ln_eps, L.ptr(self.conv1_w), L.ptr(self.conv1_b), L.ptr(self.conv2_w),
                                  L.ptr(self.conv2_b), L.ptr(self.lnf_w), L.ptr(self.lnf_b), L.ptr(self.rope_cos),
                                  L.ptr(self.rope_sin), C.cast(self.table, C.POINTER(L.P)))


LORA_PAD = 128          # rank padding of the augmented-K LoRA GEMMs (csrc/engine.cu)
LORA_PROJS = ("q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj")


class PackedLM:
    """Qwen3 weights in kernel layout (state_dict names = HF Qwen3ForCausalLM).  With `lora=True` every projection weight
    (and its transposed dgrad copy) is stored stacked over layers with LORA_PAD extra K columns that `update_lora` fills
    with alpha/r * B (forward) resp. A^T (dgrad), so the adapted projection stays ONE GEMM."""

    def __init__(self, sd: Dict[str, torch.Tensor], dims: PathDims, device, lora: bool = False):
        d = dims
        self.dims = d
        self.keep = []
        self.lora_pad = LORA_PAD if lora else 0
        P = self.lora_pad
        self._vec = {}
        self.wgrad_flat = None

        def dev(t, dtype):
            t = t.detach().to(device=device, dtype=dtype).contiguous()
            self.keep.append(t)
            return t

        V, D, F, Lyr = d.vocab, d.lm_dim, d.lm_ffn, d.lm_layers
        QD, KD = d.lm_heads * d.lm_head_dim, d.lm_kv_heads * d.lm_head_dim
        QKV = QD + 2 * KD
        self.QD, self.KD, self.QKV = QD, KD, QKV
        self.vocab_pad = _round_up(V, 256)      # 256: lets the lm_head GEMMs use the 256-wide tile
        emb = sd["model.embed_tokens.weight"]
        assert emb.shape[0] == V, f"embedding rows {emb.shape[0]} != vocab {V}"
        self.embed_f32 = dev(emb, F32)
        head = sd.get("lm_head.weight", emb)         # Qwen3-0.6B ties lm_head to embed_tokens; an untied head is packed on its own
        self.tied_head = head.data_ptr() == emb.data_ptr() or (head.shape == emb.shape and bool(torch.equal(head, emb)))
        eb = torch.zeros(self.vocab_pad, D, dtype=BF16, device=device)
        eb[:V] = (self.embed_f32 if self.tied_head else head.detach().to(device=device, dtype=F32)).to(BF16)   # autocast casts the fp32 table to bf16
        self.embed_bf16 = eb
        self.embed_bf16_t = eb.t().contiguous()
        self.final_norm_w = dev(sd["model.norm.weight"], F32)
        self.rope_cos, self.rope_sin = _rope_tables(d.lm_max_pos, d.lm_head_dim, d.lm_rope_theta, device, False)
        assert F % 64 == 0

        def stacked(n_out, k_in):
            return torch.zeros(Lyr, n_out, k_in + P, dtype=BF16, device=device)

        # [W | aB] forward operands and [W^T | A^T] dgrad operands, stacked over layers
        self.w = {"qkv": stacked(QKV, D), "o": stacked(D, QD), "gu": stacked(2 * F, D), "d": stacked(D, F)}
        self.wt = {"qkv": stacked(D, QKV), "o": stacked(QD, D), "gu": stacked(D, 2 * F), "d": stacked(F, D)}
        self.kin = {"qkv": D, "o": QD, "gu": D, "d": F}
        self.nout = {"qkv": QKV, "o": D, "gu": 2 * F, "d": D}
        if P:
            self.lora_a = {g: torch.zeros(Lyr, P, self.kin[g], dtype=BF16, device=device) for g in self.kin}
            self.lora_bt = {g: torch.zeros(Lyr, P, self.nout[g], dtype=BF16, device=device) for g in self.kin}
            self.lora_da = {g: torch.zeros(Lyr, P, self.kin[g], dtype=F32, device=device) for g in self.kin}
            self.lora_db = {g: torch.zeros(Lyr, self.nout[g], P, dtype=F32, device=device) for g in self.kin}
        ptrs = []
        for i in range(Lyr):
            p = f"model.layers.{i}."
            wq, wk, wv = (sd[p + f"self_attn.{n}_proj.weight"] for n in "qkv")
            wqkv = torch.cat([wq, wk, wv], 0).to(device=device, dtype=BF16)
            wo = sd[p + "self_attn.o_proj.weight"].to(device=device, dtype=BF16)
            g = sd[p + "mlp.gate_proj.weight"].reshape(F // 64, 1, 64, D)
            u = sd[p + "mlp.up_proj.weight"].reshape(F // 64, 1, 64, D)
            wgu = torch.cat([g, u], 1).reshape(2 * F, D).to(device=device, dtype=BF16)   # 64-row blocks: gate, up, gate, up ...
            wd = sd[p + "mlp.down_proj.weight"].to(device=device, dtype=BF16)
            for name, mat in (("qkv", wqkv), ("o", wo), ("gu", wgu), ("d", wd)):
                self.w[name][i, :, : mat.shape[1]] = mat
                self.wt[name][i, :, : mat.shape[0]] = mat.t()
            layer = [None] * L.LM_PTRS_PER_LAYER
            layer[L.LM_LN1_W] = dev(sd[p + "input_layernorm.weight"], F32)
            layer[L.LM_WQKV], layer[L.LM_WQKV_T] = self.w["qkv"][i], self.wt["qkv"][i]
            layer[L.LM_QNORM_W] = dev(sd[p + "self_attn.q_norm.weight"], F32)
            layer[L.LM_KNORM_W] = dev(sd[p + "self_attn.k_norm.weight"], F32)
            layer[L.LM_WO], layer[L.LM_WO_T] = self.w["o"][i], self.wt["o"][i]
            layer[L.LM_LN2_W] = dev(sd[p + "post_attention_layernorm.weight"], F32)
            layer[L.LM_WGU], layer[L.LM_WGU_T] = self.w["gu"][i], self.wt["gu"][i]
            layer[L.LM_WD], layer[L.LM_WD_T] = self.w["d"][i], self.wt["d"][i]
            if P:
                layer[L.LM_LORA_A_QKV], layer[L.LM_LORA_A_O] = self.lora_a["qkv"][i], self.lora_a["o"][i]
                layer[L.LM_LORA_A_GU], layer[L.LM_LORA_A_D] = self.lora_a["gu"][i], self.lora_a["d"][i]
                layer[L.LM_LORA_BT_QKV], layer[L.LM_LORA_BT_O] = self.lora_bt["qkv"][i], self.lora_bt["o"][i]
                layer[L.LM_LORA_BT_GU], layer[L.LM_LORA_BT_D] = self.lora_bt["gu"][i], self.lora_bt["d"][i]
            for slot in (L.LM_LN1_W, L.LM_QNORM_W, L.LM_KNORM_W, L.LM_LN2_W):
                self._vec[(i, slot)] = layer[slot]
            ptrs.extend(layer)
        self.table = L.pointer_table(ptrs)
        self.grad_table = None
        if P:
            gp = []
            for i in range(Lyr):
                for g in ("qkv", "o", "gu", "d"):
                    gp.extend([self.lora_da[g][i], self.lora_db[g][i]])
            self.grad_table = L.pointer_table(gp)
        self.c = L.LmWeights(Lyr, D, F, d.lm_heads, d.lm_kv_heads, d.lm_head_dim, d.lm_max_pos, V, self.vocab_pad,
                             d.lm_eps, P, L.ptr(self.embed_f32), L.ptr(self.embed_bf16), L.ptr(self.embed_bf16_t),
                             L.ptr(self.final_norm_w), L.ptr(self.rope_cos), L.ptr(self.rope_sin),
                             C.cast(self.table, C.POINTER(L.P)))

    # ------------------------------------------------------------------ unfrozen LM
    _MATS = (("self_attn.q_proj", "qkv", 0), ("self_attn.k_proj", "qkv", 1), ("self_attn.v_proj", "qkv", 2), ("self_attn.o_proj", "o", 0),
             ("mlp.gate_proj", "gu", 0), ("mlp.up_proj", "gu", 1), ("mlp.down_proj", "d", 0))

    def enable_weight_grads(self):
        """Allocate the engine's weight-gradient outputs as views of ONE flat fp32 buffer (a single NCCL all-reduce covers it)
        and the per-layer pointer table ta_lm_forward_backward fills.  Layouts follow the packed operands (header)."""
        if getattr(self, "wgrad_flat", None) is not None:
            return
        assert self.lora_pad == 0, "LoRA adapters and an unfrozen LM are mutually exclusive"
        if not self.tied_head:
            raise L.TinyAudioB200Error("unfrozen decoder with an untied lm_head is not supported (Qwen3-0.6B ties it to embed_tokens)")
        d = self.dims
        Lyr, D, F = d.lm_layers, d.lm_dim, d.lm_ffn
        shapes = {"qkv": (Lyr, self.QKV, D), "o": (Lyr, D, self.QD), "gu": (Lyr, 2 * F, D), "d": (Lyr, D, F), "ln1": (Lyr, D),
                  "ln2": (Lyr, D), "qn": (Lyr, d.lm_head_dim), "kn": (Lyr, d.lm_head_dim), "fnorm": (D,), "embed": (self.vocab_pad, D)}
        total = sum(math.prod(v) for v in shapes.values())
        self.wgrad_flat = torch.zeros(total, dtype=F32, device=self.embed_f32.device)
        self.wgrad, off = {}, 0
        for k, shp in shapes.items():
            n = math.prod(shp)
            self.wgrad[k] = self.wgrad_flat[off: off + n].view(*shp)
            off += n
        small = sum(math.prod(shapes[k]) for k in ("ln1", "ln2", "qn", "kn", "fnorm"))
        o0 = sum(math.prod(shapes[k]) for k in ("qkv", "o", "gu", "d"))
        self.wgrad_small = self.wgrad_flat[o0: o0 + small]          # the accumulated (atomic) outputs: zeroed before every step
        gp = []
        for i in range(Lyr):
            gp.extend([self.wgrad[k][i] for k in ("qkv", "o", "gu", "d", "ln1", "ln2", "qn", "kn")])
        self.wgrad_table = L.pointer_table(gp)

    def hf_grads(self) -> Dict[str, torch.Tensor]:
        """The weight gradients under HF Qwen3ForCausalLM parameter names.  q/k/v/o/down/norm/embedding gradients are views of the
        flat buffer; gate/up are de-interleaved copies (two strided copies per step)."""
        d = self.dims
        F, D, Lyr = d.lm_ffn, d.lm_dim, d.lm_layers
        g = self.wgrad
        gu = g["gu"].view(Lyr, F // 64, 2, 64, D)
        gate = gu[:, :, 0].reshape(Lyr, F, D)
        up = gu[:, :, 1].reshape(Lyr, F, D)
        out = {"model.embed_tokens.weight": g["embed"][: d.vocab], "model.norm.weight": g["fnorm"]}
        for i in range(Lyr):
            p = f"model.layers.{i}."
            out[p + "self_attn.q_proj.weight"] = g["qkv"][i, : self.QD]
            out[p + "self_attn.k_proj.weight"] = g["qkv"][i, self.QD: self.QD + self.KD]
            out[p + "self_attn.v_proj.weight"] = g["qkv"][i, self.QD + self.KD:]
            out[p + "self_attn.o_proj.weight"] = g["o"][i]
            out[p + "mlp.gate_proj.weight"] = gate[i]
            out[p + "mlp.up_proj.weight"] = up[i]
            out[p + "mlp.down_proj.weight"] = g["d"][i]
            out[p + "input_layernorm.weight"] = g["ln1"][i]
            out[p + "post_attention_layernorm.weight"] = g["ln2"][i]
            out[p + "self_attn.q_norm.weight"] = g["qn"][i]
            out[p + "self_attn.k_norm.weight"] = g["kn"][i]
        return out

    @torch.no_grad()
    def refresh_from(self, sd: Dict[str, torch.Tensor]):
        """Re-pack the bf16 operands (and their transposed dgrad copies) from the fp32 master weights after an optimiser step:
        one ta_pack_weight launch per matrix.  fp32 vectors (norm gains) and the fp32 embedding table are used in place when the
        masters already live on this device (no copies), else they are copied."""
        d = self.dims
        lib = L.load()
        st = L.stream_ptr()
        F, D = d.lm_ffn, d.lm_dim
        row_off = {"qkv": (0, self.QD, self.QD + self.KD)}
        for i in range(d.lm_layers):
            p = f"model.layers.{i}."
            for name, grp, j in self._MATS:
                src = sd[p + name + ".weight"]
                assert src.is_cuda and src.dtype == F32 and src.is_contiguous(), f"{p + name}: fp32 contiguous CUDA master expected"
                R, Cc = src.shape
                if grp == "gu":
                    blk, stride, off = 64, 128, 64 * j
                else:
                    blk, stride, off = R, R, (row_off["qkv"][j] if grp == "qkv" else 0)
                w, wt = self.w[grp][i], self.wt[grp][i]
                L.check(lib.ta_pack_weight(L.ptr(src), R, Cc, L.ptr(w), w.stride(0), L.ptr(wt), wt.stride(0), blk, stride, off, st))
        emb = sd["model.embed_tokens.weight"]
        L.check(lib.ta_pack_weight(L.ptr(emb), emb.shape[0], D, L.ptr(self.embed_bf16), D, L.ptr(self.embed_bf16_t), self.vocab_pad,
                                   emb.shape[0], emb.shape[0], 0, st))
        if emb.data_ptr() != self.embed_f32.data_ptr():
            self.embed_f32.copy_(emb)
        # norm gains: the kernels read fp32 vectors; alias-or-copy
        names = [(L.LM_LN1_W, "input_layernorm.weight"), (L.LM_QNORM_W, "self_attn.q_norm.weight"), (L.LM_KNORM_W, "self_attn.k_norm.weight"),
                 (L.LM_LN2_W, "post_attention_layernorm.weight")]
        for i in range(d.lm_layers):
            for slot, nm in names:
                src = sd[f"model.layers.{i}.{nm}"]
                dst = self._vec[(i, slot)]
                if src.data_ptr() != dst.data_ptr():
                    dst.copy_(src)
        if sd["model.norm.weight"].data_ptr() != self.final_norm_w.data_ptr():
            self.final_norm_w.copy_(sd["model.norm.weight"])

    # adapter j of a fused projection owns rank rows / columns [8 j, 8 j + r)
    _GROUPS = {"qkv": ("q_proj", "k_proj", "v_proj"), "o": ("o_proj",), "gu": ("gate_proj", "up_proj"), "d": ("down_proj",)}

    def _row_slices(self, g):
        d = self.dims
        F = d.lm_ffn
        if g == "qkv":
            return [slice(0, self.QD), slice(self.QD, self.QD + self.KD), slice(self.QD + self.KD, self.QKV)]
        if g == "gu":
            return ["gate", "up"]          # interleaved 64-row blocks, handled by a view
        return [slice(0, self.nout[g])]

    @torch.no_grad()
    def update_lora(self, lora_a: Dict[str, torch.Tensor], lora_b: Dict[str, torch.Tensor], scaling: float):
        """lora_a[proj]: [L, r, in] fp32, lora_b[proj]: [L, out, r] fp32 (stacked over layers).  Refreshes the rank-P operands
        and the extra K columns of the augmented weights (a handful of strided copies per step)."""
        P = self.lora_pad
        assert P, "PackedLM was built without LoRA"
        F = self.dims.lm_ffn
        for g, projs in self._GROUPS.items():
            K, N = self.kin[g], self.nout[g]
            A, BT = self.lora_a[g], self.lora_bt[g]
            for j, pj in enumerate(projs):
                a, bm = lora_a[pj], lora_b[pj]
                r = a.shape[1]
                assert r <= 8, "rank > 8 needs a wider block layout"
                A[:, 8 * j: 8 * j + r] = a.to(BF16)
                bs = (bm.float() * scaling).to(BF16).transpose(1, 2)       # [L, r, out]
                if g == "qkv":
                    rs = self._row_slices(g)[j]
                    BT[:, 8 * j: 8 * j + r, rs] = bs
                elif g == "gu":
                    v = BT.view(BT.shape[0], P, F // 64, 2, 64)             # columns in (block, gate|up, 64) order
                    v[:, 8 * j: 8 * j + r, :, j, :] = bs.reshape(bs.shape[0], r, F // 64, 64)
                else:
                    BT[:, 8 * j: 8 * j + r] = bs
            self.w[g][:, :, K: K + P] = BT.transpose(1, 2)                   # [W | aB]
            self.wt[g][:, :, N: N + P] = A.transpose(1, 2)                   # [W^T | A^T]

    def lora_grads(self, scaling: float, ranks: Dict[str, int]):
        """Unpack the engine's gradient outputs into per-projection stacked gradients (views / small copies)."""
        F = self.dims.lm_ffn
        ga, gb = {}, {}
        for g, projs in self._GROUPS.items():
            dA, dB = self.lora_da[g], self.lora_db[g]                       # [L, P, K], [L, N, P]
            for j, pj in enumerate(projs):
                r = ranks[pj]
                ga[pj] = dA[:, 8 * j: 8 * j + r]
                cols = dB[:, :, 8 * j: 8 * j + r] * scaling                  # d/dB = alpha/r * d/d(aB)
                if g == "qkv":
                    gb[pj] = cols[:, self._row_slices(g)[j]]
                elif g == "gu":
                    gb[pj] = cols.reshape(cols.shape[0], F // 64, 2, 64, r)[:, :, j].reshape(cols.shape[0], F, r)
                else:
                    gb[pj] = cols
        return ga, gb


class _Workspaces:
    """Grow-only device scratch keyed by name (allocated through torch's caching allocator)."""

    def __init__(self, device):
        self.device = device
        self.bufs: Dict[str, torch.Tensor] = {}

    def get(self, name: str, nbytes: int) -> torch.Tensor:
        b = self.bufs.get(name)
        if b is None or b.numel() < nbytes:
            self.bufs[name] = b = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return b

    def typed(self, name: str, shape, dtype) -> torch.Tensor:
        n = int(math.prod(shape))
        b = self.get(name, n * torch.empty((), dtype=dtype).element_size())
        return b[: n * torch.empty((), dtype=dtype).element_size()].view(dtype).view(*shape)


def label_rows_and_targets(labels_cpu: torch.Tensor):
    """Shift-by-one on the host (HF:loss/loss_utils.py:56-59): position p predicts labels[p+1].
    Returns flat row indices (int32) and targets (int32) of the non-ignored positions."""
    B, S = labels_cpu.shape
    shift = torch.full_like(labels_cpu, -100)
    shift[:, :-1] = labels_cpu[:, 1:]
    flat = shift.reshape(-1)
    rows = (flat != -100).nonzero().squeeze(-1)
    return rows.to(torch.int32), flat[rows].to(torch.int32)


class HotPath:
    """One GPU's replica of the frozen towers + the per-step driver."""

    def label_rows(self, labels: Optional[torch.Tensor]):
        """(rows int32 [n], targets int32 [n], n) of the positions that carry a label after the shift-by-one
        (HF:loss/loss_utils.py:56-59).  Host labels: built on the host, copied asynchronously (no device sync).  Device labels
        (what HF Trainer hands over): compacted on the device by ta_label_rows; only the 4-byte count is read back (it sizes
        the lm_head product) -- no labels.cpu(), no host-side nonzero()."""
        if labels is None:
            e = torch.empty(0, dtype=torch.int32, device=self.device)
            return e, e, 0
        if not labels.is_cuda:
            rows, targets = label_rows_and_targets(labels)
            return rows.to(self.device, non_blocking=True), targets.to(self.device, non_blocking=True), int(rows.numel())
        B, S = labels.shape
        lab = labels.to(device=self.device, dtype=torch.int64).contiguous()
        rows = self.ws.typed("label_rows", (B * S,), torch.int32)
        targets = self.ws.typed("label_targets", (B * S,), torch.int32)
        count = self.ws.typed("label_count", (1,), torch.int32)
        L.check(self.lib.ta_label_rows(L.ptr(lab), B, S, L.ptr(rows), L.ptr(targets), L.ptr(count), L.stream_ptr()))
        n = int(count.item())
        return rows[:n], targets[:n], n

    def apply_frame_dropout(self, enc: torch.Tensor, keep_prob: Optional[float] = None, keep_mask: Optional[torch.Tensor] = None):
        """audio_token_dropout (asr_modeling.py:458-479): whole-frame Bernoulli(keep_prob) zero mask on the encoder output, no
        rescale.  The mask is drawn with torch's generator exactly as the reference draws it (same shape, fp32, same device), so
        `torch.manual_seed` reproduces the reference's CUDA mask; `keep_mask` ([B, S_e], 1 = keep) injects a given draw (parity
        tests replay the mask the reference drew).  Applied in place by ta_frame_keep_mask."""
        if keep_mask is None:
            if keep_prob is None or keep_prob >= 1.0:
                return enc
            keep_mask = torch.bernoulli(torch.full(enc.shape[:-1], float(keep_prob), device=self.device, dtype=F32))
        keep = keep_mask.to(device=self.device, dtype=F32).contiguous()
        assert tuple(keep.shape) == tuple(enc.shape[:-1]), f"keep mask {tuple(keep.shape)} vs encoder frames {tuple(enc.shape[:-1])}"
        L.check(self.lib.ta_frame_keep_mask(L.ptr(enc), L.ptr(keep), keep.numel(), enc.shape[-1], L.stream_ptr()))
        return enc

    def __init__(self, dims: PathDims, enc_sd, lm_sd, device="cuda", lora: bool = False):
        self.lib = L.load()
        self.dims = dims
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise L.TinyAudioB200Error("HotPath needs a CUDA device (no CPU fallback)")
        self.enc = PackedEncoder(enc_sd, dims, self.device)
        self.lm = PackedLM(lm_sd, dims, self.device, lora=lora)
        self.ws = _Workspaces(self.device)
        self.launches = 0

    # ------------------------------------------------------------------ front end + encoder
    def logmel(self, wave: torch.Tensor, want_f32: bool = False):
        """wave (B, L) fp32 CUDA -> (conv1 im2col bf16 [B*T, 3*n_mels], optional fp32 (B,128,T))."""
        L.require_cuda(wave)
        assert wave.dtype == F32 and wave.dim() == 2 and wave.stride(1) == 1
        B, Ls = wave.shape
        T = Ls // self.dims.hop
        n = C.c_longlong()
        L.check(self.lib.ta_logmel_workspace_floats(B, Ls, C.byref(n)))
        ws = self.ws.typed("logmel", (n.value,), F32)
        im2 = self.ws.typed("conv1_im2col", (B * T, 3 * self.dims.n_mels), BF16)
        out = torch.empty(B, self.dims.n_mels, T, device=self.device, dtype=F32) if want_f32 else None
        L.check(self.lib.ta_logmel_fwd(L.ptr(wave), wave.stride(0), B, Ls, L.ptr(ws), L.ptr(out), L.ptr(im2), L.stream_ptr()))
        return im2, out, T

    def mel_to_im2col(self, mel: torch.Tensor):
        L.require_cuda(mel)
        mel = mel.to(F32).contiguous()
        B, nm, T = mel.shape
        assert nm == self.dims.n_mels
        im2 = self.ws.typed("conv1_im2col", (B * T, 3 * nm), BF16)
        L.check(self.lib.ta_mel_to_conv1_im2col(L.ptr(mel), B, T, L.ptr(im2), L.stream_ptr()))
        return im2, T

    def encode(self, im2col: torch.Tensor, B: int, T: int) -> torch.Tensor:
        """-> encoder last_hidden_state, bf16 [B, S_e, enc_dim]."""
        S = (T + 2 - 3) // 2 + 1
        n = C.c_longlong()
        L.check(self.lib.ta_encoder_workspace_bytes(C.byref(self.enc.c), B, T, C.byref(n)))
        ws = self.ws.get("encoder", n.value)
        out = self.ws.typed("enc_out", (B, S, self.dims.enc_dim), BF16)
        L.check(self.lib.ta_encoder_forward(C.byref(self.enc.c), L.ptr(im2col), B, T, L.ptr(ws), n.value, L.ptr(out),
                                            L.stream_ptr()))
        return out

    # ------------------------------------------------------------------ projector
    def _proj_struct(self, params: Dict[str, torch.Tensor], need_t: bool):
        d = self.dims
        H = params["linear_1.weight"].shape[0]
        w1 = self.ws.typed("proj_w1", (H, d.proj_k * d.enc_dim), BF16)
        w2 = self.ws.typed("proj_w2", (d.lm_dim, H), BF16)
        st = L.stream_ptr()
        p1, p2 = params["linear_1.weight"], params["linear_2.weight"]
        L.check(self.lib.ta_cast_f32_bf16(L.ptr(p1), L.ptr(w1), p1.numel(), st))
        L.check(self.lib.ta_cast_f32_bf16(L.ptr(p2), L.ptr(w2), p2.numel(), st))
        w2t = None
        if need_t:
            w2t = self.ws.typed("proj_w2t", (H, d.lm_dim), BF16)
            L.check(self.lib.ta_transpose_bf16(L.ptr(w2), L.ptr(w2t), d.lm_dim, H, H, d.lm_dim, st))
        n1, n2 = params["norm.weight"], params["norm_2.weight"]
        self._proj_keep = (w1, w2, w2t, n1, n2)
        return L.MlpProjectorWeights(d.proj_k * d.enc_dim, H, d.lm_dim, d.proj_eps, L.ptr(w1), L.ptr(n1), L.ptr(w2),
                                     L.ptr(w2t), L.ptr(n2)), H

    def frame_stack(self, enc_out: torch.Tensor):
        B, S, D = enc_out.shape
        k = self.dims.proj_k
        n = (S - k) // k + 1
        if n * k == S:
            return enc_out.view(B * n, k * D), n          # free view (tiny_audio/projectors.py:87)
        out = self.ws.typed("stacked", (B * n, k * D), BF16)
        L.check(self.lib.ta_frame_stack(L.ptr(enc_out), L.ptr(out), B, S, n, k, D, L.stream_ptr()))
        return out, n

    def projector_forward(self, x_stacked: torch.Tensor, params, need_bwd: bool):
        for k in ("linear_1.weight", "norm.weight", "linear_2.weight", "norm_2.weight"):
            t = params[k]
            L.require_cuda(t)
            assert t.dtype == F32 and t.is_contiguous(), f"projector param {k} must be contiguous fp32"
        pw, H = self._proj_struct(params, need_bwd)
        M = x_stacked.shape[0]
        d = self.dims
        y1 = self.ws.typed("proj_y1", (M, H), BF16)
        a1 = self.ws.typed("proj_a1", (M, H), BF16)
        y2 = self.ws.typed("proj_y2", (M, d.lm_dim), BF16)
        out = self.ws.typed("proj_out", (M, d.lm_dim), F32)
        L.check(self.lib.ta_mlp_projector_forward(C.byref(pw), L.ptr(x_stacked), M, L.ptr(y1), L.ptr(a1), L.ptr(y2), L.ptr(out),
                                                  L.stream_ptr()))
        return out, (pw, x_stacked, M, y1, a1, y2, H)

    def projector_backward(self, stash, d_out: torch.Tensor, grads: Dict[str, torch.Tensor]):
        pw, x_stacked, M, y1, a1, y2, H = stash
        n = C.c_longlong()
        L.check(self.lib.ta_mlp_projector_backward_workspace_bytes(C.byref(pw), M, C.byref(n)))
        ws = self.ws.get("proj_bwd", n.value)
        grads["norm.weight"].zero_()
        grads["norm_2.weight"].zero_()
        L.check(self.lib.ta_mlp_projector_backward(C.byref(pw), L.ptr(x_stacked), M, L.ptr(y1), L.ptr(a1), L.ptr(y2), L.ptr(d_out),
                                                   L.ptr(ws), n.value, L.ptr(grads["linear_1.weight"]), L.ptr(grads["norm.weight"]),
                                                   L.ptr(grads["linear_2.weight"]), L.ptr(grads["norm_2.weight"]), L.stream_ptr()))

    # ------------------------------------------------------------------ embed + scatter
    def embed_scatter(self, input_ids: torch.Tensor, counts: torch.Tensor, audio: torch.Tensor, n_a: int):
        B, S = input_ids.shape
        d = self.dims
        src = self.ws.typed("src_row", (B * S,), torch.int32)
        emb = self.ws.typed("inputs_embeds", (B * S, d.lm_dim), F32)
        st = L.stream_ptr()
        L.check(self.lib.ta_audio_index(L.ptr(input_ids), L.ptr(counts), L.ptr(src), B, S, n_a, d.audio_token_id, st))
        L.check(self.lib.ta_embed_scatter(L.ptr(input_ids), L.ptr(src), L.ptr(self.lm.embed_f32), L.ptr(audio), L.ptr(emb),
                                          B * S, d.lm_dim, d.vocab, st))
        return emb, src

    # ------------------------------------------------------------------ decoder + loss (+ backward to inputs_embeds)
    def lm_step(self, emb: torch.Tensor, B: int, S: int, rows: torch.Tensor, targets: torch.Tensor, inv_items: float,
                with_backward: bool, want_row_loss: bool = False, train_lm: bool = False, input_ids: Optional[torch.Tensor] = None,
                want_hidden: bool = False):
        """train_lm (unfrozen LM): the backward also fills self.lm.wgrad (PackedLM.enable_weight_grads) -- needs input_ids.
        want_hidden: also keep the last layer's output (before the final norm) in the `final_hidden` workspace, from which
        logits of arbitrary positions can be produced afterwards (HotPath.logits_all)."""
        d = self.dims
        nl = int(rows.numel())
        n = C.c_longlong()
        train_lm = bool(train_lm and with_backward)
        if train_lm:
            self.lm.enable_weight_grads()
            self.lm.wgrad_small.zero_()
        L.check(self.lib.ta_lm_workspace_bytes(C.byref(self.lm.c), B, S, nl, 2 if train_lm else int(with_backward), C.byref(n)))
        ws = self.ws.get("lm", n.value)
        loss = torch.zeros(1, device=self.device, dtype=F32)
        demb = self.ws.typed("d_inputs_embeds", (B * S, d.lm_dim), F32) if with_backward else None
        row_loss = torch.empty(nl, device=self.device, dtype=F32) if want_row_loss else None
        hid = self.ws.typed("final_hidden", (B * S, d.lm_dim), F32) if want_hidden else None
        lora_zeroed = 0
        if with_backward and self.lm.grad_table is not None:        # LoRA: clear the 8 stacked gradient tensors once, not per product
            for g in self.lm.lora_da:
                self.lm.lora_da[g].zero_()
                self.lm.lora_db[g].zero_()
            lora_zeroed = 1
        args = L.LmStepArgs(B, S, nl, int(with_backward), L.ptr(emb), L.ptr(rows), L.ptr(targets), inv_items, L.ptr(loss),
                            L.ptr(row_loss), L.ptr(demb), L.ptr(ws), n.value, L.ptr(hid),
                            C.cast(self.lm.grad_table, C.POINTER(L.P)) if (with_backward and self.lm.grad_table is not None) else None,
                            None, None, 0,
                            C.cast(self.lm.wgrad_table, C.POINTER(L.P)) if train_lm else None,
                            L.ptr(self.lm.wgrad["embed"]) if train_lm else None, L.ptr(self.lm.wgrad["fnorm"]) if train_lm else None,
                            L.ptr(input_ids) if train_lm else None, d.audio_token_id, None, None, lora_zeroed)
        L.check(self.lib.ta_lm_forward_backward(C.byref(self.lm.c), C.byref(args), L.stream_ptr()))
        return loss, demb, row_loss

    def lm_hidden(self, emb: torch.Tensor, B: int, S: int, kv_cache=None, position_ids: Optional[torch.Tensor] = None,
                  kv_start: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Forward-only decoder pass; returns the last layer's output before the final norm, fp32 [B*S, dim].
        `kv_cache` = (k, v, max_seq): the prompt's roped keys / values of every layer are stored in rows [0, S) (prefill).
        `position_ids` int32 [B, S] / `kv_start` int32 [B]: left-padded prompts (rotary positions counted from each sequence's first
        real token; padding keys masked)."""
        d = self.dims
        n = C.c_longlong()
        L.check(self.lib.ta_lm_workspace_bytes(C.byref(self.lm.c), B, S, 0, 0, C.byref(n)))
        ws = self.ws.get("lm", n.value)
        loss = torch.zeros(1, device=self.device, dtype=F32)
        hid = self.ws.typed("final_hidden", (B * S, d.lm_dim), F32)
        kc, vc, ms = kv_cache if kv_cache is not None else (None, None, 0)
        args = L.LmStepArgs(B, S, 0, 0, L.ptr(emb), None, None, 1.0, L.ptr(loss), None, None, L.ptr(ws), n.value, L.ptr(hid), None,
                            L.ptr(kc), L.ptr(vc), ms, None, None, None, None, 0, L.ptr(position_ids), L.ptr(kv_start), 0)
        L.check(self.lib.ta_lm_forward_backward(C.byref(self.lm.c), C.byref(args), L.stream_ptr()))
        return hid

    def new_kv_cache(self, B: int, max_seq: int):
        """bf16 [layers, B, max_seq, Hkv*head_dim] x 2 (keys are stored normed + roped)."""
        d = self.dims
        shape = (d.lm_layers, B, max_seq, d.lm_kv_heads * d.lm_head_dim)
        return (torch.empty(shape, device=self.device, dtype=BF16), torch.empty(shape, device=self.device, dtype=BF16), max_seq)

    def _decode_graph(self, fed, nxt_buf, pos_dev, cache, logits, pos_host: int, kv_start=None):
        """CUDA graph of one decode step over the given (persistent, workspace-owned) buffers; cached per pointer set."""
        key = (int(fed.numel()), cache[2], fed.data_ptr(), nxt_buf.data_ptr(), pos_dev.data_ptr(), cache[0].data_ptr(),
               cache[1].data_ptr(), logits.data_ptr(), kv_start.data_ptr() if kv_start is not None else 0)
        graphs = self.__dict__.setdefault("_decode_graphs", {})
        g = graphs.get(key)
        if g is None:
            n = C.c_longlong()
            L.check(self.lib.ta_lm_decode_workspace_bytes(C.byref(self.lm.c), int(fed.numel()), C.byref(n)))
            self.ws.get("lm_decode", n.value)                  # allocate outside the capture
            keep = pos_dev.clone()
            self.decode_step(fed, pos_dev, pos_host, cache, logits, nxt_buf, kv_start)      # eager warm-up (module load) before the capture;
            pos_dev.copy_(keep)                                                   # its cache row is rewritten by the real step
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                with torch.cuda.graph(g, stream=side):
                    self.decode_step(fed, pos_dev, pos_host, cache, logits, nxt_buf, kv_start)
            torch.cuda.current_stream().wait_stream(side)
            pos_dev.copy_(keep)
            graphs.clear()                                     # one live graph per HotPath is enough
            graphs[key] = g
        return g

    def decode_step(self, ids: torch.Tensor, pos_dev: torch.Tensor, pos_host: int, kv_cache, logits: torch.Tensor,
                    next_ids: torch.Tensor, kv_start: Optional[torch.Tensor] = None):
        """One KV-cache decode step (ta_lm_decode_step): feeds ids [B] at cache row *pos_dev, writes the bf16 logits
        [B, vocab_pad] and next_ids [B] = argmax, and advances *pos_dev on the device.  kv_start int32 [B]: first real cache row
        of each left-padded sequence."""
        B = int(ids.numel())
        kc, vc, ms = kv_cache
        n = C.c_longlong()
        L.check(self.lib.ta_lm_decode_workspace_bytes(C.byref(self.lm.c), B, C.byref(n)))
        ws = self.ws.get("lm_decode", n.value)
        L.check(self.lib.ta_lm_decode_step(C.byref(self.lm.c), L.ptr(ids), L.ptr(pos_dev), int(pos_host), L.ptr(kc), L.ptr(vc), ms, B,
                                           L.ptr(ws), n.value, L.ptr(logits), L.ptr(next_ids), L.ptr(kv_start), L.stream_ptr()))

    def logits_rows(self, hidden: torch.Tensor, rows: torch.Tensor) -> torch.Tensor:
        """final norm + tied lm_head on the given flat token rows -> bf16 logits [n_rows, vocab] (padding sliced off)."""
        d = self.dims
        nr = int(rows.numel())
        rows = rows.to(device=self.device, dtype=torch.int32).contiguous()
        normed = self.ws.typed("gen_normed", (nr, d.lm_dim), BF16)
        logits = torch.empty(nr, self.lm.vocab_pad, device=self.device, dtype=BF16)
        L.check(self.lib.ta_lm_hidden_to_logits(C.byref(self.lm.c), L.ptr(hidden), L.ptr(rows), nr, L.ptr(normed), L.ptr(logits),
                                                L.stream_ptr()))
        return logits[:, : d.vocab]

    def logits_all(self, B: int, S: int, hidden: Optional[torch.Tensor] = None) -> torch.Tensor:
        """bf16 logits [B, S, vocab] of every position from the kept final hidden states (lm_step(want_hidden=True) or lm_hidden):
        what the reference's forward returns as `outputs.logits` (asr_modeling.py:517-533; bf16 under the autocast recipe)."""
        hid = hidden if hidden is not None else self.ws.typed("final_hidden", (B * S, self.dims.lm_dim), F32)
        rows = torch.arange(B * S, device=self.device, dtype=torch.int32)
        return self.logits_rows(hid, rows).view(B, S, self.dims.vocab)

    def text_embeds(self, input_ids: torch.Tensor) -> torch.Tensor:
        """fp32 embedding lookup [B*S, dim] for a batch without audio (the <audio> scatter has nothing to place)."""
        B, S = input_ids.shape
        ids = input_ids.to(device=self.device, dtype=torch.int64).contiguous()
        zero = torch.zeros(B, dtype=torch.int64, device=self.device)
        dummy = self.ws.typed("no_audio", (1, self.dims.lm_dim), F32)
        emb, _ = self.embed_scatter(ids, zero, dummy, 1)
        return emb

    @torch.no_grad()
    def audio_embeds(self, *, waveform=None, input_features=None, proj_params=None):
        B = (waveform if waveform is not None else input_features).shape[0]
        if waveform is not None:
            im2, _, T = self.logmel(waveform)
        else:
            im2, T = self.mel_to_im2col(input_features)
        enc = self.encode(im2, B, T)
        xs, n_a = self.frame_stack(enc)
        audio, _ = self.projector_forward(xs, proj_params, False)
        return audio, n_a

    @torch.no_grad()
    def greedy_generate(self, *, input_ids: torch.Tensor, proj_params=None, waveform=None, input_features=None, audio_embeds=None,
                        audio_token_counts=None, attention_mask=None, max_new_tokens: int = 16, eos_token_ids=(), pad_token_id: int = 0,
                        use_cache: bool = True, sync_every: int = 8, use_graph: bool = False):
        """Greedy decoding (num_beams=1, do_sample=False: the reference's generation defaults, asr_config.py:103-111).
        use_cache=True (default, like HF generate): one prefill pass that also fills the KV cache, then one
        ta_lm_decode_step per new token (HBM-bound skinny kernels, csrc/decode.cu).  use_cache=False re-runs the decoder
        over the whole sequence for every new token (kept as the A/B reference of the cache path).  use_graph: the decode
        step (~200 PDL launches, position counter on the device) is captured once per buffer set in a CUDA graph and replayed
        (-5 % per token; off by default because the capture costs more than it saves on short transcripts).

        Ragged prompts (asr_modeling.py:587-640 -> HF generate with attention_mask): `attention_mask` [B, S0] marks LEFT-padded
        prompts; rotary positions are counted from each sequence's first real token (HF: cumsum(mask) - 1), padding keys are never
        attended to, and per-sample `audio_token_counts` place each clip's own number of audio embeddings.  Batches larger than
        32 sequences (the decode kernels' row limit) are decoded in chunks of 32."""
        d = self.dims
        ids = input_ids.to(device=self.device, dtype=torch.int64).contiguous()
        B = ids.shape[0]
        if audio_embeds is not None:      # any projector module's output, fp32 [B, n_a, lm_dim] (QFormer: asr_modeling.generate)
            n_a = int(audio_embeds.shape[1])
            audio = audio_embeds.detach().to(device=self.device, dtype=F32).contiguous().view(B * n_a, d.lm_dim)
        else:
            audio, n_a = self.audio_embeds(waveform=waveform, input_features=input_features, proj_params=proj_params)
            audio = audio.clone()
        if audio_token_counts is None:
            audio_token_counts = (ids == d.audio_token_id).sum(-1)
        counts = audio_token_counts.to(device=self.device, dtype=torch.int64).contiguous()
        kv_start = pos_ids = None
        if attention_mask is not None:
            am = attention_mask.to("cpu", torch.int64)
            if tuple(am.shape) != tuple(ids.shape):
                raise L.TinyAudioB200Error(f"attention_mask {tuple(am.shape)} does not match input_ids {tuple(ids.shape)}")
            if bool((am[:, 1:] < am[:, :-1]).any()):
                raise L.TinyAudioB200Error("generate(): prompts must be LEFT-padded (attention_mask 0...01...1), as HF generate requires "
                                           "for decoder-only models; right-padded prompts would continue after the padding")
            if bool((am == 0).any()):
                kv_start = (am == 0).sum(-1).to(torch.int32).to(self.device)
                pos_ids = (am.cumsum(-1) - 1).clamp_(min=0).to(torch.int32).to(self.device).contiguous()
        if B > 32 and use_cache and max_new_tokens > 0:      # the skinny decode kernels hold <= 32 rows: decode in chunks
            outs = []
            for lo in range(0, B, 32):
                sl = slice(lo, min(B, lo + 32))
                outs.append(self.greedy_generate(input_ids=ids[sl], audio_embeds=audio.view(B, n_a, d.lm_dim)[sl], audio_token_counts=counts[sl],
                                                 attention_mask=(attention_mask[sl] if attention_mask is not None else None),
                                                 max_new_tokens=max_new_tokens, eos_token_ids=eos_token_ids, pad_token_id=pad_token_id,
                                                 use_cache=True, sync_every=sync_every, use_graph=use_graph))
            T = max(o.shape[1] for o in outs)      # a chunk that finished early is padded: same as running on after every eos
            return torch.cat([torch.nn.functional.pad(o, (0, T - o.shape[1]), value=pad_token_id) for o in outs], 0)
        eos = torch.tensor(list(eos_token_ids), device=self.device, dtype=torch.int64)
        done = torch.zeros(B, dtype=torch.bool, device=self.device)
        out = []
        if use_cache and max_new_tokens > 0:
            S0 = ids.shape[1]
            if S0 + max_new_tokens > d.lm_max_pos:
                raise L.TinyAudioB200Error(f"prompt {S0} + {max_new_tokens} new tokens exceed the rotary table ({d.lm_max_pos})")
            max_seq = _round_up(S0 + max_new_tokens, 64)
            kv_shape = (d.lm_layers, B, max_seq, d.lm_kv_heads * d.lm_head_dim)
            cache = (self.ws.typed("kv_cache_k", kv_shape, BF16), self.ws.typed("kv_cache_v", kv_shape, BF16), max_seq)
            emb, _ = self.embed_scatter(ids, counts, audio, n_a)
            hid = self.lm_hidden(emb, B, S0, kv_cache=cache, position_ids=pos_ids, kv_start=kv_start)
            last = torch.arange(B, device=self.device, dtype=torch.int32) * S0 + (S0 - 1)
            nxt = self.logits_rows(hid, last).float().argmax(-1)
            pos_dev = self.ws.typed("decode_pos", (1,), torch.int32)
            pos_dev.fill_(S0)
            logits = self.ws.typed("decode_logits", (B, self.lm.vocab_pad), BF16)
            fed = self.ws.typed("decode_ids_in", (B,), torch.int64)
            nxt_buf = self.ws.typed("decode_ids_out", (B,), torch.int64)
            step = self._decode_graph(fed, nxt_buf, pos_dev, cache, logits, S0, kv_start) if (use_graph and max_new_tokens >= 8) else None
            flags = []                                                      # per step: have all sequences finished?
            for t in range(max_new_tokens):
                nxt = torch.where(done, torch.full_like(nxt, pad_token_id), nxt)
                out.append(nxt)
                if eos.numel():
                    done = done | (nxt[:, None] == eos[None, :]).any(-1)
                    flags.append(done.all())
                    if (t + 1) % sync_every == 0 and bool(flags[-1]):       # host sync only every few tokens
                        break
                if t + 1 == max_new_tokens:
                    break
                fed.copy_(nxt)
                if step is not None:
                    step.replay()
                else:
                    self.decode_step(fed, pos_dev, S0 + t, cache, logits, nxt_buf, kv_start)
                nxt = nxt_buf.clone()
            res = torch.stack(out, dim=1)
            if flags:      # stop exactly where the token-by-token loop would have: the first step after which all are done
                f = torch.stack(flags).nonzero()
                if f.numel():
                    res = res[:, : int(f[0]) + 1]
            return res
        for _ in range(max_new_tokens):
            S = ids.shape[1]
            emb, _ = self.embed_scatter(ids, counts, audio, n_a)
            hid = self.lm_hidden(emb, B, S, position_ids=pos_ids, kv_start=kv_start)
            last = torch.arange(B, device=self.device, dtype=torch.int32) * S + (S - 1)
            nxt = self.logits_rows(hid, last).float().argmax(-1)
            nxt = torch.where(done, torch.full_like(nxt, pad_token_id), nxt)
            out.append(nxt)
            if eos.numel():
                done = done | (nxt[:, None] == eos[None, :]).any(-1)
            ids = torch.cat([ids, nxt[:, None]], dim=1).contiguous()
            if pos_ids is not None:
                pos_ids = torch.cat([pos_ids, pos_ids[:, -1:] + 1], dim=1).contiguous()
            if bool(done.all()):
                break
        return torch.stack(out, dim=1)

    def lm_loss_and_audio_grad(self, *, input_ids=None, audio=None, n_a=0, inputs_embeds=None, labels=None, labels_cpu=None,
                               audio_token_counts=None, num_items_in_batch=None, with_backward=True, lm_backward=None, train_lm=False,
                               want_hidden=False):
        """<audio> scatter -> Qwen3 -> CE (-> backward to the audio embeddings).  `audio` fp32 [B*n_a, lm_dim], or None for a batch
        without audio (plain embedding lookup); `inputs_embeds` fp32 [B, S, lm_dim] bypasses the lookup altogether.
        Returns (loss [1], d_audio [B*n_a, lm_dim] or None).  Projector-agnostic: any module that produces the audio embeddings
        can sit in front of it (ASRModel uses it for every projector except the fused MLP path).  lm_backward / train_lm: the
        decoder's own trainable tensors (LoRA A / B, or all Qwen3 weights) get their gradients in the same pass."""
        d = self.dims
        src = None
        if inputs_embeds is not None:
            B, S = int(inputs_embeds.shape[0]), int(inputs_embeds.shape[1])
            emb = inputs_embeds.detach().to(device=self.device, dtype=F32).contiguous().view(B * S, d.lm_dim)
            # no token ids: the embedding table receives no input-side gradient (every position counts as a placeholder)
            ids = (input_ids.to(device=self.device, dtype=torch.int64).contiguous() if input_ids is not None
                   else torch.full((B, S), d.audio_token_id, dtype=torch.int64, device=self.device))
        else:
            B, S = input_ids.shape
            ids = input_ids.to(device=self.device, dtype=torch.int64).contiguous()
            if audio is None:
                emb = self.text_embeds(ids)
            else:
                if audio_token_counts is None:
                    audio_token_counts = (ids == d.audio_token_id).sum(-1)
                counts = audio_token_counts.to(device=self.device, dtype=torch.int64).contiguous()
                emb, src = self.embed_scatter(ids, counts, audio, n_a)
        rows_d, tg_d, n_lab = self.label_rows(labels if labels is not None else labels_cpu)
        n_items = float(n_lab) if num_items_in_batch is None else float(num_items_in_batch)
        want_audio_grad = bool(with_backward and src is not None)
        lm_bwd = bool(want_audio_grad or lm_backward or train_lm)   # LoRA / unfrozen decoder need the LM backward even for a frozen projector
        loss, demb, _ = self.lm_step(emb, B, S, rows_d, tg_d, 1.0 / max(n_items, 1.0), lm_bwd, train_lm=train_lm, input_ids=ids,
                                     want_hidden=want_hidden)
        d_audio = None
        if want_audio_grad:
            d_audio = self.ws.typed("d_audio", (B * n_a, d.lm_dim), F32)
            d_audio.zero_()
            L.check(self.lib.ta_audio_grad_gather(L.ptr(src), L.ptr(demb), L.ptr(d_audio), B * S, d.lm_dim, L.stream_ptr()))
        return loss, d_audio

    @torch.no_grad()
    def encode_audio(self, *, waveform=None, input_features=None, frame_keep_prob=None, frame_keep_mask=None):
        """log-mel (if needed) + frozen encoder (+ audio-token dropout) -> bf16 [B, S_e, enc_dim]."""
        B = (waveform if waveform is not None else input_features).shape[0]
        if waveform is not None:
            im2, _, T = self.logmel(waveform)
        else:
            im2, T = self.mel_to_im2col(input_features)
        return self.apply_frame_dropout(self.encode(im2, B, T), frame_keep_prob, frame_keep_mask)

    # ------------------------------------------------------------------ the whole step
    def forward_backward(self, *, input_ids: torch.Tensor, proj_params, labels: Optional[torch.Tensor] = None,
                         labels_cpu: Optional[torch.Tensor] = None,
                         waveform: Optional[torch.Tensor] = None, input_features: Optional[torch.Tensor] = None,
                         audio_token_counts: Optional[torch.Tensor] = None, num_items_in_batch: Optional[float] = None,
                         grads: Optional[Dict[str, torch.Tensor]] = None, return_parts: bool = False,
                         frame_keep_prob: Optional[float] = None, frame_keep_mask: Optional[torch.Tensor] = None,
                         lm_backward: Optional[bool] = None, train_lm: bool = False, want_hidden: bool = False):
        """Returns (loss [1] fp32 device tensor, parts).  When `grads` is given (fp32 tensors shaped like the projector
        params) the backward runs and fills them with d(loss)/d(param).  `labels` may live on the host or on the device
        (`labels_cpu` is the older name of the same argument)."""
        if labels is None:
            labels = labels_cpu
        d = self.dims
        B, S = input_ids.shape
        parts = {}
        if waveform is not None:
            im2, mel_f32, T = self.logmel(waveform, want_f32=return_parts)
            if return_parts:
                parts["mel"] = mel_f32
        else:
            im2, T = self.mel_to_im2col(input_features)
        enc = self.apply_frame_dropout(self.encode(im2, B, T), frame_keep_prob, frame_keep_mask)
        xs, n_a = self.frame_stack(enc)
        with_bwd = grads is not None                           # projector gradients wanted
        lm_bwd = with_bwd if lm_backward is None else (lm_backward or with_bwd)   # LoRA-only training still needs the LM backward
        lm_bwd = lm_bwd or train_lm                                               # unfrozen LM: weight gradients -> self.lm.wgrad
        audio, stash = self.projector_forward(xs, proj_params, with_bwd)
        if audio_token_counts is None:
            audio_token_counts = (input_ids == d.audio_token_id).sum(-1)
        counts = audio_token_counts.to(device=self.device, dtype=torch.int64).contiguous()
        ids = input_ids.to(device=self.device, dtype=torch.int64).contiguous()
        emb, src = self.embed_scatter(ids, counts, audio, n_a)
        rows_d, tg_d, n_lab = self.label_rows(labels)
        n_items = float(n_lab) if num_items_in_batch is None else float(num_items_in_batch)
        inv = 1.0 / max(n_items, 1.0)
        loss, demb, _ = self.lm_step(emb, B, S, rows_d, tg_d, inv, lm_bwd, train_lm=train_lm, input_ids=ids, want_hidden=want_hidden)
        if with_bwd:
            d_audio = self.ws.typed("d_audio", (B * n_a, d.lm_dim), F32)
            d_audio.zero_()
            L.check(self.lib.ta_audio_grad_gather(L.ptr(src), L.ptr(demb), L.ptr(d_audio), B * S, d.lm_dim, L.stream_ptr()))
            self.projector_backward(stash, d_audio, grads)
        if return_parts:
            parts.update(encoder_out=enc, projector_out=audio.view(B, n_a, d.lm_dim), inputs_embeds=emb.view(B, S, d.lm_dim))
        return loss, parts


class FusedClipAdamW:
    """clip_grad_norm_(max_norm) + AdamW over a list of fp32 tensors, all on device, no host sync
    (replaces HF Trainer's clip + torch.optim.AdamW(fused=True); configs/training/production.yaml:5-9)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=1.0,
                 no_decay=()):
        self.lib = L.load()
        self.params = list(params)
        self.lr, self.betas, self.eps, self.wd, self.max_norm = lr, betas, eps, weight_decay, max_grad_norm
        self.no_decay = set(id(p) for p in no_decay)
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]
        self.step_count = 0
        self.gnorm_sq = torch.zeros(1, device=self.params[0].device, dtype=F32)

    def step(self, grads, lr: Optional[float] = None):
        st = L.stream_ptr()
        self.step_count += 1
        self.gnorm_sq.zero_()
        for g in grads:
            L.check(self.lib.ta_grad_sumsq(L.ptr(g), g.numel(), L.ptr(self.gnorm_sq), st))
        lr = self.lr if lr is None else lr
        for p, g, m, v in zip(self.params, grads, self.m, self.v):
            wd = 0.0 if id(p) in self.no_decay else self.wd
            L.check(self.lib.ta_adamw_clip_step(L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), p.numel(), lr, self.betas[0],
                                                self.betas[1], self.eps, wd, self.step_count, self.max_norm,
                                                L.ptr(self.gnorm_sq), st))

    def grad_norm(self) -> torch.Tensor:
        return self.gnorm_sq.sqrt()
