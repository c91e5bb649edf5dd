"""CPU oracle for the tiny-audio training hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a plain, functional (no nn.Module) fp32 restatement of the arithmetic on the
reference's training path:

    waveform -> log-mel -> GLM-ASR encoder -> MLP projector -> gather/scatter -> Qwen3 -> CE
             -> backward into the projector -> global-norm clip -> AdamW

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import it.  The product (``tiny_audio_b200``) never does and fails loudly without its
CUDA library.

Parity pin: ``oracle/make_golden.py`` loads the SAME seeded weights into the unmodified reference
(``/root/reference/tiny_audio`` + the ``transformers`` 5.5.0 modules it calls) and stores the
reference's outputs under ``tests/golden/``;  ``tests/test_oracle_golden.py`` checks this file
against those fixtures.  The reference's own tests hold no floating-point golden vectors
(SURVEY.md section 4), so the fixtures generated from the reference itself ARE the pin; the
integer/layout semantics are additionally pinned by the reference's unit-test cases restated in
``tests/test_host_logic.py``.

Citations: ``HF:`` = site-packages/transformers (5.5.0); other paths are under /root/reference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field, asdict
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# configuration of the path (dims of the two frozen towers + projector)
# --------------------------------------------------------------------------------------
@dataclass
class PathConfig:
    # log-mel (HF:models/whisper/feature_extraction_whisper.py:60-103)
    sample_rate: int = 16000
    n_fft: int = 400
    hop: int = 160
    n_mels: int = 128
    # GLM-ASR encoder (HF:models/glmasr/configuration_glmasr.py:44-61)
    enc_dim: int = 1280
    enc_ffn: int = 5120
    enc_layers: int = 32
    enc_heads: int = 20            # head_dim = enc_dim // enc_heads = 64
    enc_rope_theta: float = 10000.0
    enc_partial_rotary: float = 0.5
    enc_ln_eps: float = 1e-5
    # projector (tiny_audio/projectors.py:23-71)
    proj_k: int = 4
    proj_hidden: int = 1024
    proj_eps: float = 1e-6
    # Qwen3-0.6B (public checkpoint config; SURVEY.md section 8 header)
    lm_dim: int = 1024
    lm_ffn: int = 3072
    lm_layers: int = 28
    lm_heads: int = 16
    lm_kv_heads: int = 8
    lm_head_dim: int = 128
    lm_rope_theta: float = 1e6
    lm_eps: float = 1e-6
    vocab: int = 151936
    audio_token_id: int = 151669

    @property
    def enc_head_dim(self) -> int:
        return self.enc_dim // self.enc_heads

    def to_dict(self):
        return asdict(self)


FULL = PathConfig()


def small_config(**kw) -> PathConfig:
    """Reduced-depth / reduced-vocab variant used by the fast tests (widths stay full size so
    that every kernel sees its production tile shapes)."""
    base = dict(enc_layers=2, lm_layers=2, vocab=5003, audio_token_id=5002)
    base.update(kw)
    return PathConfig(**base)


# --------------------------------------------------------------------------------------
# seeded weights (names = the HF state_dict names so they load into the reference 1:1)
# --------------------------------------------------------------------------------------
def init_weights(cfg: PathConfig, seed: int = 1234, emb_std: float = 0.02) -> Dict[str, Dict[str, Tensor]]:
    """Deterministic fp32 weights from torch's CPU generator.

    Unlike HF's default init, biases and norm gains/offsets are non-trivial so that a kernel which
    drops a bias or a gain cannot pass parity.  Linear weights use 1/sqrt(fan_in) so activations
    stay O(1) through 32 + 28 layers.
    """
    g = torch.Generator(device="cpu").manual_seed(seed)

    def lin(o, i, scale=1.0):
        return torch.randn(o, i, generator=g) * (scale / math.sqrt(i))

    def vec(n, mean=0.0, std=0.02):
        return mean + std * torch.randn(n, generator=g)

    enc: Dict[str, Tensor] = {}
    D, Fe = cfg.enc_dim, cfg.enc_ffn
    enc["conv1.weight"] = torch.randn(D, cfg.n_mels, 3, generator=g) / math.sqrt(3 * cfg.n_mels)
    enc["conv1.bias"] = vec(D)
    enc["conv2.weight"] = torch.randn(D, D, 3, generator=g) / math.sqrt(3 * D)
    enc["conv2.bias"] = vec(D)
    for i in range(cfg.enc_layers):
        p = f"layers.{i}."
        enc[p + "input_layernorm.weight"] = vec(D, 1.0, 0.1)
        enc[p + "input_layernorm.bias"] = vec(D)
        enc[p + "self_attn.q_proj.weight"] = lin(D, D)
        enc[p + "self_attn.q_proj.bias"] = vec(D)
        enc[p + "self_attn.k_proj.weight"] = lin(D, D)           # no bias (HF:glmasr:188)
        enc[p + "self_attn.v_proj.weight"] = lin(D, D)
        enc[p + "self_attn.v_proj.bias"] = vec(D)
        enc[p + "self_attn.o_proj.weight"] = lin(D, D, 0.5)
        enc[p + "self_attn.o_proj.bias"] = vec(D)
        enc[p + "post_attention_layernorm.weight"] = vec(D, 1.0, 0.1)
        enc[p + "post_attention_layernorm.bias"] = vec(D)
        enc[p + "mlp.fc1.weight"] = lin(Fe, D)
        enc[p + "mlp.fc1.bias"] = vec(Fe)
        enc[p + "mlp.fc2.weight"] = lin(D, Fe, 0.5)
        enc[p + "mlp.fc2.bias"] = vec(D)
    enc["norm.weight"] = vec(D, 1.0, 0.1)
    enc["norm.bias"] = vec(D)

    proj: Dict[str, Tensor] = {}
    H = cfg.proj_hidden
    proj["linear_1.weight"] = lin(H, cfg.proj_k * D)
    proj["norm.weight"] = vec(H, 1.0, 0.1)
    proj["linear_2.weight"] = lin(cfg.lm_dim, H)
    proj["norm_2.weight"] = vec(cfg.lm_dim, 1.0, 0.1)

    lm: Dict[str, Tensor] = {}
    Dl, Fl, hd = cfg.lm_dim, cfg.lm_ffn, cfg.lm_head_dim
    lm["model.embed_tokens.weight"] = torch.randn(cfg.vocab, Dl, generator=g) * emb_std
    for i in range(cfg.lm_layers):
        p = f"model.layers.{i}."
        lm[p + "input_layernorm.weight"] = vec(Dl, 1.0, 0.1)
        lm[p + "self_attn.q_proj.weight"] = lin(cfg.lm_heads * hd, Dl)
        lm[p + "self_attn.k_proj.weight"] = lin(cfg.lm_kv_heads * hd, Dl)
        lm[p + "self_attn.v_proj.weight"] = lin(cfg.lm_kv_heads * hd, Dl)
        lm[p + "self_attn.o_proj.weight"] = lin(Dl, cfg.lm_heads * hd, 0.5)
        lm[p + "self_attn.q_norm.weight"] = vec(hd, 1.0, 0.1)
        lm[p + "self_attn.k_norm.weight"] = vec(hd, 1.0, 0.1)
        lm[p + "post_attention_layernorm.weight"] = vec(Dl, 1.0, 0.1)
        lm[p + "mlp.gate_proj.weight"] = lin(Fl, Dl)
        lm[p + "mlp.up_proj.weight"] = lin(Fl, Dl)
        lm[p + "mlp.down_proj.weight"] = lin(Dl, Fl, 0.5)
    lm["model.norm.weight"] = vec(Dl, 1.0, 0.1)
    # tied lm_head (Qwen3-0.6B tie_word_embeddings=True): same storage
    lm["lm_head.weight"] = lm["model.embed_tokens.weight"]
    return {"encoder": enc, "projector": proj, "lm": lm}


# --------------------------------------------------------------------------------------
# a1. log-mel   (HF:models/whisper/feature_extraction_whisper.py:135-164, 95-103, 328-337)
# --------------------------------------------------------------------------------------
def _hz_to_mel_slaney(f: np.ndarray) -> np.ndarray:
    # HF:audio_utils.py:285-296 (slaney scale: linear below 1 kHz, log above)
    f = np.asarray(f, dtype=np.float64)
    mels = 3.0 * f / 200.0
    logstep = 27.0 / np.log(6.4)
    hi = f >= 1000.0
    mels = np.where(hi, 15.0 + np.log(np.maximum(f, 1e-30) / 1000.0) * logstep, mels)
    return mels


def _mel_to_hz_slaney(m: np.ndarray) -> np.ndarray:
    m = np.asarray(m, dtype=np.float64)
    f = 200.0 * m / 3.0
    logstep = np.log(6.4) / 27.0
    hi = m >= 15.0
    return np.where(hi, 1000.0 * np.exp(logstep * (m - 15.0)), f)


def mel_filter_bank(n_freq: int = 201, n_mels: int = 128, fmin: float = 0.0, fmax: float = 8000.0,
                    sr: int = 16000) -> np.ndarray:
    """Slaney-scale, slaney-normalised triangular bank, float64 (n_freq, n_mels).
    Follows HF:audio_utils.py:515-535 + _create_triangular_filter_bank."""
    mel_pts = np.linspace(_hz_to_mel_slaney(fmin), _hz_to_mel_slaney(fmax), n_mels + 2)
    hz_pts = _mel_to_hz_slaney(mel_pts)
    fft_freqs = np.linspace(0, sr // 2, n_freq)
    fdiff = np.diff(hz_pts)
    slopes = hz_pts[None, :] - fft_freqs[:, None]
    down = -slopes[:, :-2] / fdiff[:-1]
    up = slopes[:, 2:] / fdiff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    enorm = 2.0 / (hz_pts[2:n_mels + 2] - hz_pts[:n_mels])
    return fb * enorm[None, :]


def log_mel(wave: Tensor, cfg: PathConfig = FULL) -> Tensor:
    """wave (B, L) float32 (already zero-padded) -> (B, n_mels, L // hop) float32.

    reflect-pad n_fft/2, periodic Hann, |rfft|^2, drop the last frame, mel, log10(clamp 1e-10),
    per-clip floor at max-8, (x+4)/4   (HF:whisper/fe:135-164)."""
    assert wave.dim() == 2
    B, L = wave.shape
    n_fft, hop = cfg.n_fft, cfg.hop
    x = F.pad(wave.float().unsqueeze(1), (n_fft // 2, n_fft // 2), mode="reflect").squeeze(1)
    frames = x.unfold(-1, n_fft, hop)                       # (B, 1 + L//hop, n_fft)
    win = torch.hann_window(n_fft, periodic=True, dtype=torch.float32)
    spec = torch.fft.rfft(frames * win, dim=-1)             # (B, T+1, 201)
    power = (spec.real ** 2 + spec.imag ** 2)[:, :-1, :]    # drop last frame
    fb = torch.from_numpy(mel_filter_bank(n_fft // 2 + 1, cfg.n_mels, 0.0, 8000.0, cfg.sample_rate)).float()
    mel = torch.einsum("fm,btf->bmt", fb, power)
    logm = torch.clamp(mel, min=1e-10).log10()
    mx = logm.amax(dim=(1, 2), keepdim=True)
    logm = torch.maximum(logm, mx - 8.0)
    return (logm + 4.0) / 4.0


def mel_attention_mask(sample_lengths: Tensor, padded_len: int, hop: int = 160) -> Tensor:
    """sample mask[:, ::hop], minus the last column when padded_len % hop != 0 (HF:whisper/fe:328-337)."""
    idx = torch.arange(0, padded_len, hop)
    m = (idx[None, :] < sample_lengths[:, None]).to(torch.int32)
    if padded_len % hop != 0:
        m = m[:, :-1]
    return m


# --------------------------------------------------------------------------------------
# a2. token-count arithmetic (tiny_audio/asr_config.py:9-19, projectors.py:52-55) -- integers
# --------------------------------------------------------------------------------------
DEFAULT_CONV_LAYERS = [(1, 3, 1), (1, 3, 2)]


def encoder_output_length(mel_len, conv_layers=None):
    n = mel_len
    for pad, k, s in (conv_layers or DEFAULT_CONV_LAYERS):
        n = (n + 2 * pad - (k - 1) - 1) // s + 1
    return n


def projector_output_length(enc_len, k: int = 4):
    return (enc_len - k) // k + 1


# --------------------------------------------------------------------------------------
# a3. GLM-ASR encoder forward (HF:models/glmasr/modeling_glmasr.py:316-330, 253-274, 192-225)
# --------------------------------------------------------------------------------------
def _rope_tables(seq: int, dim: int, theta: float) -> Tuple[Tensor, Tensor]:
    inv = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.int64).float() / dim))
    fr = torch.arange(seq).float()[:, None] * inv[None, :]
    emb = torch.cat([fr, fr], dim=-1)
    return emb.cos(), emb.sin()


def _rotate_half(x: Tensor) -> Tensor:
    h = x.shape[-1] // 2
    return torch.cat([-x[..., h:], x[..., :h]], dim=-1)


def encoder_forward(w: Dict[str, Tensor], mel: Tensor, cfg: PathConfig = FULL) -> Tensor:
    """mel (B, n_mels, T) -> (B, S_e, enc_dim).  No attention mask (HF:glmasr:217)."""
    x = F.gelu(F.conv1d(mel, w["conv1.weight"], w["conv1.bias"], padding=1))
    x = F.gelu(F.conv1d(x, w["conv2.weight"], w["conv2.bias"], stride=2, padding=1))
    x = x.transpose(1, 2)                                   # (B, S, D)
    B, S, D = x.shape
    H, hd = cfg.enc_heads, cfg.enc_head_dim
    rd = int(hd * cfg.enc_partial_rotary)
    cos, sin = _rope_tables(S, rd, cfg.enc_rope_theta)      # (S, rd)
    for i in range(cfg.enc_layers):
        p = f"layers.{i}."
        h = F.layer_norm(x, (D,), w[p + "input_layernorm.weight"], w[p + "input_layernorm.bias"], cfg.enc_ln_eps)
        q = F.linear(h, w[p + "self_attn.q_proj.weight"], w[p + "self_attn.q_proj.bias"]).view(B, S, H, hd).transpose(1, 2)
        k = F.linear(h, w[p + "self_attn.k_proj.weight"]).view(B, S, H, hd).transpose(1, 2)
        v = F.linear(h, w[p + "self_attn.v_proj.weight"], w[p + "self_attn.v_proj.bias"]).view(B, S, H, hd).transpose(1, 2)
        qr, qp = q[..., :rd], q[..., rd:]
        kr, kp = k[..., :rd], k[..., rd:]
        q = torch.cat([qr * cos + _rotate_half(qr) * sin, qp], dim=-1)
        k = torch.cat([kr * cos + _rotate_half(kr) * sin, kp], dim=-1)
        att = torch.softmax((q @ k.transpose(-1, -2)) * (hd ** -0.5), dim=-1) @ v
        att = att.transpose(1, 2).reshape(B, S, D)
        x = x + F.linear(att, w[p + "self_attn.o_proj.weight"], w[p + "self_attn.o_proj.bias"])
        h = F.layer_norm(x, (D,), w[p + "post_attention_layernorm.weight"], w[p + "post_attention_layernorm.bias"], cfg.enc_ln_eps)
        h = F.gelu(F.linear(h, w[p + "mlp.fc1.weight"], w[p + "mlp.fc1.bias"]))
        x = x + F.linear(h, w[p + "mlp.fc2.weight"], w[p + "mlp.fc2.bias"])
    return F.layer_norm(x, (D,), w["norm.weight"], w["norm.bias"], cfg.enc_ln_eps)


# --------------------------------------------------------------------------------------
# a5. frame-stack + MLP projector (tiny_audio/projectors.py:57-71, 79-87)
# --------------------------------------------------------------------------------------
def rms_norm(x: Tensor, weight: Tensor, eps: float) -> Tensor:
    # LlamaRMSNorm / Qwen3RMSNorm (HF:qwen3:59-64): fp32 variance, gain applied after
    v = x.float().pow(2).mean(-1, keepdim=True)
    return weight * (x.float() * torch.rsqrt(v + eps)).to(x.dtype)


def frame_stack(x: Tensor, k: int) -> Tensor:
    B, S, D = x.shape
    n = (S - k) // k + 1
    return x[:, : n * k, :].reshape(B, n, D * k)


def frame_stack_indices(S: int, k: int, D: int) -> np.ndarray:
    """For output row j, column c: source (frame, feature) = (k*j + c // D, c % D).  Returns the
    (n, k*D, 2) int64 table -- the 'frame-stack indices' that must be bit-exact."""
    n = (S - k) // k + 1
    j = np.arange(n)[:, None]
    c = np.arange(k * D)[None, :]
    return np.stack([k * j + c // D, np.broadcast_to(c % D, (n, k * D))], axis=-1).astype(np.int64)


def projector_forward(w: Dict[str, Tensor], enc_out: Tensor, cfg: PathConfig = FULL) -> Tensor:
    x = frame_stack(enc_out, cfg.proj_k)
    x = F.linear(x, w["linear_1.weight"])
    x = rms_norm(x, w["norm.weight"], cfg.proj_eps)
    x = F.gelu(x)
    x = F.linear(x, w["linear_2.weight"])
    return rms_norm(x, w["norm_2.weight"], cfg.proj_eps)


# --------------------------------------------------------------------------------------
# a6. QFormer projector (tiny_audio/projectors.py:359-475; HF:models/blip_2/modeling_blip_2.py:537-1042), eval mode (no dropout)
# --------------------------------------------------------------------------------------
QF_WINDOW, QF_DOWNSAMPLE, QF_HEADS, QF_LAYERS, QF_EPS = 15, 5, 16, 2, 1e-12


def init_qformer_weights(cfg: PathConfig, seed: int = 77) -> Dict[str, Tensor]:
    """Seeded weights under the reference's parameter names (QFormerAudioProjector.state_dict())."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    H, Fq = cfg.enc_dim, 4 * cfg.enc_dim

    def lin(o, i):
        return torch.randn(o, i, generator=g) / math.sqrt(i)

    def vec(n, mean=0.0, std=0.05):
        return mean + std * torch.randn(n, generator=g)

    w = {"query": torch.randn(1, QF_WINDOW // QF_DOWNSAMPLE, H, generator=g),
         "qformer.layernorm.weight": vec(H, 1.0, 0.1), "qformer.layernorm.bias": vec(H)}
    for i in range(QF_LAYERS):
        for att in ("attention", "crossattention"):
            p = f"qformer.encoder.layer.{i}.{att}."
            for n in ("query", "key", "value"):
                w[p + f"attention.{n}.weight"] = lin(H, H)
                w[p + f"attention.{n}.bias"] = vec(H)
            w[p + "output.dense.weight"] = lin(H, H)
            w[p + "output.dense.bias"] = vec(H)
            w[p + "output.LayerNorm.weight"] = vec(H, 1.0, 0.1)
            w[p + "output.LayerNorm.bias"] = vec(H)
        p = f"qformer.encoder.layer.{i}."
        w[p + "intermediate_query.dense.weight"] = lin(Fq, H)
        w[p + "intermediate_query.dense.bias"] = vec(Fq)
        w[p + "output_query.dense.weight"] = lin(H, Fq)
        w[p + "output_query.dense.bias"] = vec(H)
        w[p + "output_query.LayerNorm.weight"] = vec(H, 1.0, 0.1)
        w[p + "output_query.LayerNorm.bias"] = vec(H)
    w["linear.weight"] = lin(cfg.lm_dim, H)
    w["linear.bias"] = vec(cfg.lm_dim)
    return w


def qformer_output_length(enc_len):
    return ((enc_len + QF_WINDOW - 1) // QF_WINDOW) * (QF_WINDOW // QF_DOWNSAMPLE)


def qformer_projector_forward(w: Dict[str, Tensor], enc_out: Tensor, cfg: PathConfig = FULL, drop_masks=None) -> Tensor:
    """`drop_masks`: optional iterable of hidden-dropout multiplier tensors (0 or 1/(1-p), [windows * queries, H]) in the order the
    reference draws them -- after the query LayerNorm (HF:models/blip_2/modeling_blip_2.py:985-986), then per layer after the
    self-attention, cross-attention and FFN output projections (:644-648, :700-704); None = eval mode.  (Attention-probability
    dropout, :622, is not injectable here.)"""
    masks = iter(drop_masks) if drop_masks is not None else None

    def drop(t):
        return t if masks is None else t * next(masks).reshape(t.shape).to(t.dtype)
    B, S, H = enc_out.shape
    nb = -(-S // QF_WINDOW)
    x_enc = F.pad(enc_out.float(), (0, 0, 0, nb * QF_WINDOW - S)).reshape(B * nb, QF_WINDOW, H)
    hd = H // QF_HEADS

    def attend(p, x, src):
        Wn = x.shape[0]
        q = F.linear(x, w[p + "attention.query.weight"], w[p + "attention.query.bias"]).view(Wn, -1, QF_HEADS, hd).transpose(1, 2)
        k = F.linear(src, w[p + "attention.key.weight"], w[p + "attention.key.bias"]).view(Wn, -1, QF_HEADS, hd).transpose(1, 2)
        v = F.linear(src, w[p + "attention.value.weight"], w[p + "attention.value.bias"]).view(Wn, -1, QF_HEADS, hd).transpose(1, 2)
        ctx = (torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(hd), -1) @ v).transpose(1, 2).reshape(Wn, -1, H)
        o = drop(F.linear(ctx, w[p + "output.dense.weight"], w[p + "output.dense.bias"]))
        return F.layer_norm(o + x, (H,), w[p + "output.LayerNorm.weight"], w[p + "output.LayerNorm.bias"], QF_EPS)

    x = drop(F.layer_norm(w["query"], (H,), w["qformer.layernorm.weight"], w["qformer.layernorm.bias"], QF_EPS).expand(B * nb, -1, -1))
    for i in range(QF_LAYERS):
        p = f"qformer.encoder.layer.{i}."
        x = attend(p + "attention.", x, x)
        x = attend(p + "crossattention.", x, x_enc)
        h = F.gelu(F.linear(x, w[p + "intermediate_query.dense.weight"], w[p + "intermediate_query.dense.bias"]))
        f = drop(F.linear(h, w[p + "output_query.dense.weight"], w[p + "output_query.dense.bias"]))
        x = F.layer_norm(f + x, (H,), w[p + "output_query.LayerNorm.weight"], w[p + "output_query.LayerNorm.bias"], QF_EPS)
    return F.linear(x.reshape(B, nb * (QF_WINDOW // QF_DOWNSAMPLE), H), w["linear.weight"], w["linear.bias"])


# --------------------------------------------------------------------------------------
# f4. MOSA projector (tiny_audio/projectors.py:103-177): conv downsampler x2 -> dense softmax mixture of 2-layer GELU adapters
# --------------------------------------------------------------------------------------
MOSA_ADAPTER_HIDDEN, MOSA_ROUTER_HIDDEN, MOSA_EXPERTS = 4096, 512, 4


def _adapter_weights(w, prefix, g, i_dim, h_dim, o_dim, fc2_std=None):
    """SimpleAdapter parameters (projectors.py:90-100) under the reference's names: fc1 / fc2 with biases."""
    w[prefix + "fc1.weight"] = torch.randn(h_dim, i_dim, generator=g) / math.sqrt(i_dim)
    w[prefix + "fc1.bias"] = 0.05 * torch.randn(h_dim, generator=g)
    w[prefix + "fc2.weight"] = torch.randn(o_dim, h_dim, generator=g) * (fc2_std if fc2_std else 1.0 / math.sqrt(h_dim))
    w[prefix + "fc2.bias"] = 0.05 * torch.randn(o_dim, generator=g)


def init_mosa_weights(cfg: PathConfig, seed: int = 78) -> Dict[str, Tensor]:
    """Seeded weights under MOSAProjector.state_dict() names (projectors.py:133-151)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    D, O = cfg.enc_dim, cfg.lm_dim
    w = {"downsampler.0.weight": torch.randn(D, D, 3, generator=g) / math.sqrt(3 * D),
         "downsampler.0.bias": 0.05 * torch.randn(D, generator=g),
         "downsampler.2.weight": torch.randn(O, D, 3, generator=g) / math.sqrt(3 * D),
         "downsampler.2.bias": 0.05 * torch.randn(O, generator=g),
         "router.0.weight": torch.randn(MOSA_ROUTER_HIDDEN, O, generator=g) / math.sqrt(O),
         "router.0.bias": 0.05 * torch.randn(MOSA_ROUTER_HIDDEN, generator=g),
         "router.2.weight": torch.randn(MOSA_EXPERTS, MOSA_ROUTER_HIDDEN, generator=g) * (2.0 / math.sqrt(MOSA_ROUTER_HIDDEN)),
         "router.2.bias": 0.05 * torch.randn(MOSA_EXPERTS, generator=g)}
    for i in range(MOSA_EXPERTS):
        _adapter_weights(w, f"experts.{i}.", g, O, MOSA_ADAPTER_HIDDEN, O)
    return w


def mosa_output_length(enc_len):
    """Two k=3, s=2, p=1 convolutions (projectors.py:172-177)."""
    for _ in range(2):
        enc_len = (enc_len + 2 * 1 - 3) // 2 + 1
    return enc_len


def mosa_projector_forward(w: Dict[str, Tensor], enc_out: Tensor, cfg: PathConfig = FULL) -> Tensor:
    """projectors.py:153-170: every expert sees every token; outputs are mixed with the router's softmax."""
    x = enc_out.float().transpose(1, 2)
    for idx in ("0", "2"):                      # nn.Sequential(Conv1d, GELU, Conv1d, GELU)
        x = F.gelu(F.conv1d(x, w[f"downsampler.{idx}.weight"], w[f"downsampler.{idx}.bias"], stride=2, padding=1))
    x = x.transpose(1, 2)
    hidden = F.relu(F.linear(x, w["router.0.weight"], w["router.0.bias"]))
    mix = torch.softmax(F.linear(hidden, w["router.2.weight"], w["router.2.bias"]), dim=-1)
    out = None
    for i in range(mix.shape[-1]):
        y = F.linear(F.gelu(F.linear(x, w[f"experts.{i}.fc1.weight"], w[f"experts.{i}.fc1.bias"])),
                     w[f"experts.{i}.fc2.weight"], w[f"experts.{i}.fc2.bias"]) * mix[..., i:i + 1]
        out = y if out is None else out + y
    return out


# --------------------------------------------------------------------------------------
# f4. shared + sparse MoE projector (tiny_audio/projectors.py:185-351): frame-stack -> RMSNorm -> shared adapter + top-k of E adapters
# --------------------------------------------------------------------------------------
MOE_EXPERTS, MOE_TOP_K, MOE_AUX_COEF, MOE_Z_COEF = 4, 2, 0.01, 1e-4


def init_moe_weights(cfg: PathConfig, seed: int = 79) -> Dict[str, Tensor]:
    """Seeded weights under MoEAudioProjector.state_dict() names (projectors.py:223-235); the scales follow its
    _init_weights (:242-251): router std 0.02, fc2 std 0.01."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    I, H, O = cfg.enc_dim * cfg.proj_k, cfg.proj_hidden, cfg.lm_dim
    w = {"norm.weight": 1.0 + 0.1 * torch.randn(I, generator=g),
         "router.weight": 0.02 * torch.randn(MOE_EXPERTS, I, generator=g)}
    for i in range(MOE_EXPERTS):
        _adapter_weights(w, f"experts.{i}.", g, I, H, O, fc2_std=0.01)
    _adapter_weights(w, "shared_expert.", g, I, H, O, fc2_std=0.01)
    return w


def moe_router(w: Dict[str, Tensor], flat: Tensor):
    """projectors.py:291-309 without jitter: fp32 softmax over the router logits, top-k, weights renormalised with +1e-6."""
    logits = F.linear(flat, w["router.weight"])
    probs = torch.softmax(logits.float(), dim=-1)
    top_w, top_i = torch.topk(probs, MOE_TOP_K, dim=-1)
    top_w = top_w / (top_w.sum(dim=-1, keepdim=True) + 1e-6)
    return logits, probs, top_w, top_i


def moe_projector_forward(w: Dict[str, Tensor], enc_out: Tensor, cfg: PathConfig = FULL, training: bool = True):
    """-> (audio embeddings, aux loss).  projectors.py:257-347 with router_jitter_noise = 0 (the jitter is RNG: no parity
    definition).  aux = load-balance + z-loss in training mode, 0 in eval mode (:311-325)."""
    x = rms_norm(frame_stack(enc_out.float(), cfg.proj_k), w["norm.weight"], cfg.proj_eps)
    B, n, I = x.shape
    flat = x.reshape(B * n, I)

    def adapter(prefix, t):
        return F.linear(F.gelu(F.linear(t, w[prefix + "fc1.weight"], w[prefix + "fc1.bias"])), w[prefix + "fc2.weight"], w[prefix + "fc2.bias"])

    out = adapter("shared_expert.", flat)
    logits, probs, top_w, top_i = moe_router(w, flat)
    aux = torch.zeros((), dtype=torch.float32)
    if training:
        n_e = probs.shape[-1]
        balance = MOE_AUX_COEF * ((probs.mean(0) - 1.0 / n_e) ** 2).mean() * n_e
        aux = balance + MOE_Z_COEF * torch.logsumexp(logits, dim=-1).pow(2).mean()
    for e in range(probs.shape[-1]):                          # sparse dispatch: only the tokens that picked expert e
        rows, slot = torch.nonzero(top_i == e, as_tuple=True)
        if rows.numel():
            out = out.index_add(0, rows, adapter(f"experts.{e}.", flat[rows]) * top_w[rows, slot].unsqueeze(-1))
    return out.view(B, n, -1), aux


def projector_kind(weights: Dict[str, Tensor]) -> str:
    if "query" in weights:
        return "qformer"
    if "downsampler.0.weight" in weights:
        return "mosa"
    if "router.weight" in weights:
        return "moe"
    return "mlp"


# --------------------------------------------------------------------------------------
# a7. ragged gather + masked_scatter (tiny_audio/asr_modeling.py:27-44, 497-515)
# --------------------------------------------------------------------------------------
def gather_audio_embeds(audio_embeds: Tensor, token_counts: Tensor) -> Tensor:
    B, n, D = audio_embeds.shape
    rows = []
    for i in range(B):
        c = int(token_counts[i])
        take = audio_embeds[i, : min(c, n)]
        if c > n:
            take = torch.cat([take, audio_embeds.new_zeros(c - n, D)], dim=0)
        rows.append(take)
    return torch.cat(rows, dim=0) if rows else audio_embeds.new_zeros(0, D)


def scatter_audio(inputs_embeds: Tensor, input_ids: Tensor, packed: Tensor, audio_token_id: int) -> Tensor:
    """masked_scatter semantics: the j-th <audio> position in row-major (b, s) order receives
    packed[j]."""
    out = inputs_embeds.clone()
    pos = (input_ids == audio_token_id).reshape(-1).nonzero().squeeze(-1)
    flat = out.view(-1, out.shape[-1])
    assert packed.shape[0] >= pos.numel(), "masked_scatter needs at least as many source rows"
    flat[pos] = packed[: pos.numel()].to(flat.dtype)
    return out


# --------------------------------------------------------------------------------------
# a8/a9. Qwen3 forward + CE (HF:models/qwen3/modeling_qwen3.py:378-517; HF:loss/loss_utils.py:28-67)
# --------------------------------------------------------------------------------------
def init_lora_weights(cfg: PathConfig, seed: int = 55, rank: int = 8, alpha: float = 32.0, b_std: float = 0.02):
    """peft-style adapters for q,k,v,o,gate,up,down of every decoder layer, stacked over layers: A [L, r, in] (kaiming-uniform
    like peft), B [L, out, r].  peft initialises B to zero; a small random B is used here so that every gradient path is live."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    D, Fl, hd = cfg.lm_dim, cfg.lm_ffn, cfg.lm_head_dim
    dims = {"q_proj": (D, cfg.lm_heads * hd), "k_proj": (D, cfg.lm_kv_heads * hd), "v_proj": (D, cfg.lm_kv_heads * hd),
            "o_proj": (cfg.lm_heads * hd, D), "gate_proj": (D, Fl), "up_proj": (D, Fl), "down_proj": (Fl, D)}
    A, Bm = {}, {}
    for k, (i, o) in dims.items():
        bound = 1.0 / math.sqrt(i)          # kaiming_uniform_(a=sqrt(5)) on [r, in]
        A[k] = (torch.rand(cfg.lm_layers, rank, i, generator=g) * 2 - 1) * bound
        Bm[k] = torch.randn(cfg.lm_layers, o, rank, generator=g) * b_std
    return {"A": A, "B": Bm, "scaling": alpha / rank}


def lm_forward(w: Dict[str, Tensor], inputs_embeds: Tensor, cfg: PathConfig = FULL,
               attention_mask: Optional[Tensor] = None, lora=None, position_ids: Optional[Tensor] = None) -> Tensor:
    """inputs_embeds (B, S, D) -> final-norm hidden states (B, S, D).  `lora`: adapters of init_lora_weights
    (peft semantics: y = W x + alpha/r * B(A(x)), dropout 0; tiny_audio/asr_modeling.py:289-301).
    `position_ids` (B, S): rotary positions (HF generate with a left-padded attention_mask: cumsum(mask) - 1); default arange(S).
    A query row whose keys are ALL masked (a left-padding row) keeps plain causal attention, so that it stays finite -- its output
    is never used (HF's sdpa path un-masks such rows for the same reason)."""
    def lin(x, name, layer, proj):
        y = F.linear(x, w[name])
        if lora is not None and proj in lora["A"]:
            y = y + lora["scaling"] * F.linear(F.linear(x, lora["A"][proj][layer]), lora["B"][proj][layer])
        return y

    x = inputs_embeds
    B, S, D = x.shape
    Hq, Hkv, hd = cfg.lm_heads, cfg.lm_kv_heads, cfg.lm_head_dim
    if position_ids is None:
        cos, sin = _rope_tables(S, hd, cfg.lm_rope_theta)
    else:
        tc, ts = _rope_tables(int(position_ids.max()) + 1, hd, cfg.lm_rope_theta)
        cos, sin = tc[position_ids][:, None], ts[position_ids][:, None]          # (B, 1, S, hd)
    causal = torch.ones(S, S, dtype=torch.bool).tril()
    mask = causal[None, None]
    if attention_mask is not None:
        mask = mask & attention_mask.bool()[:, None, None, :]
        dead = ~mask.any(-1, keepdim=True)                                        # padding query rows
        mask = mask | (dead & causal[None, None])
    for i in range(cfg.lm_layers):
        p = f"model.layers.{i}."
        h = rms_norm(x, w[p + "input_layernorm.weight"], cfg.lm_eps)
        q = lin(h, p + "self_attn.q_proj.weight", i, "q_proj").view(B, S, Hq, hd)
        k = lin(h, p + "self_attn.k_proj.weight", i, "k_proj").view(B, S, Hkv, hd)
        v = lin(h, p + "self_attn.v_proj.weight", i, "v_proj").view(B, S, Hkv, hd)
        q = rms_norm(q, w[p + "self_attn.q_norm.weight"], cfg.lm_eps).transpose(1, 2)
        k = rms_norm(k, w[p + "self_attn.k_norm.weight"], cfg.lm_eps).transpose(1, 2)
        v = v.transpose(1, 2)
        q = q * cos + _rotate_half(q) * sin
        k = k * cos + _rotate_half(k) * sin
        rep = Hq // Hkv
        k = k.repeat_interleave(rep, dim=1)
        v = v.repeat_interleave(rep, dim=1)
        s = (q @ k.transpose(-1, -2)) * (hd ** -0.5)
        s = s.masked_fill(~mask, float("-inf"))
        att = (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(B, S, Hq * hd)
        x = x + lin(att, p + "self_attn.o_proj.weight", i, "o_proj")
        h = rms_norm(x, w[p + "post_attention_layernorm.weight"], cfg.lm_eps)
        h = F.silu(lin(h, p + "mlp.gate_proj.weight", i, "gate_proj")) * lin(h, p + "mlp.up_proj.weight", i, "up_proj")
        x = x + lin(h, p + "mlp.down_proj.weight", i, "down_proj")
    return rms_norm(x, w["model.norm.weight"], cfg.lm_eps)


def causal_lm_loss(logits: Tensor, labels: Tensor, num_items_in_batch=None) -> Tensor:
    """fp32 upcast, shift-by-one, ignore -100; mean, or sum / num_items (HF:loss_utils:28-67)."""
    V = logits.shape[-1]
    shift = F.pad(labels, (0, 1), value=-100)[..., 1:].reshape(-1)
    if num_items_in_batch is None:
        return F.cross_entropy(logits.float().view(-1, V), shift, ignore_index=-100, reduction="mean")
    return F.cross_entropy(logits.float().view(-1, V), shift, ignore_index=-100, reduction="sum") / num_items_in_batch


# --------------------------------------------------------------------------------------
# whole forward (tiny_audio/asr_modeling.py:481-533) and one train step
# --------------------------------------------------------------------------------------
def model_forward(W, batch: Dict[str, Tensor], cfg: PathConfig = FULL, num_items_in_batch=None,
                  return_parts: bool = False):
    """batch keys: input_features (B,n_mels,T) [or waveform (B,L)], input_ids, labels,
    attention_mask (optional), audio_token_counts (optional), frame_keep_mask (optional, (B,S_e) of {0,1}: the Bernoulli draw of
    `_maybe_drop_audio_tokens`, tiny_audio/asr_modeling.py:458-479 -- whole encoder frames are zeroed, no rescale)."""
    parts = {}
    if "input_features" in batch:
        mel = batch["input_features"].float()
    else:
        mel = log_mel(batch["waveform"], cfg)
    parts["mel"] = mel
    with torch.no_grad():
        enc = encoder_forward(W["encoder"], mel, cfg)
    parts["encoder_out"] = enc
    if batch.get("frame_keep_mask") is not None:
        enc = enc * batch["frame_keep_mask"].to(enc.dtype).unsqueeze(-1)
    kind, aux = projector_kind(W["projector"]), None
    if kind == "qformer":
        audio = qformer_projector_forward(W["projector"], enc, cfg)
    elif kind == "mosa":
        audio = mosa_projector_forward(W["projector"], enc, cfg)
    elif kind == "moe":
        audio, aux = moe_projector_forward(W["projector"], enc, cfg, training=batch.get("projector_training", True))
    else:
        audio = projector_forward(W["projector"], enc, cfg)
    parts["projector_out"] = audio
    ids = batch["input_ids"]
    counts = batch.get("audio_token_counts")
    if counts is None:
        counts = (ids == cfg.audio_token_id).sum(-1)
    packed = gather_audio_embeds(audio, counts)
    emb = F.embedding(ids, W["lm"]["model.embed_tokens.weight"])
    emb = scatter_audio(emb, ids, packed, cfg.audio_token_id)
    parts["inputs_embeds"] = emb
    hid = lm_forward(W["lm"], emb, cfg, batch.get("attention_mask"), W.get("lora"))
    parts["hidden"] = hid
    logits = F.linear(hid, W["lm"]["lm_head.weight"])
    loss = None
    if batch.get("labels") is not None:
        loss = causal_lm_loss(logits, batch["labels"], num_items_in_batch)
        if aux is not None:                  # projector.get_aux_loss() is added to the LM loss (asr_modeling.py:528-531)
            loss = loss + aux
            parts["aux_loss"] = aux
    if return_parts:
        return loss, logits, parts
    return loss, logits


@torch.no_grad()
def greedy_generate(W, batch, cfg: PathConfig = FULL, max_new_tokens: int = 8, attention_mask: Optional[Tensor] = None):
    """Greedy ids by re-running the whole forward per token (reference: asr_modeling.py:562-646, greedy defaults).
    Returns (ids [B, T_new], top-1 minus top-2 logit margin [B, T_new]).  `attention_mask` (B, S0): LEFT-padded prompts (ragged
    batches) -- HF generate masks the padding keys and counts rotary positions from each sequence's first real token."""
    ids = batch["input_ids"].clone()
    am = attention_mask.clone() if attention_mask is not None else None
    counts = batch["audio_token_counts"]
    mel = batch["input_features"].float() if "input_features" in batch else log_mel(batch["waveform"], cfg)
    audio = projector_forward(W["projector"], encoder_forward(W["encoder"], mel, cfg), cfg)
    packed = gather_audio_embeds(audio, counts)
    out, margins = [], []
    for _ in range(max_new_tokens):
        emb = F.embedding(ids, W["lm"]["model.embed_tokens.weight"])
        emb = scatter_audio(emb, ids, packed, cfg.audio_token_id)
        pos = (am.cumsum(-1) - 1).clamp(min=0) if am is not None else None
        hid = lm_forward(W["lm"], emb, cfg, attention_mask=am, position_ids=pos)[:, -1]
        logits = F.linear(hid, W["lm"]["lm_head.weight"])
        top = logits.topk(2, -1).values
        nxt = logits.argmax(-1)
        out.append(nxt)
        margins.append(top[:, 0] - top[:, 1])
        ids = torch.cat([ids, nxt[:, None]], 1)
        if am is not None:
            am = torch.cat([am, torch.ones_like(am[:, :1])], 1)
    return torch.stack(out, 1), torch.stack(margins, 1)


def clip_grad_norm(grads: Dict[str, Tensor], max_norm: float) -> Tuple[Tensor, float]:
    """torch.nn.utils.clip_grad_norm_ semantics: total 2-norm; scale by max_norm/(norm+1e-6) clamped to 1."""
    total = torch.sqrt(sum((g.float() ** 2).sum() for g in grads.values()))
    coef = min(1.0, float(max_norm / (total + 1e-6)))
    return total, coef


def adamw_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float, beta1=0.9, beta2=0.999,
               eps=1e-8, weight_decay=0.0):
    """torch.optim.AdamW single-tensor semantics (decoupled decay, bias correction, eps outside sqrt)."""
    p = p * (1.0 - lr * weight_decay)
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)) + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v


def train_step(W, batch, cfg: PathConfig = FULL, lr=1e-3, max_grad_norm=1.0, weight_decay=0.0, state=None,
               num_items_in_batch=None, train_lm: bool = False):
    """One optimiser step on the projector (configs 1-3).  Returns loss, grads, new params, state.
    train_lm (configs/experiments/embedded.yaml:19-33, `freeze_language_model: false`; asr_modeling.py:251-254): the whole
    decoder is trainable too -- its gradients are returned as `lm_grads` under HF parameter names (the tied
    embed_tokens / lm_head table receives both contributions); the optimiser update below still covers the projector only."""
    proj = {k: v.detach().clone().requires_grad_(True) for k, v in W["projector"].items()}
    lm_w = W["lm"]
    if train_lm:
        lm_w = {k: v.detach().clone().requires_grad_(True) for k, v in W["lm"].items() if k != "lm_head.weight"}
        lm_w["lm_head.weight"] = lm_w["model.embed_tokens.weight"]          # tie_word_embeddings
    W2 = {"encoder": W["encoder"], "lm": lm_w, "projector": proj}
    lora = None
    if W.get("lora") is not None:
        lora = {"scaling": W["lora"]["scaling"],
                "A": {k: v.detach().clone().requires_grad_(True) for k, v in W["lora"]["A"].items()},
                "B": {k: v.detach().clone().requires_grad_(True) for k, v in W["lora"]["B"].items()}}
        W2["lora"] = lora
    loss, _ = model_forward(W2, batch, cfg, num_items_in_batch)
    loss.backward()
    grads = {k: (v.grad.detach() if v.grad is not None else torch.zeros_like(v)) for k, v in proj.items()}   # unselected MoE experts: None
    lora_grads = None
    if lora is not None:
        lora_grads = {"A": {k: v.grad.detach() for k, v in lora["A"].items()}, "B": {k: v.grad.detach() for k, v in lora["B"].items()}}
    gnorm, coef = clip_grad_norm(grads, max_grad_norm)
    if state is None:
        state = {"step": 0, "m": {k: torch.zeros_like(v) for k, v in proj.items()},
                 "v": {k: torch.zeros_like(v) for k, v in proj.items()}}
    state["step"] += 1
    new_p = {}
    for k in proj:
        # HF Trainer default param groups: no decay on norm gains / biases (HF:trainer.py:1280-1290)
        wd = 0.0 if k.startswith("norm") else weight_decay
        new_p[k], state["m"][k], state["v"][k] = adamw_step(
            proj[k].detach(), grads[k] * coef, state["m"][k], state["v"][k], state["step"], lr,
            weight_decay=wd)
    lm_grads = {k: v.grad.detach() for k, v in lm_w.items() if k != "lm_head.weight"} if train_lm else None
    return {"loss": loss.detach(), "grads": grads, "grad_norm": gnorm, "clip_coef": coef,
            "params": new_p, "state": state, "lora_grads": lora_grads, "lm_grads": lm_grads}


# --------------------------------------------------------------------------------------
# synthetic batch (scripts/debug/check_gradient_flow.py:81-145 is the template)
# --------------------------------------------------------------------------------------
# Qwen3 chat-template token ids (public tokenizer):  <|im_start|>=151644 <|im_end|>=151645
# "user"=872 "assistant"=77091 "\n"=198 ; " Transcribe the speech to text" is 6 tokens whose ids
# cannot be verified offline -- a fixed 6-id stand-in is used and recorded here.
IM_START, IM_END, NL, USER, ASSISTANT = 151644, 151645, 198, 872, 77091
PROMPT_TAIL = [4058, 3114, 279, 8806, 311, 1467]   # stand-in for " Transcribe the speech to text"
THINK_EMPTY = [151667, 271, 151668, 271]            # "<think>\n\n</think>\n\n" (enable_thinking=False)


def synthetic_batch(cfg: PathConfig, batch: int, clip_seconds: float, seed: int = 0, response_len: int = 64,
                    pad_to_seconds: Optional[float] = None, projector: str = "mlp") -> Dict[str, Tensor]:
    """Equal-length clips, 0.1*N(0,1) waveform, chat-template prompt with N_a <audio> tokens, R seeded
    response ids; labels = -100 except response + <|im_end|>."""
    rng = np.random.default_rng(seed)
    L = int(round(clip_seconds * cfg.sample_rate))
    Lp = int(round((pad_to_seconds or clip_seconds) * cfg.sample_rate))
    wave = np.zeros((batch, Lp), dtype=np.float32)
    wave[:, :L] = 0.1 * rng.standard_normal((batch, L)).astype(np.float32)
    mel_len = L // cfg.hop
    n_a = int(projector_output_length(encoder_output_length(mel_len), cfg.proj_k))
    if projector == "qformer":
        n_a = int(qformer_output_length(encoder_output_length(mel_len)))
    elif projector == "mosa":
        n_a = int(mosa_output_length(encoder_output_length(mel_len)))

    def tid(t):   # map real-tokenizer ids into a reduced vocab deterministically
        return t if t < cfg.vocab - 1 else (t % (cfg.vocab - 1))

    ids, labels = [], []
    for b in range(batch):
        resp = rng.integers(0, min(cfg.vocab - 1, 151643), size=response_len).tolist()
        prompt = [tid(IM_START), tid(USER), tid(NL)] + [cfg.audio_token_id] * n_a + [tid(t) for t in PROMPT_TAIL] \
            + [tid(IM_END), tid(NL), tid(IM_START), tid(ASSISTANT), tid(NL)] + [tid(t) for t in THINK_EMPTY]
        tail = resp + [tid(IM_END), tid(NL)]
        ids.append(prompt + tail)
        labels.append([-100] * len(prompt) + resp + [tid(IM_END)] + [-100])
    return {
        "waveform": torch.from_numpy(wave),
        "sample_lengths": torch.full((batch,), L, dtype=torch.int64),
        "input_ids": torch.tensor(ids, dtype=torch.int64),
        "labels": torch.tensor(labels, dtype=torch.int64),
        "attention_mask": torch.ones(batch, len(ids[0]), dtype=torch.int64),
        "audio_token_counts": torch.full((batch,), n_a, dtype=torch.int64),
    }
