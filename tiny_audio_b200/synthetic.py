"""Offline construction helpers: random-initialised towers of the named architecture (there is no network for
checkpoints) and synthetic 16 kHz batches of the shape BASELINE.json names.  Used by bench.py and the tests;
the seams overridden here are the same four loader methods the reference exposes.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from .engine import PathDims

# Qwen3 chat-template ids (public tokenizer); the 6-token instruction is a fixed stand-in (no tokenizer offline)
IM_START, IM_END, NL, USER, ASSISTANT = 151644, 151645, 198, 872, 77091
PROMPT_TAIL = [4058, 3114, 279, 8806, 311, 1467]
THINK_EMPTY = [151667, 271, 151668, 271]


def num_audio_tokens(n_samples: int, hop: int = 160, k: int = 4) -> int:
    mel = n_samples // hop
    enc = (mel + 2 - 3) // 2 + 1          # conv1 keeps the length, conv2 halves it (asr_config.py:9-19)
    return (enc - k) // k + 1


def synthetic_batch(dims: PathDims, batch: int, clip_seconds: float, seed: int = 0, response_len: int = 64,
                    pad_to_seconds: Optional[float] = None, pin: bool = False, projector: str = "mlp") -> Dict[str, torch.Tensor]:
    """Equal-length clips `0.1*N(0,1)`, prompt = chat template with N_a `<audio>` tokens, R seeded response ids,
    labels = -100 except response + <|im_end|>  (SURVEY.md section 8d)."""
    rng = np.random.default_rng(seed)
    sr = 16000
    n = int(round(clip_seconds * sr))
    n_pad = int(round((pad_to_seconds or clip_seconds) * sr))
    wave = np.zeros((batch, n_pad), dtype=np.float32)
    wave[:, :n] = 0.1 * rng.standard_normal((batch, n)).astype(np.float32)
    n_a = num_audio_tokens(n, dims.hop, dims.proj_k)
    if projector == "qformer":                      # ceil(S_e / 15) windows x 3 queries (projectors.py:422-430)
        n_a = (((n // dims.hop + 2 - 3) // 2 + 1) + 14) // 15 * 3
    elif projector == "mosa":                       # two k=3, s=2, p=1 convolutions (projectors.py:172-177)
        n_a = (n // dims.hop + 2 - 3) // 2 + 1
        for _ in range(2):
            n_a = (n_a + 2 - 3) // 2 + 1
    V = dims.vocab

    def tid(t):
        return t if t < V - 1 else t % (V - 1)

    ids, labels = [], []
    for _ in range(batch):
        resp = rng.integers(0, min(V - 1, 151643), size=response_len).tolist()
        prompt = ([tid(IM_START), tid(USER), tid(NL)] + [dims.audio_token_id] * n_a + [tid(t) for t in PROMPT_TAIL]
                  + [tid(IM_END), tid(NL), tid(IM_START), tid(ASSISTANT), tid(NL)] + [tid(t) for t in THINK_EMPTY])
        ids.append(prompt + resp + [tid(IM_END), tid(NL)])
        labels.append([-100] * len(prompt) + resp + [tid(IM_END)] + [-100])
    out = {
        "input_features": torch.from_numpy(wave),                       # waveform fast path (B, L)
        "input_ids": torch.tensor(ids, dtype=torch.int64),
        "labels": torch.tensor(labels, dtype=torch.int64),
        "attention_mask": torch.ones(batch, len(ids[0]), dtype=torch.int64),
        "audio_token_counts": torch.full((batch,), n_a, dtype=torch.int64),
    }
    if pin and torch.cuda.is_available():
        out = {k: v.pin_memory() for k, v in out.items()}
    return out


class StubTokenizer:
    """The attributes ASRModel reads from a tokenizer (asr_modeling.py:163-168, 305-342 of the reference)."""
    pad_token = "<|finetune_right_pad_id|>"
    eos_token = "<|im_end|>"
    padding_side = "right"
    chat_template = ""
    additional_special_tokens = ["<audio>"]
    bos_token_id = None

    def __init__(self, vocab: int, audio_token_id: int):
        self.vocab, self.audio_token_id = vocab, audio_token_id
        self.pad_token_id = min(151643, vocab - 2)
        self.eos_token_id = IM_END if IM_END < vocab - 1 else IM_END % (vocab - 1)

    def convert_tokens_to_ids(self, t):
        return {"<audio>": self.audio_token_id, "<|im_end|>": self.eos_token_id, "<|endoftext|>": self.pad_token_id}.get(t)

    def get_vocab(self):
        return {"<audio>": self.audio_token_id}

    def __len__(self):
        return self.vocab

    def save_pretrained(self, *a, **k):
        return None


def build_offline_model(dims: PathDims, device="cuda", seed: int = 1234, enc_state=None, lm_state=None, proj_state=None,
                        audio_token_dropout: float = 0.0, projector_type: str = "mlp", use_lora: bool = False,
                        freeze_projector: bool = False, freeze_language_model: bool = True, **config_extras):
    """ASRModel (tiny_audio_b200.asr_modeling) with GLM-ASR / Qwen3 modules of the given dims, random (seeded) or
    supplied weights, fp32 masters -- no network, no checkpoints."""
    from transformers import GlmAsrEncoderConfig, Qwen3Config, Qwen3ForCausalLM
    from transformers.models.glmasr.modeling_glmasr import GlmAsrEncoder

    from .asr_config import ASRConfig
    from .asr_modeling import ASRModel
    from .asr_processing import WaveformFeatureExtractor

    enc_cfg = GlmAsrEncoderConfig(hidden_size=dims.enc_dim, intermediate_size=dims.enc_ffn, num_hidden_layers=dims.enc_layers,
                                  num_attention_heads=dims.enc_heads, num_key_value_heads=dims.enc_heads,
                                  num_mel_bins=dims.n_mels)
    txt_cfg = Qwen3Config(hidden_size=dims.lm_dim, intermediate_size=dims.lm_ffn, num_hidden_layers=dims.lm_layers,
                          num_attention_heads=dims.lm_heads, num_key_value_heads=dims.lm_kv_heads, head_dim=dims.lm_head_dim,
                          vocab_size=dims.vocab, rms_norm_eps=dims.lm_eps, tie_word_embeddings=True,
                          max_position_embeddings=40960,
                          rope_parameters={"rope_theta": dims.lm_rope_theta, "rope_type": "default"})
    dev = torch.device(device)

    class _Offline(ASRModel):
        @classmethod
        def _load_audio_encoder(cls, config, dtype):
            torch.manual_seed(seed + 1)
            with torch.device(dev):
                m = GlmAsrEncoder._from_config(enc_cfg, attn_implementation="sdpa").to(dtype)
            if enc_state is not None:
                m.load_state_dict(enc_state, strict=True)
            m.requires_grad_(False)
            m.eval()
            return m

        @classmethod
        def _load_language_model(cls, config, dtype):
            torch.manual_seed(seed + 2)
            with torch.device(dev):
                m = Qwen3ForCausalLM._from_config(txt_cfg, attn_implementation="sdpa").to(dtype)
            if lm_state is not None:
                m.load_state_dict(lm_state, strict=True)
                m.tie_weights()
            if getattr(config, "freeze_language_model", True):
                m.requires_grad_(False)
                m.train(False)
            return m

        def _init_tokenizer(self, config):
            self.tokenizer = StubTokenizer(dims.vocab, dims.audio_token_id)
            self.audio_token_id = dims.audio_token_id

        def _create_feature_extractor(self, config):
            return WaveformFeatureExtractor(feature_size=dims.n_mels)

    _Offline.__name__ = "ASRModel"
    cfg = ASRConfig(audio_config=enc_cfg, text_config=txt_cfg, model_dtype="float32", attn_implementation="sdpa",
                    projector_type=projector_type, projector_pool_stride=dims.proj_k, projector_hidden_dim=dims.proj_hidden,
                    audio_token_dropout=audio_token_dropout, use_lora=use_lora, freeze_projector=freeze_projector,
                    freeze_language_model=freeze_language_model, **config_extras)
    torch.manual_seed(seed + 3)
    model = _Offline(cfg)
    if proj_state is not None:
        model.projector.load_state_dict(proj_state, strict=True)
    return model.to(dev)
