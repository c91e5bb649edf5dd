"""ASRConfig -- the flag system of the drop-in surface.

Mirrors the reference's `tiny_audio/asr_config.py:22-223` field for field (names, defaults, meaning), so
that `scripts/train.py:494-506` can build it from the same Hydra dicts, and keeps the conv-length formula
(`asr_config.py:9-19`) that the collator and the processor use for the audio-token arithmetic.
"""
from __future__ import annotations

from typing import Optional

import transformers

# (padding, kernel, stride) of the two encoder convs (Whisper / GLM-ASR front end)
DEFAULT_ENCODER_CONV_LAYERS = [(1, 3, 1), (1, 3, 2)]

_ALL_LINEAR = ["q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj"]

_GENERATION_DEFAULTS = dict(num_beams=1, max_new_tokens=128, min_new_tokens=0, repetition_penalty=1.0,
                            length_penalty=1.0, no_repeat_ngram_size=0, use_cache=True)


def compute_encoder_output_length(mel_length, conv_layers=None):
    """`(L + 2p - (k-1) - 1) // s + 1` per conv layer; works on ints and on integer tensors."""
    out = mel_length
    for pad, kernel, stride in (conv_layers if conv_layers is not None else DEFAULT_ENCODER_CONV_LAYERS):
        out = (out + 2 * pad - (kernel - 1) - 1) // stride + 1
    return out


def _rebuild_sub_config(cfg):
    if isinstance(cfg, dict):
        model_type = cfg.get("model_type")
        if model_type:
            return transformers.AutoConfig.for_model(model_type).__class__(**cfg)
    return cfg


class ASRConfig(transformers.PretrainedConfig):
    """Audio encoder + projector + text decoder + generation + LoRA flags (same surface as the reference)."""

    model_type = "asr_model"
    is_composition = True

    def __init__(
        self,
        audio_model_id: str = "zai-org/GLM-ASR-Nano-2512",
        text_model_id: str = "Qwen/Qwen3-0.6B",
        attn_implementation: str = "flash_attention_2",
        model_dtype: str = "bfloat16",
        num_beams: Optional[int] = None,
        system_prompt: str = "You are a helpful assistant.",
        encoder_dim: Optional[int] = None,
        llm_dim: Optional[int] = None,
        encoder_conv_layers: Optional[list] = None,
        audio_sample_rate: int = 16000,
        projector_pool_stride: int = 4,
        downsample_rate: int = 5,
        projector_hidden_dim: Optional[int] = None,
        projector_type: str = "mlp",
        audio_token_dropout: float = 0.0,
        num_experts: int = 4,
        num_experts_per_tok: int = 2,
        router_aux_loss_coef: float = 0.01,
        qformer_window_size: int = 15,
        qformer_hidden_size: Optional[int] = None,
        qformer_num_layers: int = 2,
        qformer_num_heads: int = 16,
        qformer_intermediate_size: Optional[int] = None,
        use_lora: bool = False,
        lora_rank: int = 8,
        lora_alpha: int = 32,
        lora_dropout: float = 0.0,
        lora_target_modules: Optional[list] = None,
        freeze_projector: bool = False,
        freeze_language_model: bool = True,
        do_sample: bool = False,
        temperature: Optional[float] = None,
        top_p: Optional[float] = None,
        top_k: Optional[int] = None,
        max_new_tokens: Optional[int] = None,
        min_new_tokens: Optional[int] = None,
        repetition_penalty: Optional[float] = None,
        length_penalty: Optional[float] = None,
        no_repeat_ngram_size: Optional[int] = None,
        use_cache: Optional[bool] = None,
        **kwargs,
    ):
        local = dict(locals())
        for name in ("audio_model_id", "text_model_id", "attn_implementation", "model_dtype", "system_prompt", "encoder_dim",
                     "llm_dim", "audio_sample_rate", "projector_pool_stride", "downsample_rate", "projector_hidden_dim",
                     "projector_type", "audio_token_dropout", "num_experts", "num_experts_per_tok", "router_aux_loss_coef",
                     "qformer_window_size", "qformer_hidden_size", "qformer_num_layers", "qformer_num_heads",
                     "qformer_intermediate_size", "use_lora", "lora_rank", "lora_alpha", "lora_dropout", "freeze_projector",
                     "freeze_language_model", "do_sample", "temperature", "top_p", "top_k"):
            setattr(self, name, local[name])
        self.encoder_conv_layers = encoder_conv_layers or DEFAULT_ENCODER_CONV_LAYERS
        self.lora_target_modules = lora_target_modules or list(_ALL_LINEAR)
        # greedy-decoding defaults unless given explicitly (kept out of **kwargs so the base class cannot undo them)
        for name, default in _GENERATION_DEFAULTS.items():
            value = local[name]
            setattr(self, name, default if value is None else value)

        audio_config = kwargs.pop("audio_config", None)
        if audio_config is None:
            audio_config = transformers.AutoConfig.from_pretrained(audio_model_id)
            audio_config.dtype = model_dtype
        text_config = kwargs.pop("text_config", None)
        if text_config is None:
            text_config = transformers.AutoConfig.from_pretrained(text_model_id, trust_remote_code=True)
            text_config.dtype = model_dtype
        self.audio_config = _rebuild_sub_config(audio_config)
        self.text_config = _rebuild_sub_config(text_config)

        super().__init__(**kwargs)

        self.encoder = self.audio_config   # the HF pipeline resolves the feature extractor through config.encoder
        self.auto_map = {
            "AutoConfig": "asr_config.ASRConfig",
            "AutoModel": "asr_modeling.ASRModel",
            "AutoModelForSpeechSeq2Seq": "asr_modeling.ASRModel",
            "AutoProcessor": "asr_processing.ASRProcessor",
        }
        self.custom_pipelines = {
            "automatic-speech-recognition": {
                "impl": "asr_pipeline.ASRPipeline", "pt": ["AutoModelForSpeechSeq2Seq"], "tf": [], "type": "audio"}
        }
        self.architectures = ["ASRModel"]
        self.pipeline_tag = "automatic-speech-recognition"


    def to_diff_dict(self):
        """HF serialises the difference to a default-constructed config; a default ASRConfig() resolves the two tower configs
        from the hub, which is unavailable offline -- fall back to the full dict there (loads back identically)."""
        try:
            return super().to_diff_dict()
        except OSError:
            return self.to_dict()


try:  # registering twice (e.g. next to the reference in one process) is tolerated by transformers 5.x
    transformers.AutoConfig.register("asr_model", ASRConfig, exist_ok=True)
except TypeError:  # older signature
    transformers.AutoConfig.register("asr_model", ASRConfig)
