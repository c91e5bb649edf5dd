"""ASRProcessor + the GPU-mel feature extractor (drop-in for `tiny_audio/asr_processing.py:17-128`).

`ASRProcessor` keeps the reference's call contract: feature-extract the audio, derive the number of `<audio>`
placeholders from the attention mask with the conv-length formula and `projector.get_output_length`, and build
the chat prompt.  `WaveformFeatureExtractor` is the fast-path extractor: it has the call signature of
`WhisperFeatureExtractor` (what `scripts/train.py:327-333` calls inside the dataloader workers) but returns the
zero-padded waveform as `input_features` -- the log-mel then runs on the GPU inside `ASRModel.forward`
(`ta_logmel_fwd`) instead of on the CPU workers, which the reference names as its bottleneck
(configs/experiments/embedded.yaml:37-41).  The frame mask it returns is the same arithmetic as
HF:models/whisper/feature_extraction_whisper.py:328-337.
"""
from __future__ import annotations

from typing import Optional, Union

import numpy as np
import torch
import transformers
from transformers import ProcessorMixin
from transformers.feature_extraction_utils import BatchFeature

from .asr_config import DEFAULT_ENCODER_CONV_LAYERS, ASRConfig, compute_encoder_output_length


class WhisperFeatureExtractor:  # noqa: D101  (the class name is what scripts/train.py:260-264 inspects)
    """Waveform pass-through with Whisper's padding rules.  `type(fe).__name__ == "WhisperFeatureExtractor"` on
    purpose: the reference collator picks `padding="max_length"` (30 s) by that name, and this keeps its behaviour."""

    sampling_rate = 16000
    hop_length = 160
    n_fft = 400
    chunk_length = 30
    feature_size = 128
    padding_value = 0.0
    returns_waveform = True

    def __init__(self, feature_size: int = 128, sampling_rate: int = 16000, hop_length: int = 160, chunk_length: int = 30,
                 **_):
        self.feature_size, self.sampling_rate, self.hop_length, self.chunk_length = feature_size, sampling_rate, hop_length, chunk_length
        self.n_samples = chunk_length * sampling_rate
        self.nb_max_frames = self.n_samples // hop_length

    def __call__(self, raw_speech, sampling_rate=None, padding="max_length", max_length=None, truncation=True,
                 return_attention_mask=None, return_tensors=None, **_):
        if sampling_rate is not None and sampling_rate != self.sampling_rate:
            raise ValueError(f"expected {self.sampling_rate} Hz audio, got {sampling_rate}")
        if isinstance(raw_speech, np.ndarray) and raw_speech.ndim == 1 or (
                isinstance(raw_speech, (list, tuple)) and len(raw_speech) and np.isscalar(raw_speech[0])):
            raw_speech = [raw_speech]
        clips = [np.asarray(c, dtype=np.float32).reshape(-1) for c in raw_speech]
        if padding == "max_length" or padding is True and max_length:
            target = max_length or self.n_samples
        elif padding in ("longest", True):
            target = max(len(c) for c in clips)
        else:   # no padding: all clips must already agree
            target = max(len(c) for c in clips)
        if truncation and padding == "max_length":
            clips = [c[:target] for c in clips]
        wave = np.zeros((len(clips), target), dtype=np.float32)
        mask = np.zeros((len(clips), target), dtype=np.int32)
        for i, c in enumerate(clips):
            wave[i, : len(c)] = c
            mask[i, : len(c)] = 1
        frame_mask = mask[:, :: self.hop_length]
        if target % self.hop_length != 0:
            frame_mask = frame_mask[:, :-1]
        out = {"input_features": wave}
        if return_attention_mask:
            out["attention_mask"] = frame_mask
        return BatchFeature(out, tensor_type=return_tensors)

    def save_pretrained(self, path, **_):
        import json
        import os
        os.makedirs(path, exist_ok=True)
        with open(os.path.join(path, "preprocessor_config.json"), "w") as f:
            json.dump({"feature_extractor_type": "WhisperFeatureExtractor", "feature_size": self.feature_size,
                       "sampling_rate": self.sampling_rate, "hop_length": self.hop_length, "chunk_length": self.chunk_length,
                       "n_fft": self.n_fft, "tiny_audio_b200_waveform_passthrough": True}, f)


WaveformFeatureExtractor = WhisperFeatureExtractor


class ASRProcessor(ProcessorMixin):
    """Same constructor and __call__ as the reference's ASRProcessor."""

    attributes = ["feature_extractor", "tokenizer"]
    feature_extractor_class = "AutoFeatureExtractor"
    tokenizer_class = "AutoTokenizer"
    AUDIO_TOKEN = "<audio>"
    TRANSCRIBE_PROMPT = "Transcribe the speech to text"

    def __init__(self, feature_extractor, tokenizer, projector=None, encoder_conv_layers: Optional[list] = None):
        self.feature_extractor = feature_extractor
        self.tokenizer = tokenizer
        self.audio_token_id = tokenizer.convert_tokens_to_ids(self.AUDIO_TOKEN)
        self.projector = projector
        self.encoder_conv_layers = encoder_conv_layers or DEFAULT_ENCODER_CONV_LAYERS

    def _compute_encoder_output_length(self, mel_length: int) -> int:
        return compute_encoder_output_length(mel_length, self.encoder_conv_layers)

    def __call__(self, audio: Optional[Union[list, "torch.Tensor"]] = None, text: Optional[str] = None,
                 system_prompt: Optional[str] = None, return_tensors: str = "pt", **kwargs) -> dict:
        out = {}
        n_audio = 0
        if audio is not None:
            feats = self.feature_extractor(audio, sampling_rate=getattr(self.feature_extractor, "sampling_rate", 16000),
                                           return_attention_mask=True, return_tensors=return_tensors, **kwargs)
            out["input_features"] = feats["input_features"]
            out["audio_attention_mask"] = feats["attention_mask"]
            real_frames = int(feats["attention_mask"].sum(dim=-1).max().item())
            n_audio = self.projector.get_output_length(self._compute_encoder_output_length(real_frames))
        if n_audio > 0:
            content = self.AUDIO_TOKEN * n_audio + (" " + self.TRANSCRIBE_PROMPT if self.TRANSCRIBE_PROMPT else "")
        else:
            content = self.TRANSCRIBE_PROMPT or ""
        messages = ([{"role": "system", "content": system_prompt}] if system_prompt else []) + [{"role": "user", "content": content}]
        if text is not None:
            messages.append({"role": "assistant", "content": text})
        tok = self.tokenizer.apply_chat_template(messages, tokenize=True, add_generation_prompt=(text is None),
                                                 return_tensors=return_tensors, enable_thinking=False)
        ids = tok if isinstance(tok, torch.Tensor) else tok.get("input_ids", getattr(tok, "input_ids", None))
        if ids.dim() == 1:
            ids = ids.unsqueeze(0)
        out["input_ids"] = ids
        out["attention_mask"] = torch.ones_like(ids)
        return out


try:
    ASRProcessor.register_for_auto_class()
    transformers.AutoProcessor.register(ASRConfig, ASRProcessor, exist_ok=True)
except Exception:
    pass
