"""ASRModel -- drop-in for the reference's `tiny_audio/asr_modeling.py:47` on the training hot path.

Same constructor seams, attribute names and call surface as the reference (so `scripts/train.py` drives it
unchanged: SURVEY.md section 8b), but `forward` does not run the HF modules: the frozen GLM-ASR encoder, the
projector, the frozen Qwen3 decoder, the CE loss and the whole backward into the projector parameters run as
hand-written sm_100a kernels behind libtinyaudio_b200.so (tiny_audio_b200/engine.py).  The HF modules are kept
only as the owners of the fp32 master weights (`audio_tower`, `language_model`) -- for `.to()`, `state_dict`
naming, and weight loading -- and are never called on the hot path.

No CPU fallback: a forward with CPU tensors raises.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Optional

import torch
import torch.nn as nn
from transformers import PreTrainedModel
from transformers.generation import GenerationMixin
from transformers.modeling_outputs import CausalLMOutputWithPast

from . import lib as L
from .asr_config import ASRConfig, compute_encoder_output_length
from .engine import HotPath, PathDims
from .projectors import PROJECTOR_CLASSES

_PROJ_KEYS = ("linear_1.weight", "norm.weight", "linear_2.weight", "norm_2.weight")


def _gather_audio_embeds(audio_embeds: torch.Tensor, token_counts: torch.Tensor) -> torch.Tensor:
    """Reference helper (asr_modeling.py:27-44) restated with plain indexing: first `token_counts[i]` rows of each
    sample, zero rows when a count exceeds the available length.  The CUDA path implements the same index
    semantics inside `ta_audio_index` / `ta_embed_scatter`; this function exists for API parity and tests."""
    _, n, _ = audio_embeds.shape
    need = int(token_counts.max()) if token_counts.numel() else 0
    if need > n:
        audio_embeds = torch.nn.functional.pad(audio_embeds, (0, 0, 0, need - n))
        n = need
    keep = torch.arange(n, device=audio_embeds.device)[None, :] < token_counts[:, None]
    return audio_embeds[keep]


def _prepare_decoder(model, hot: HotPath, extra) -> str:
    """Bring the engine's decoder operands in line with the trainable decoder tensors `extra` (the tensors ASRModel._decoder_tensors
    returned): LoRA -> refresh the rank-padded A / B operands; unfrozen LM -> re-pack the bf16 operands when the optimiser moved the
    fp32 masters.  Returns the decoder mode: "frozen" | "lora" | "unfrozen"."""
    adapters = model.lora_adapters
    if adapters is not None:
        n = len(extra) // 2
        names = adapters.targets
        hot.lm.update_lora({t: a.detach() for t, a in zip(names, extra[:n])}, {t: b.detach() for t, b in zip(names, extra[n:])},
                           adapters.scaling)
        return "lora"
    if len(extra) > 0:
        model._refresh_packed_lm(hot)
        return "unfrozen"
    return "frozen"


def _decoder_grads(model, hot: HotPath, mode: str):
    """Gradients of the trainable decoder tensors, in ASRModel._decoder_tensors order (valid until the next engine call)."""
    if mode == "lora":
        adapters = model.lora_adapters
        ga, gb = hot.lm.lora_grads(adapters.scaling, {t: adapters.rank for t in adapters.targets})
        return [ga[t].clone() for t in adapters.targets] + [gb[t].clone() for t in adapters.targets]
    hg = hot.lm.hf_grads()          # views of the engine's flat buffer; scaled (= copied) in backward
    return [hg[n] for n, _ in model.language_model.named_parameters()]


class _FusedPathLoss(torch.autograd.Function):
    """loss = CE(Qwen3(scatter(projector(encoder(audio))))) with d(loss)/d(projector params) -- and d(loss)/d(trainable decoder
    tensors) -- computed in the same pass; backward only rescales the stored gradients by the incoming scalar.
    Tensor arguments: the 4 projector parameters, then ASRModel._decoder_tensors(): EITHER the stacked lora_A tensors followed by
    the stacked lora_B tensors OR (freeze_language_model=False) every Qwen3 parameter in `language_model.named_parameters()`
    order, OR nothing (frozen decoder)."""

    @staticmethod
    def forward(ctx, model, call, w1, n1, w2, n2, *extra):
        hot: HotPath = model._hot_path()
        params = {k: p.detach().float().contiguous() for k, p in zip(_PROJ_KEYS, (w1, n1, w2, n2))}
        need = any(ctx.needs_input_grad[2:6])
        grads = {k: torch.empty_like(v) for k, v in params.items()} if need else None
        mode = _prepare_decoder(model, hot, extra)
        need_dec = mode != "frozen" and any(ctx.needs_input_grad[6:])
        loss, _ = hot.forward_backward(proj_params=params, grads=grads, lm_backward=need_dec, train_lm=(need_dec and mode == "unfrozen"),
                                       **call)
        ctx.grads = grads
        ctx.dtypes = (w1.dtype, n1.dtype, w2.dtype, n2.dtype)
        ctx.dec_grads = _decoder_grads(model, hot, mode) if need_dec else None
        ctx.dec_dtypes = [t.dtype for t in extra]
        return loss.reshape(())

    @staticmethod
    def backward(ctx, gout):
        proj = (None,) * 4
        if ctx.grads is not None:
            proj = tuple((ctx.grads[k] * gout).to(dt) for k, dt in zip(_PROJ_KEYS, ctx.dtypes))
        dec = (None,) * len(ctx.dec_dtypes)
        if ctx.dec_grads is not None:
            dec = tuple((g * gout).to(dt) for g, dt in zip(ctx.dec_grads, ctx.dec_dtypes))
        return (None, None) + proj + dec


class _LmLossFn(torch.autograd.Function):
    """loss(audio_embeds) for any projector: <audio> scatter -> Qwen3 -> CE with d(loss)/d(audio_embeds) -- and d(loss)/d(trainable
    decoder tensors: LoRA A / B or the unfrozen Qwen3 parameters) -- computed in the same pass by the CUDA engine; autograd then
    continues into the projector that produced `audio_embeds`.  `audio` is None for a batch without audio (text-only forward, or
    `inputs_embeds` given in `call`)."""

    @staticmethod
    def forward(ctx, model, call, audio, *extra):
        hot: HotPath = model._hot_path()
        mode = _prepare_decoder(model, hot, extra)
        need_audio = audio is not None and ctx.needs_input_grad[2]
        need_dec = mode != "frozen" and any(ctx.needs_input_grad[3:])
        kw = dict(with_backward=need_audio, lm_backward=need_dec, train_lm=(need_dec and mode == "unfrozen"))
        if audio is not None:
            B, n_a, D = audio.shape
            flat = audio.detach().float().contiguous().view(B * n_a, D)
            loss, d_audio = hot.lm_loss_and_audio_grad(audio=flat, n_a=n_a, **kw, **call)
            ctx.d_audio = d_audio.view(B, n_a, D).clone() if d_audio is not None else None
            ctx.dtype = audio.dtype
        else:
            loss, _ = hot.lm_loss_and_audio_grad(audio=None, n_a=0, **kw, **call)
            ctx.d_audio = None
        ctx.dec_grads = _decoder_grads(model, hot, mode) if need_dec else None
        ctx.dec_dtypes = [t.dtype for t in extra]
        return loss.reshape(())

    @staticmethod
    def backward(ctx, gout):
        d_audio = (ctx.d_audio * gout).to(ctx.dtype) if ctx.d_audio is not None else None
        dec = (None,) * len(ctx.dec_dtypes)
        if ctx.dec_grads is not None:
            dec = tuple((g * gout).to(dt) for g, dt in zip(ctx.dec_grads, ctx.dec_dtypes))
        return (None, None, d_audio) + dec


class ASRModel(PreTrainedModel, GenerationMixin):
    """Audio encoder (frozen) + projector (trainable) + causal LM (frozen): the reference's composition."""

    config_class = ASRConfig
    base_model_prefix = "model"
    main_input_name = "input_features"
    _supports_flash_attn_2 = True
    supports_gradient_checkpointing = True
    _is_loading_from_pretrained: bool = False

    TRANSCRIBE_PROMPT = "Transcribe the speech to text"

    # ------------------------------------------------------------------ construction (same seams as the reference)
    def __init__(self, config: ASRConfig, **kwargs) -> None:
        super().__init__(config)
        self.system_prompt = config.system_prompt
        dtype = getattr(torch, config.model_dtype)
        self.audio_tower = self._load_audio_encoder(config, dtype)
        self.language_model = self._load_language_model(config, dtype)
        self._init_tokenizer(config)

        gc = self.language_model.generation_config
        self.generation_config = gc
        for name in ("max_new_tokens", "min_new_tokens", "num_beams", "do_sample", "temperature", "top_p", "top_k",
                     "use_cache", "length_penalty", "repetition_penalty", "no_repeat_ngram_size"):
            setattr(gc, name, getattr(config, name))
        eos = [self.tokenizer.convert_tokens_to_ids(t) for t in ("<|im_end|>", "<|endoftext|>")]
        gc.eos_token_id = [t for t in eos if t is not None]
        gc.pad_token_id = self.tokenizer.pad_token_id

        self.feature_extractor = self._create_feature_extractor(config)
        self.projector = self._create_projector(config, dtype)
        if getattr(config, "use_lora", False) and not getattr(type(self), "_is_loading_from_pretrained", False):
            self._setup_lora(config)
        if getattr(config, "freeze_projector", False):
            self.projector.requires_grad_(False)
        self._no_split_modules = getattr(self.language_model, "_no_split_modules", [])
        self._hot: Optional[HotPath] = None
        self._hot_key = None

    def _create_feature_extractor(self, config: ASRConfig):
        from transformers import AutoFeatureExtractor
        fe = AutoFeatureExtractor.from_pretrained(config.audio_model_id)
        if "whisper" not in config.audio_model_id.lower():
            fe.padding = False
        return fe

    @classmethod
    def _load_audio_encoder(cls, config: ASRConfig, dtype: torch.dtype) -> nn.Module:
        kw = dict(attn_implementation=config.attn_implementation, low_cpu_mem_usage=True, dtype=dtype)
        name = config.audio_model_id.lower()
        if "glm" in name:
            from transformers import AutoModelForSeq2SeqLM
            full = AutoModelForSeq2SeqLM.from_pretrained(config.audio_model_id, trust_remote_code=True, **kw)
            enc = full.audio_tower
            full.language_model = None
            full.multi_modal_projector = None
            del full
        else:
            raise NotImplementedError("tiny_audio_b200's hot path is built for the GLM-ASR encoder "
                                      f"(audio_model_id={config.audio_model_id!r}); see DESIGN.md")
        enc.requires_grad_(False)
        enc.eval()
        return enc

    @classmethod
    def _load_language_model(cls, config: ASRConfig, dtype: torch.dtype) -> PreTrainedModel:
        from transformers import AutoModelForCausalLM
        lm = AutoModelForCausalLM.from_pretrained(config.text_model_id, attn_implementation=config.attn_implementation,
                                                  trust_remote_code=True, low_cpu_mem_usage=True, dtype=dtype)
        lm.config.use_cache = getattr(config, "use_cache", True)
        if getattr(config, "freeze_language_model", True):
            lm.requires_grad_(False)
            lm.train(False)
        # else: full decoder fine-tuning (configs/experiments/embedded.yaml:19-33): the parameters stay trainable fp32 masters;
        # the CUDA engine computes their gradients and re-packs its bf16 operands after every update (engine.py:PackedLM)
        return lm

    def _create_projector(self, config: ASRConfig, dtype: torch.dtype) -> nn.Module:
        if config.encoder_dim is None:
            ec = self.audio_tower.config
            config.encoder_dim = getattr(ec, "hidden_size", None) or getattr(ec, "d_model", None)
        if config.llm_dim is None:
            dc = self.language_model.config
            config.llm_dim = getattr(dc, "hidden_size", None) or getattr(dc, "d_model", None)
        if config.encoder_dim is None or config.llm_dim is None:
            raise ValueError("could not infer encoder_dim / llm_dim; set them in the config")
        kind = getattr(config, "projector_type", "mlp")
        if kind not in PROJECTOR_CLASSES:
            raise ValueError(f"Unknown projector_type: {kind}. Valid options: {list(PROJECTOR_CLASSES.keys())}")
        proj = PROJECTOR_CLASSES[kind](config)
        device = next(self.language_model.parameters()).device
        return proj.to(device=device, dtype=dtype)

    def _setup_lora(self, config: ASRConfig):
        """Attach LoRA adapters to the decoder (reference: asr_modeling.py:289-301 via peft; restated in lora.py).  The
        container is registered under the language model so that its parameters are named `language_model.*`, which is
        what scripts/train.py:413-418 keys the decoder learning-rate group on."""
        from .lora import LoraAdapters
        adapters = LoraAdapters(self.language_model.config, rank=config.lora_rank, alpha=config.lora_alpha,
                                target_modules=config.lora_target_modules, dropout=config.lora_dropout)
        dev = next(self.language_model.parameters()).device
        self.language_model.add_module("lora_adapters", adapters.to(dev))
        self._hot = None

    @property
    def lora_adapters(self):
        return getattr(self.language_model, "lora_adapters", None)

    def _init_tokenizer(self, config: ASRConfig):
        from transformers import AutoTokenizer
        tok = AutoTokenizer.from_pretrained(config.text_model_id, trust_remote_code=True)
        if tok.pad_token is None or tok.pad_token_id == tok.eos_token_id:
            if "<|finetune_right_pad_id|>" in tok.get_vocab():
                tok.pad_token = "<|finetune_right_pad_id|>"
            elif tok.pad_token is None:
                tok.pad_token = tok.eos_token
        special = list(getattr(tok, "additional_special_tokens", None) or [])
        if "<audio>" not in special:
            tok.add_special_tokens({"additional_special_tokens": special + ["<audio>"]})
            self.language_model.resize_token_embeddings(len(tok), mean_resizing=True)
        tok.padding_side = "right"
        self.tokenizer = tok
        self.audio_token_id = tok.convert_tokens_to_ids("<audio>")
        for cfg in (self.config.text_config, self.language_model.config, getattr(self, "generation_config", None)):
            if cfg is not None:
                cfg.pad_token_id, cfg.eos_token_id, cfg.bos_token_id = tok.pad_token_id, tok.eos_token_id, tok.bos_token_id

    # ------------------------------------------------------------------ bookkeeping the Trainer relies on
    def train(self, mode: bool = True):
        super().train(mode)
        self.audio_tower.train(False)
        if getattr(self.config, "freeze_language_model", True):
            self.language_model.train(False)
        return self

    def _set_gradient_checkpointing(self, enable: bool = True, gradient_checkpointing_func=None):
        # activations of the decoder live in one preallocated workspace (engine.cu); there is nothing to checkpoint
        return None

    def get_input_embeddings(self):
        return self.language_model.get_input_embeddings()

    def set_input_embeddings(self, value):
        self.language_model.set_input_embeddings(value)
        self._hot = None

    def get_output_embeddings(self):
        return self.language_model.get_output_embeddings()

    def set_output_embeddings(self, value):
        self.language_model.set_output_embeddings(value)
        self._hot = None

    def get_processor(self):
        from .asr_processing import ASRProcessor
        return ASRProcessor(feature_extractor=self.feature_extractor, tokenizer=self.tokenizer, projector=self.projector,
                            encoder_conv_layers=self.config.encoder_conv_layers)

    def state_dict(self, *args, **kwargs):
        """Trainable weights only (projector), reference key names (`projector.linear_1.weight`, ...).  LoRA adapters are
        serialised separately (peft layout, `lora_adapters.peft_state_dict()`), as in the reference (asr_modeling.py:398-422)."""
        sd = {f"projector.{k}": v for k, v in self.projector.state_dict().items()}
        if not getattr(self.config, "freeze_language_model", True):      # fine-tuned decoder travels with the checkpoint (:409-422)
            sd.update({f"language_model.{k}": v for k, v in self.language_model.state_dict().items()
                       if not k.startswith("lora_adapters.")})
        return sd

    def _compute_encoder_output_lengths(self, audio_attention_mask: torch.Tensor) -> torch.Tensor:
        return compute_encoder_output_length(audio_attention_mask.sum(dim=-1), self.config.encoder_conv_layers)

    def _get_num_audio_tokens(self, audio_attention_mask: torch.Tensor) -> int:
        n = int(self._compute_encoder_output_lengths(audio_attention_mask).max().item())
        return int(self.projector.get_output_length(n))

    # ------------------------------------------------------------------ the B200 hot path
    def path_dims(self) -> PathDims:
        ec, tc = self.audio_tower.config, self.language_model.config
        rope_e = getattr(ec, "rope_parameters", None) or {}
        rope_t = getattr(tc, "rope_parameters", None) or {}
        emb = self.language_model.get_input_embeddings().weight
        hidden = getattr(self.config, "projector_hidden_dim", None) or tc.hidden_size
        return PathDims(
            n_mels=ec.num_mel_bins, enc_dim=ec.hidden_size, enc_ffn=ec.intermediate_size, enc_layers=ec.num_hidden_layers,
            enc_heads=ec.num_attention_heads, enc_rope_theta=float(rope_e.get("rope_theta", 10000.0)),
            enc_partial_rotary=float(rope_e.get("partial_rotary_factor", getattr(ec, "partial_rotary_factor", 0.5))),
            enc_max_pos=max(int(getattr(ec, "max_position_embeddings", 1500)), 1500),
            proj_k=self.config.projector_pool_stride, proj_hidden=hidden,
            lm_dim=tc.hidden_size, lm_ffn=tc.intermediate_size, lm_layers=tc.num_hidden_layers,
            lm_heads=tc.num_attention_heads, lm_kv_heads=tc.num_key_value_heads,
            lm_head_dim=getattr(tc, "head_dim", None) or tc.hidden_size // tc.num_attention_heads,
            lm_rope_theta=float(rope_t.get("rope_theta", getattr(tc, "rope_theta", 1e6))), lm_eps=tc.rms_norm_eps,
            vocab=emb.shape[0], audio_token_id=int(self.audio_token_id))

    def _frozen_versions(self) -> int:
        """Sum of the in-place version counters of the frozen towers' tensors: an in-place weight load (load_state_dict, a
        copy_ into a parameter) after the operands were packed must not go unnoticed."""
        ts = self.__dict__.get("_frozen_tensors")
        if ts is None:
            # parameters only: buffers (rotary inv_freq) are not used by the engine, and a DDP wrapper re-broadcasts them every step
            ts = list(self.audio_tower.parameters())
            if getattr(self.config, "freeze_language_model", True):
                ts += [p for n, p in self.language_model.named_parameters() if not n.startswith("lora_adapters.")]
            self.__dict__["_frozen_tensors"] = ts
        return sum(int(t._version) for t in ts)

    def _hot_path(self) -> HotPath:
        dev = next(self.projector.parameters()).device
        if dev.type != "cuda":
            raise L.TinyAudioB200Error("ASRModel.forward needs the model on a CUDA device (no CPU fallback)")
        emb = self.language_model.get_input_embeddings().weight
        key = (dev.index, emb.data_ptr(), self._frozen_versions())
        if self._hot is None or self._hot_key != key:
            self.__dict__.pop("_frozen_tensors", None)
            with torch.no_grad():
                lm_sd = {k: v for k, v in self.language_model.state_dict().items() if not k.startswith("lora_adapters.")}
                self._hot = HotPath(self.path_dims(), self.audio_tower.state_dict(), lm_sd, dev, lora=self.lora_adapters is not None)
            self._hot_key = (dev.index, emb.data_ptr(), self._frozen_versions())
            self._lm_pack_version = None
        return self._hot

    def _decoder_tensors(self):
        """The decoder's trainable tensors the CUDA engine produces gradients for: stacked LoRA A then B tensors, or every Qwen3
        parameter when the decoder is unfrozen, or nothing."""
        if self.lora_adapters is not None:
            la, lb = self.lora_adapters.tensors()
            return tuple(la.values()) + tuple(lb.values())
        if not getattr(self.config, "freeze_language_model", True):
            return tuple(p for _, p in self.language_model.named_parameters())
        return ()

    def _refresh_packed_lm(self, hot: HotPath) -> None:
        """Unfrozen decoder: re-pack the engine's bf16 operands when the fp32 masters moved (optimiser step, in-place load)."""
        named = list(self.language_model.named_parameters())
        version = sum(int(t._version) for _, t in named)
        if getattr(self, "_lm_pack_version", None) != version:
            hot.lm.refresh_from({n: t.detach() for n, t in named})
            self._lm_pack_version = version

    def _encode_audio(self, audio_features: torch.Tensor, expected_token_counts: torch.Tensor) -> torch.Tensor:
        """Reference signature (asr_modeling.py:434-456): packed audio embeddings (sum(counts), llm_dim)."""
        hot = self._hot_path()
        B = audio_features.shape[0]
        if audio_features.dim() == 2:
            im2, _, T = hot.logmel(audio_features.float().contiguous())
        else:
            im2, T = hot.mel_to_im2col(audio_features)
        enc = hot.encode(im2, B, T).clone()
        enc = self._maybe_drop_audio_tokens(enc)
        audio = self.projector(enc)
        return _gather_audio_embeds(audio, expected_token_counts.to(device=audio.device, dtype=torch.long))

    def _maybe_drop_audio_tokens(self, hidden_states: torch.Tensor) -> torch.Tensor:
        p = float(getattr(self.config, "audio_token_dropout", 0.0))
        if not self.training or p <= 0.0:
            return hidden_states
        keep = torch.bernoulli(torch.full(hidden_states.shape[:-1], 1.0 - p, device=hidden_states.device,
                                          dtype=hidden_states.dtype)).unsqueeze(-1)
        return hidden_states * keep

    def forward(self, input_ids: Optional[torch.Tensor] = None, input_features: Optional[torch.Tensor] = None,
                audio_attention_mask: Optional[torch.Tensor] = None, attention_mask: Optional[torch.Tensor] = None,
                position_ids: Optional[torch.Tensor] = None, past_key_values=None, inputs_embeds: Optional[torch.Tensor] = None,
                labels: Optional[torch.Tensor] = None, use_cache: Optional[bool] = None,
                cache_position: Optional[torch.Tensor] = None, audio_token_counts: Optional[torch.Tensor] = None,
                **kwargs) -> CausalLMOutputWithPast:
        """Training / scoring forward (reference: asr_modeling.py:481-533).

        `input_features` is either the reference's (B, n_mels, T) log-mel tensor or, on the fast path, the zero-padded 16 kHz
        waveform (B, L) -- the log-mel then runs on the GPU (ta_logmel_fwd).  Without `input_features` the batch is text-only
        (or `inputs_embeds`) and only the CUDA decoder runs.  `labels` may live on the host or on the device.

        `outputs.logits`: the reference always returns the decoder's logits for every position.  Here they are produced when
        there are no labels (scoring / text-only calls), or on request (`return_logits=True`, or `config.return_logits`); a
        training call with labels returns `logits=None` by default, because the fused path evaluates the lm_head only on the
        labelled rows (that is what keeps 9 GB of logits out of HBM).  Logits are bf16 [B, S, vocab], the dtype the reference
        produces under its bf16-autocast recipe.

        `attention_mask`: right padding (what the tokenizer this model configures produces, padding_side = "right") needs no
        mask under causal attention -- padded positions only influence themselves and carry label -100.  Left-padded batches
        are served by generate()."""
        if past_key_values is not None:
            raise NotImplementedError("past_key_values: cached decoding goes through ASRModel.generate() on the B200 path")
        if input_ids is None and inputs_embeds is None:
            raise ValueError("You must specify exactly one of input_ids or inputs_embeds")
        hot = self._hot_path()
        dev = hot.device
        return_logits = kwargs.pop("return_logits", None)
        if return_logits is None:
            return_logits = bool(getattr(self.config, "return_logits", False)) or labels is None
        nib = kwargs.get("num_items_in_batch")
        call = dict(labels=labels, want_hidden=bool(return_logits))
        if nib is not None:
            call["num_items_in_batch"] = float(nib)
        extra = self._decoder_tensors()
        with_audio = input_features is not None and input_ids is not None and inputs_embeds is None
        if labels is None and torch.is_grad_enabled():      # nothing to differentiate: a scoring call must not run the backward
            with torch.no_grad():
                return self.forward(input_ids=input_ids, input_features=input_features, audio_attention_mask=audio_attention_mask,
                                    attention_mask=attention_mask, position_ids=position_ids, inputs_embeds=inputs_embeds,
                                    audio_token_counts=audio_token_counts, return_logits=return_logits, **kwargs)
        if not with_audio:
            # text-only / inputs_embeds forward (asr_modeling.py:496-497, 517-526): the CUDA decoder alone
            if inputs_embeds is not None:
                call["inputs_embeds"] = inputs_embeds
                B, S = int(inputs_embeds.shape[0]), int(inputs_embeds.shape[1])
            else:
                B, S = input_ids.shape
            if input_ids is not None:
                call["input_ids"] = input_ids.to(dev, non_blocking=True)
            loss = _LmLossFn.apply(self, call, None, *extra)
        else:
            B, S = input_ids.shape
            feats = input_features.to(dev, non_blocking=True)
            call.update(input_ids=input_ids.to(dev, non_blocking=True),
                        audio_token_counts=(audio_token_counts.to(dev, non_blocking=True) if audio_token_counts is not None else None))
            audio_kw = dict(waveform=feats.float().contiguous()) if feats.dim() == 2 else dict(input_features=feats)
            p = float(getattr(self.config, "audio_token_dropout", 0.0))
            if self.training and p > 0.0:
                audio_kw["frame_keep_prob"] = 1.0 - p
            keep_mask = kwargs.pop("audio_frame_keep_mask", None)     # parity hook: replay a given Bernoulli draw (tests)
            if keep_mask is not None:
                audio_kw["frame_keep_mask"] = keep_mask
            pr = self.projector
            from .projectors import MLPAudioProjector
            if isinstance(pr, MLPAudioProjector):
                # fully fused path: projector forward/backward run inside the CUDA engine together with the towers
                call.update(audio_kw)
                loss = _FusedPathLoss.apply(self, call, pr.linear_1.weight, pr.norm.weight, pr.linear_2.weight, pr.norm_2.weight, *extra)
            else:
                # generic projector (qformer / mosa / moe): frozen encoder -> projector module (autograd) -> CUDA decoder + CE with
                # d(loss)/d(audio embeddings) handed back to autograd
                enc = hot.encode_audio(**audio_kw).clone()
                audio = pr(enc)
                loss = _LmLossFn.apply(self, call, audio.float(), *extra)
                if labels is not None and hasattr(pr, "get_aux_loss"):       # MoE load-balance + z-loss (asr_modeling.py:528-531)
                    aux = pr.get_aux_loss()
                    if aux is not None and aux.numel() > 0:
                        loss = loss + aux.to(loss.device)
        logits = hot.logits_all(B, S) if return_logits else None     # from the final hidden states the engine kept (want_hidden)
        if labels is None:
            loss = None
        return CausalLMOutputWithPast(loss=loss, logits=logits)

    @torch.no_grad()
    def generate(self, input_ids: Optional[torch.Tensor] = None, input_features: Optional[torch.Tensor] = None,
                 audio_attention_mask: Optional[torch.Tensor] = None, attention_mask: Optional[torch.Tensor] = None,
                 audio_token_counts: Optional[torch.Tensor] = None, max_new_tokens: Optional[int] = None, **kwargs) -> torch.Tensor:
        """Greedy transcription ids (reference: asr_modeling.py:562-646 with num_beams=1, do_sample=False).  Returns only the
        newly generated tokens, like the reference (it strips the prompt, :644-646).  `input_ids` holds the prompt with its <audio>
        placeholders; when it is None the prompt is built from the tokenizer's chat template exactly as the reference does
        (:588-617; needs `audio_attention_mask` to size the placeholder run).  Ragged batches: `audio_attention_mask` gives every clip
        its own audio token count and `attention_mask` marks LEFT-padded prompts (rotary positions start at each sequence's first real
        token, padding keys are masked) -- ids equal to the reference's HF generate (tests/golden/generate_ragged.npz)."""
        if input_features is None:
            raise ValueError("input_features required for generation")
        if input_ids is None:
            # build the prompt like the reference (asr_modeling.py:588-617): N_a <audio> placeholders + the transcribe instruction
            # through the tokenizer's chat template (thinking mode off), one prompt shared by the whole batch
            if audio_attention_mask is None:
                raise ValueError("audio_attention_mask required for generation")
            if not hasattr(self.tokenizer, "apply_chat_template"):
                raise NotImplementedError("generate() without input_ids needs a tokenizer with a chat template")
            n_audio = self._get_num_audio_tokens(audio_attention_mask)
            system_prompt = kwargs.pop("system_prompt", None) or self.system_prompt
            messages = []
            if system_prompt:
                messages.append({"role": "system", "content": system_prompt})
            content = "<audio>" * n_audio
            if self.TRANSCRIBE_PROMPT:
                content += " " + self.TRANSCRIBE_PROMPT
            messages.append({"role": "user", "content": content})
            chat = self.tokenizer.apply_chat_template(messages, tokenize=True, add_generation_prompt=True, return_tensors="pt",
                                                      enable_thinking=False)
            input_ids = chat.input_ids if hasattr(chat, "input_ids") else chat
            input_ids = torch.as_tensor(input_ids)
            if input_ids.dim() == 1:
                input_ids = input_ids.unsqueeze(0)
            if input_ids.shape[0] == 1 and input_features.shape[0] > 1:
                input_ids = input_ids.expand(input_features.shape[0], -1)
        kwargs.pop("system_prompt", None)
        if kwargs.get("num_beams", 1) != 1 or kwargs.get("do_sample", False):
            raise NotImplementedError("only greedy decoding is implemented on the B200 path")
        # the reference's defaults are neutral (asr_config.py:103-111); a logits processor that would change the greedy choice must not
        # be dropped silently
        gcfg = self.generation_config
        for name, neutral in (("repetition_penalty", 1.0), ("no_repeat_ngram_size", 0), ("min_new_tokens", 0), ("num_beams", 1),
                              ("do_sample", False)):
            value = kwargs.get(name, getattr(gcfg, name, neutral))
            if value not in (None, neutral):
                raise NotImplementedError(f"generate(): {name}={value!r} is not supported on the B200 path (greedy decoding without "
                                          f"logits processors only; the reference's default is {neutral!r})")
        hot = self._hot_path()
        feats = input_features.to(hot.device)
        pr = self.projector
        kw = dict(waveform=feats.float().contiguous()) if feats.dim() == 2 else dict(input_features=feats)
        from .projectors import MLPAudioProjector
        if isinstance(pr, MLPAudioProjector):
            params = {k: p.detach().float().contiguous() for k, p in zip(_PROJ_KEYS, (pr.linear_1.weight, pr.norm.weight,
                                                                                   pr.linear_2.weight, pr.norm_2.weight))}
        else:       # generic projector (qformer): CUDA encoder -> projector module -> CUDA decoder
            params = None
            kw = dict(audio_embeds=pr(hot.encode_audio(**kw).clone()).float())
        gc = self.generation_config
        eos = gc.eos_token_id if isinstance(gc.eos_token_id, (list, tuple)) else [gc.eos_token_id]
        _prepare_decoder(self, hot, self._decoder_tensors())      # current LoRA operands / re-packed decoder after optimiser steps
        use_cache = kwargs.get("use_cache")
        kw["use_cache"] = bool(getattr(self.config, "use_cache", True) if use_cache is None else use_cache)
        if audio_token_counts is None and audio_attention_mask is not None:
            # per-sample audio token counts from the frame mask, as the reference derives them (asr_modeling.py:587-589): clips of
            # different lengths in one batch place different numbers of audio embeddings
            enc_len = self._compute_encoder_output_lengths(audio_attention_mask)
            audio_token_counts = self.projector.get_output_length(enc_len).to(torch.long)
        kw["attention_mask"] = attention_mask      # left-padded prompts of a ragged batch (HF generate semantics)
        return hot.greedy_generate(input_ids=input_ids, proj_params=params, audio_token_counts=audio_token_counts,
                                   max_new_tokens=int(max_new_tokens or gc.max_new_tokens or 128),
                                   eos_token_ids=[e for e in eos if e is not None], pad_token_id=int(gc.pad_token_id or 0), **kw)

    # ------------------------------------------------------------------ persistence (reference checkpoint layout)
    def save_pretrained(self, save_directory, **kwargs):
        """Checkpoint in the reference's on-disk layout (asr_modeling.py:769-852): `config.json`, `model.safetensors` holding
        the overridden state_dict (trainable weights only, `projector.*` keys), tokenizer + feature-extractor files,
        `preprocessor_config.json` with the ASRProcessor auto_map, and -- with LoRA attached -- peft's
        `adapter_model.safetensors` / `adapter_config.json` (key names and fields as peft 0.19 writes them, so the files load
        with `PeftModel.from_pretrained` in the stock stack).  The reference also copies its own python sources next to the
        weights for trust_remote_code loading; this package's sources need the CUDA library and are not copied."""
        from safetensors.torch import save_file
        out = Path(save_directory)
        out.mkdir(parents=True, exist_ok=True)
        self.config.vocab_size = self.language_model.config.vocab_size
        if getattr(self.config, "text_config", None) is not None:
            self.config.text_config.vocab_size = self.language_model.config.vocab_size
        if hasattr(self.audio_tower.config, "num_mel_bins"):
            self.config.audio_config.num_mel_bins = self.audio_tower.config.num_mel_bins
        try:
            self.config.save_pretrained(out)
        except OSError:      # offline: the diff against a default ASRConfig() needs the hub (tower configs) -- write the full dict
            self.config.to_json_file(str(out / "config.json"), use_diff=False)
        tensors, seen = {}, {}
        for k, v in self.state_dict().items():          # tied tensors (lm_head / embed_tokens) are written once, as HF's save_pretrained does
            key = (v.data_ptr(), tuple(v.shape), tuple(v.stride()))
            if key in seen and v.numel() > 0:
                continue
            seen[key] = k
            tensors[k] = v.detach().cpu().contiguous().clone()
        save_file(tensors, str(out / "model.safetensors"), metadata={"format": "pt"})
        for obj in (getattr(self, "tokenizer", None), getattr(self, "feature_extractor", None)):
            if obj is not None and hasattr(obj, "save_pretrained"):
                obj.save_pretrained(out)
        adapters = self.lora_adapters
        if adapters is not None:
            save_file({k: v.cpu().contiguous() for k, v in adapters.peft_state_dict().items()}, str(out / "adapter_model.safetensors"),
                      metadata={"format": "pt"})
            repo_id = (kwargs.get("repo_id") or kwargs.get("push_to_hub_model_id")
                       or getattr(self.config, "pretrained_model_path", None) or "")
            (out / "adapter_config.json").write_text(json.dumps({
                "peft_type": "LORA", "task_type": "CAUSAL_LM", "base_model_name_or_path": repo_id, "r": adapters.rank,
                "lora_alpha": adapters.alpha, "lora_dropout": 0.0, "bias": "none", "target_modules": list(adapters.targets),
                "fan_in_fan_out": False, "inference_mode": False, "init_lora_weights": True, "modules_to_save": None,
                "use_rslora": False, "use_dora": False}, indent=2))
        pre = out / "preprocessor_config.json"
        pc = json.loads(pre.read_text()) if pre.exists() else {}
        pc.update({"processor_class": "ASRProcessor", "auto_map": {"AutoProcessor": "asr_processing.ASRProcessor"}})
        pre.write_text(json.dumps(pc, indent=2))

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, *args, **kwargs):
        """Rebuild from a checkpoint written by this class or by the reference (asr_modeling.py:59-131): towers come from the
        loader seams (config.audio_model_id / text_model_id), then `model.safetensors` -- resolved like the reference does, through
        transformers' `cached_file` (local directory or hub repo id, `subfolder` / `revision` honoured) -- is overlaid with
        `load_state_dict(strict=False)`: the projector, and the fine-tuned decoder (`language_model.*`) when the checkpoint was
        trained with freeze_language_model=False.  With config.use_lora the adapters are attached afterwards and
        `adapter_model.safetensors` is loaded when `adapter_config.json` exists (fresh adapters otherwise)."""
        from safetensors.torch import load_file
        from transformers.utils.hub import cached_file
        config = kwargs.pop("config", None)
        name = str(pretrained_model_name_or_path)
        if config is None:
            config = ASRConfig.from_pretrained(name, **kwargs)
        cache_kwargs = {k: kwargs[k] for k in ("subfolder", "revision") if kwargs.get(k)}

        def resolve(fname):
            return cached_file(name, fname, _raise_exceptions_for_missing_entries=False, **cache_kwargs)

        cls._is_loading_from_pretrained = True
        try:
            model = cls(config)
            wfile = resolve("model.safetensors")
            if wfile is not None:
                sd = load_file(wfile)
                result = model.load_state_dict(sd, strict=False)
                lost = [k for k in result.unexpected_keys]
                if lost:
                    raise RuntimeError(f"checkpoint {wfile} holds weights this model has no place for: {lost[:8]}")
                want = [f"projector.{k}" for k in model.projector.state_dict()]
                absent = [k for k in want if k not in sd]
                if absent:
                    raise RuntimeError(f"checkpoint {wfile} lacks projector weights {absent[:8]}")
                model._hot = None                   # packed operands (if any) are stale now
                model._lm_pack_version = None
            else:       # the reference proceeds silently with a freshly initialised projector; say so
                import warnings
                warnings.warn(f"{name}: no model.safetensors found -- the projector keeps its random initialisation", stacklevel=2)
            if getattr(config, "use_lora", False):       # adapters are attached after the base weights, as in the reference
                model._setup_lora(config)
                if resolve("adapter_config.json") is not None:
                    afile = resolve("adapter_model.safetensors")
                    if afile is None:
                        raise RuntimeError(f"{name}: adapter_config.json without adapter_model.safetensors")
                    model.lora_adapters.load_peft_state_dict(load_file(afile))
            return model
        finally:
            cls._is_loading_from_pretrained = False

    def push_to_hub(self, repo_id: str, **kwargs):
        """Reference behaviour (asr_modeling.py:854-865): remember the repo id so that save_pretrained writes it into
        adapter_config.json (`base_model_name_or_path`), then defer to the stock uploader."""
        self.config.pretrained_model_path = repo_id
        return super().push_to_hub(repo_id, **kwargs)

    def load_projector(self, path: str):
        from safetensors.torch import load_file
        sd = load_file(str(Path(path) / "model.safetensors"))
        self.projector.load_state_dict({k[len("projector."):]: v for k, v in sd.items() if k.startswith("projector.")})

try:
    import transformers
    transformers.AutoModel.register(ASRConfig, ASRModel, exist_ok=True)
except Exception:   # registration is a convenience, never a requirement of the hot path
    pass
