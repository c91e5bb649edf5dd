"""Integer golden vectors from the UNMODIFIED reference (runs only where /root/reference exists): the token-count arithmetic the
collator and the processor rely on (scripts/train.py:335-338, tiny_audio/asr_processing.py:95-110) for every registered projector,
the ragged gather of audio embeddings (tiny_audio/asr_modeling.py:27-44), and what ASRProcessor.__call__ builds (chat messages,
ids, masks; asr_processing.py:51-128).  Output: tests/golden/integer_semantics.npz, tests/golden/processor_calls.json

usage:  python oracle/make_integer_golden.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import load_reference  # noqa: E402


class Cfg:
    encoder_dim, llm_dim, projector_pool_stride, projector_hidden_dim = 64, 128, 4, None
    num_experts, num_experts_per_tok, router_aux_loss_coef = 4, 2, 0.01
    qformer_window_size, downsample_rate, qformer_hidden_size, qformer_num_layers, qformer_num_heads = 15, 5, None, 1, 4
    qformer_intermediate_size = None


class RecordingTokenizer:
    """Stands in for the Qwen3 tokenizer (not available offline): records what apply_chat_template receives and returns ids that are a
    deterministic function of the messages (one 777 per <audio> placeholder), so two processors agree iff they build the same chat."""

    def __init__(self):
        self.calls = []

    def convert_tokens_to_ids(self, t):
        return 777 if t == "<audio>" else None

    def apply_chat_template(self, messages, tokenize=True, add_generation_prompt=False, return_tensors=None, enable_thinking=None):
        self.calls.append(dict(messages=messages, add_generation_prompt=add_generation_prompt, enable_thinking=enable_thinking,
                               tokenize=tokenize))
        ids = []
        for m in messages:
            rest = m["content"].replace("<audio>", "")
            ids += [1 + len(m["role"])] + [777] * m["content"].count("<audio>") + [10 + sum(map(ord, rest)) % 1000]
        if add_generation_prompt:
            ids.append(5)
        return torch.tensor([ids])


class StackProjector:
    def get_output_length(self, n):
        return (n - 4) // 4 + 1


PROCESSOR_CASES = ((16000, None, None), (12345, "hello world", None), (48000, None, "You are helpful."), (None, None, None),
                   (8000, "x", "sys"))


def run_processor_cases(processor_cls):
    """ASRProcessor.__call__ (tiny_audio/asr_processing.py:51-128) on audio only / audio + text / system prompt / text only."""
    from transformers import WhisperFeatureExtractor
    rng = np.random.default_rng(3)
    out = []
    for audio_len, text, system_prompt in PROCESSOR_CASES:
        tok = RecordingTokenizer()
        proc = processor_cls(WhisperFeatureExtractor(feature_size=128), tok, projector=StackProjector())
        audio = rng.standard_normal(audio_len).astype(np.float32) if audio_len else None
        res = proc(audio=audio, text=text, system_prompt=system_prompt)
        rec = {"calls": tok.calls, "input_ids": res["input_ids"].tolist(), "attention_mask": res["attention_mask"].tolist(),
               "keys": sorted(res.keys())}
        if audio_len:
            rec["audio_mask_sum"] = int(res["audio_attention_mask"].sum())
            rec["feat_shape"] = list(res["input_features"].shape)
        out.append(rec)
    return out


def config_dicts(config_cls):
    """ASRConfig(...).to_dict() for the default recipe and for two overridden ones (asr_config.py:36-220), sub-configs built offline."""
    from transformers import GlmAsrEncoderConfig, Qwen3Config
    out = {}
    for name, kw in (("default", {}), ("qformer_lora", dict(projector_type="qformer", use_lora=True, lora_rank=16, max_new_tokens=64,
                                                           freeze_projector=True)),
                     ("unfrozen_moe", dict(projector_type="moe", freeze_language_model=False, num_experts=8, projector_hidden_dim=2048,
                                           audio_token_dropout=0.05, system_prompt=""))):
        d = config_cls(audio_config=GlmAsrEncoderConfig(), text_config=Qwen3Config(), **kw).to_dict()
        d.pop("transformers_version", None)
        out[name] = d
    return out


def model_surface(build):
    """What scripts/train.py and the HF Trainer see of the model: `state_dict()` keys in order (checkpoint layout,
    tiny_audio/asr_modeling.py:398-422), names of the trainable parameters (optimizer groups, train.py:384-437) and class-level
    attributes, for every projector type and for the unfrozen-decoder recipe.  `build(kind, freeze_lm)` -> model."""
    out = {}
    for kind, freeze_lm in (("mlp", True), ("qformer", True), ("mosa", True), ("moe", True), ("mlp", False)):
        m = build(kind, freeze_lm)
        out[f"{kind}{'' if freeze_lm else '+unfrozen_lm'}"] = {
            "state_dict_keys": list(m.state_dict().keys()),
            "trainable": [n for n, p in m.named_parameters() if p.requires_grad],
            "n_parameters": sum(p.numel() for p in m.parameters()),
            "class_attrs": {a: getattr(type(m), a, None) for a in ("base_model_prefix", "main_input_name", "_supports_flash_attn_2",
                                                                      "supports_gradient_checkpointing", "TRANSCRIBE_PROMPT")},
            "training_flags_after_train()": [m.train().audio_tower.training, m.language_model.training, m.projector.training],
        }
    return out


def _reference_builder(mods):
    from oracle import path_oracle as po
    from oracle.make_golden import PROJECTOR_CONFIG_EXTRAS, PROJECTOR_INIT, build_reference_model
    cfg = po.small_config(enc_layers=1, lm_layers=1)

    def build(kind, freeze_lm):
        W = po.init_weights(cfg, seed=3)
        if kind in PROJECTOR_INIT:
            W["projector"] = PROJECTOR_INIT[kind](cfg, seed=5)
        return build_reference_model(cfg, W, mods, kind, freeze_lm=freeze_lm, **PROJECTOR_CONFIG_EXTRAS.get(kind, {}))
    return build


GENERATE_CASE = dict(seed=31, batch=2, clip_s=2.0, response_len=4, new_tokens=8)


def generate_inputs():
    """Seeded small model + batch whose prompt ends with the assistant header (everything before the first label)."""
    from oracle import path_oracle as po
    c = GENERATE_CASE
    cfg = po.small_config(enc_layers=2, lm_layers=2)
    W = po.init_weights(cfg, seed=c["seed"])
    batch = po.synthetic_batch(cfg, c["batch"], c["clip_s"], seed=c["seed"], response_len=c["response_len"])
    first = int((batch["labels"][0] != -100).nonzero()[0])
    return cfg, W, batch, batch["input_ids"][:, :first].contiguous()


def reference_generate(mods):
    """Greedy ids from the reference's OWN ASRModel.generate (tiny_audio/asr_modeling.py:562-646 -> HF generate, greedy defaults)."""
    from oracle.make_golden import build_reference_model
    cfg, W, batch, prompt = generate_inputs()
    ref = build_reference_model(cfg, W, mods, "mlp")
    ref.eval()
    L = int(batch["sample_lengths"][0])
    feats = ref.feature_extractor([batch["waveform"][b, :L].numpy() for b in range(prompt.shape[0])], sampling_rate=16000,
                                  padding="longest", return_attention_mask=True, return_tensors="pt")
    out = ref.generate(input_ids=prompt, input_features=feats.input_features, audio_attention_mask=feats.attention_mask,
                       attention_mask=torch.ones_like(prompt), max_new_tokens=GENERATE_CASE["new_tokens"])
    return prompt.numpy(), out.numpy()


def main():
    mods = load_reference()
    import json
    prompt, ids = reference_generate(mods)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "generate_ids.npz"), prompt=prompt, ids=ids)
    with open(os.path.join(ROOT, "tests", "golden", "model_surface.json"), "w") as f:
        json.dump(model_surface(_reference_builder(mods)), f, indent=1)
    with open(os.path.join(ROOT, "tests", "golden", "asr_config_dicts.json"), "w") as f:
        json.dump(config_dicts(mods["asr_config"].ASRConfig), f, indent=1, sort_keys=True, default=str)
    with open(os.path.join(ROOT, "tests", "golden", "processor_calls.json"), "w") as f:
        json.dump(run_processor_cases(mods["asr_processing"].ASRProcessor), f, indent=1)
    P, A, M = mods["projectors"], mods["asr_config"], mods["asr_modeling"]
    mel = np.arange(1, 3001, dtype=np.int64)                       # mel frames 1 .. 3000 (30 s)
    enc = np.array([int(A.compute_encoder_output_length(int(t))) for t in mel], dtype=np.int64)
    enc_t = A.compute_encoder_output_length(torch.from_numpy(mel)).numpy()
    assert np.array_equal(enc, enc_t)
    fx = {"mel_frames": mel, "encoder_frames": enc}
    torch.manual_seed(0)
    for kind in ("mlp", "mosa", "moe", "qformer"):
        proj = P.PROJECTOR_CLASSES[kind](Cfg())
        fx[f"audio_tokens.{kind}"] = np.array([int(proj.get_output_length(int(e))) for e in enc], dtype=np.int64)
    # other pool strides of the frame-stacking projectors (projector_pool_stride is a config field)
    for k in (2, 5):
        c = Cfg()
        c.projector_pool_stride = k
        fx[f"audio_tokens.mlp.k{k}"] = np.array([int(P.MLPAudioProjector(c).get_output_length(int(e))) for e in enc], dtype=np.int64)
    # ragged gather: 64 random cases, counts both below and above the available rows (zero-padding branch)
    rng = np.random.default_rng(7)
    cases = []
    for i in range(64):
        B, n, D = int(rng.integers(1, 6)), int(rng.integers(1, 12)), 3
        x = torch.from_numpy(rng.standard_normal((B, n, D)).astype(np.float32))
        counts = torch.from_numpy(rng.integers(0, n + 4, size=B).astype(np.int64))
        out = M._gather_audio_embeds(x, counts)
        cases.append((x.numpy(), counts.numpy(), out.numpy()))
    fx["gather.n_cases"] = np.array(len(cases))
    for i, (x, c, o) in enumerate(cases):
        fx[f"gather.{i}.x"], fx[f"gather.{i}.counts"], fx[f"gather.{i}.out"] = x, c, o
    out = os.path.join(ROOT, "tests", "golden", "integer_semantics.npz")
    np.savez_compressed(out, **fx)
    print("wrote", out, {k: v.shape for k, v in fx.items() if not k.startswith("gather.")})


if __name__ == "__main__":
    main()
