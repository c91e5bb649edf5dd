"""The UNMODIFIED reference `scripts/train.py` driven against this repo's plugin surface ("drops into
scripts/train.py unchanged").  hydra / omegaconf / wandb / trl / audiomentations are not installed here, so the module is imported
with stand-ins for exactly those imports (none of them is on the path under test); `tiny_audio.*` resolves to THIS repo through the
`tiny_audio/` shim, as it would with `PYTHONPATH=<this repo>:<tiny-audio checkout>`.

Runs only where the reference checkout exists (the build container); the GPU box has no /root/reference, and nothing here needs a GPU:
  * `DataCollator` (train.py:240-348) with this repo's WaveformFeatureExtractor + projector: padding mode, waveform pass-through,
    frame mask -> encoder frames -> `<audio>` placeholder count, bad-sample filtering;
  * `ASRTrainer.create_optimizer` (train.py:384-437) over this repo's ASRModel: the `language_model.` / decay parameter groups it
    builds from `named_parameters()` equal the ones it builds for the reference's own ASRModel, and ClipAdamW accepts them."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "scripts", "train.py")), reason="reference checkout not present")

AUDIO_ID = 5002


class ChatMLStandIn:
    """Stand-in for trl.experimental.utils.DataCollatorForChatML (trl is absent: its exact padding / masking is 'parity unpinned',
    SURVEY.md section 8c).  Same contract: a list of {"messages": [...]} -> input_ids / attention_mask / labels, prompt masked with
    -100, right padding.  `<audio>` maps to the audio token id; every other character to a small id."""

    def __init__(self, tokenizer=None, max_length=2048):
        self.tokenizer, self.max_length = tokenizer, max_length

    @staticmethod
    def _ids(text):
        out = []
        while text:
            if text.startswith("<audio>"):
                out.append(AUDIO_ID)
                text = text[len("<audio>"):]
            else:
                out.append(10 + (ord(text[0]) % 4000))
                text = text[1:]
        return out

    def __call__(self, examples):
        rows = []
        for ex in examples:
            prompt, resp = [], []
            for m in ex["messages"]:
                ids = [1] + self._ids(m["content"]) + [2]
                (resp if m["role"] == "assistant" else prompt).extend(ids)
            rows.append((prompt, resp))
        n = max(len(p) + len(r) for p, r in rows)
        pad = 0
        ids = torch.full((len(rows), n), pad, dtype=torch.long)
        lab = torch.full((len(rows), n), -100, dtype=torch.long)
        att = torch.zeros((len(rows), n), dtype=torch.long)
        for i, (p, r) in enumerate(rows):
            ids[i, : len(p) + len(r)] = torch.tensor(p + r)
            lab[i, len(p): len(p) + len(r)] = torch.tensor(r)
            att[i, : len(p) + len(r)] = 1
        return {"input_ids": ids, "attention_mask": att, "labels": lab}


@pytest.fixture(scope="module")
def train_py():
    saved = dict(sys.modules)
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("hydra", main=lambda **k: (lambda f: f))
    mod("wandb")
    mod("omegaconf", DictConfig=dict, OmegaConf=types.SimpleNamespace(to_container=lambda c, **k: dict(c)))
    mod("trl")
    mod("trl.experimental")
    mod("trl.experimental.utils", DataCollatorForChatML=ChatMLStandIn)
    import tiny_audio  # noqa: F401  (this repo's shim)
    mod("tiny_audio.augmentation", NoiseAugmentation=object, RIRAugmentation=object)
    spec = importlib.util.spec_from_file_location("reference_train_py", os.path.join(REF, "scripts", "train.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    yield m
    for k in list(sys.modules):
        if k not in saved:
            del sys.modules[k]


def _model(**kw):
    from tiny_audio_b200.engine import PathDims
    from tiny_audio_b200.synthetic import build_offline_model
    dims = PathDims(enc_layers=1, lm_layers=2, vocab=5003, audio_token_id=AUDIO_ID)
    return build_offline_model(dims, device="cpu", **kw)


def test_reference_imports_resolve_to_this_repo(train_py):
    import tiny_audio_b200.asr_config as cfg
    import tiny_audio_b200.asr_modeling as mdl
    assert train_py.ASRModel is mdl.ASRModel and train_py.ASRConfig is cfg.ASRConfig
    assert train_py.compute_encoder_output_length is cfg.compute_encoder_output_length


def test_reference_data_collator_with_the_waveform_feature_extractor(train_py):
    m = _model()
    col = train_py.DataCollator(tokenizer=m.tokenizer, feature_extractor=m.feature_extractor, sample_rate=16000, system_prompt=None,
                                projector=m.projector, encoder_conv_layers=m.config.encoder_conv_layers)
    # the collator keys its padding mode on the extractor's class NAME (train.py:260-264): ours must read "WhisperFeatureExtractor"
    assert col._audio_padding == "max_length"
    rng = np.random.default_rng(0)
    lens = [16000, 40000, 479999, 480000]
    feats = [{"audio": {"array": 0.1 * rng.standard_normal(n).astype(np.float32)}, "text": f"hello world {i}"} for i, n in enumerate(lens)]
    feats += [{"audio": {"array": np.zeros(0, np.float32)}, "text": "empty audio"},                     # dropped: empty
              {"audio": {"array": np.full(1000, np.nan, np.float32)}, "text": "nan audio"},              # dropped: non-finite
              {"audio": {"array": np.zeros(16000 * 31, np.float32)}, "text": "too long"},                # dropped: > 30 s
              {"audio": {"array": np.zeros(16000, np.float32)}, "text": ""}]                             # dropped: empty label
    batch = col(feats)
    assert set(batch) >= {"input_ids", "attention_mask", "labels", "input_features", "audio_attention_mask", "audio_token_counts"}
    B = len(lens)
    wave, mask = batch["input_features"], batch["audio_attention_mask"]
    assert tuple(wave.shape) == (B, 480000) and wave.dtype == torch.float32          # zero-padded waveform: the mel runs on the GPU
    assert tuple(mask.shape) == (B, 3000)
    for i, n in enumerate(lens):
        assert float(wave[i, n:].abs().sum()) == 0.0 and float(wave[i, :n].abs().sum()) > 0
        # the frame mask the reference's extractor would return (HF:whisper/feature_extraction_whisper.py:328-337): every hop-th sample mask
        assert int(mask[i].sum()) == (n + 159) // 160 if n < 480000 else 3000
        mel = int(mask[i].sum())
        enc = int(train_py.compute_encoder_output_length(torch.tensor(mel), m.config.encoder_conv_layers))
        want = int(m.projector.get_output_length(enc))
        assert int(batch["audio_token_counts"][i]) == want == int((batch["input_ids"][i] == AUDIO_ID).sum())
    assert batch["audio_token_counts"].dtype == torch.long and int(batch["audio_token_counts"][-1]) == 375
    assert (batch["labels"][batch["attention_mask"] == 0] == -100).all()


def test_reference_trainer_create_optimizer_groups(train_py):
    """ASRTrainer.create_optimizer (unmodified) over this repo's ASRModel with an unfrozen decoder: four groups keyed on the
    `language_model.` prefix and on decay membership -- identical (by parameter name) to what it builds for the reference's own
    ASRModel of the same architecture -- and ClipAdamW takes the groups as torch.optim.AdamW would."""
    from transformers import Trainer
    from oracle import path_oracle as po
    from oracle.make_golden import build_reference_model, load_reference
    from tiny_audio_b200 import lib, optim

    def groups_of(model, opt_cls):
        fake = types.SimpleNamespace(model=model, optimizer=None, decoder_learning_rate=2e-5, decoder_weight_decay=0.01,
                                     projector_weight_decay=None, args=types.SimpleNamespace(learning_rate=1e-3, weight_decay=0.0))
        real = Trainer.get_optimizer_cls_and_kwargs
        Trainer.get_optimizer_cls_and_kwargs = staticmethod(lambda args, model=None: (opt_cls, dict(betas=(0.9, 0.999), eps=1e-8)))
        try:
            opt = train_py.ASRTrainer.create_optimizer(fake)
        finally:
            Trainer.get_optimizer_cls_and_kwargs = real
        names = {id(p): n for n, p in model.named_parameters()}
        return opt, [(g["lr"], g["weight_decay"], sorted(names[id(p)] for p in g["params"])) for g in opt.param_groups]

    ours = _model(freeze_language_model=False)
    real_load, real_req = lib.load, lib.require_cuda
    lib.load, lib.require_cuda = (lambda: None), (lambda *a: None)          # construction launches nothing; the CPU has no library
    try:
        opt, g_ours = groups_of(ours, optim.ClipAdamW)
    finally:
        lib.load, lib.require_cuda = real_load, real_req
    # two groups: Qwen3's RMSNorm is not an nn.LayerNorm and no trainable tensor is a bias, so the reference's own split puts every
    # trainable tensor into a "decay" group -- (projector: lr, weight_decay) and (language_model.*: decoder lr, decoder weight decay)
    assert isinstance(opt, optim.ClipAdamW) and len(g_ours) == 2
    cfg = po.small_config(enc_layers=1, lm_layers=2)
    ref = build_reference_model(cfg, po.init_weights(cfg, seed=1), load_reference(), "mlp", freeze_lm=False)
    _, g_ref = groups_of(ref, torch.optim.AdamW)
    assert [(lr, wd, n) for lr, wd, n in g_ours] == [(lr, wd, n) for lr, wd, n in g_ref]
    lrs = {round(lr, 8) for lr, _, _ in g_ours}
    assert lrs == {1e-3, 2e-5}
    dec = [n for lr, _, ns in g_ours if lr == 2e-5 for n in ns]
    assert dec and all(n.startswith("language_model.") for n in dec)
    assert sum(p.numel() for p in opt._params) == sum(p.numel() for p in ours.parameters() if p.requires_grad)
