"""The oracle against the fixtures generated from the UNMODIFIED reference (oracle/make_golden.py): this is what
pins the oracle.  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import path_oracle as po
from oracle.make_golden import CASES, PROJECTOR_INIT, case_config, sub

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load_case(name):
    spec, B, clip_s, pad_s, R, seed = CASES[name]
    cfg, kind = case_config(spec)
    fx = np.load(os.path.join(GOLD, name + ".npz"))
    W = po.init_weights(cfg, seed=seed)
    if kind in PROJECTOR_INIT:
        W["projector"] = PROJECTOR_INIT[kind](cfg, seed=seed + 1000)
    batch = po.synthetic_batch(cfg, B, clip_s, seed=seed, response_len=R, pad_to_seconds=pad_s, projector=kind)
    batch["input_ids"] = torch.from_numpy(fx["input_ids"])
    batch["labels"] = torch.from_numpy(fx["labels"])
    batch["attention_mask"] = torch.from_numpy(fx["attention_mask"])
    if "frame_keep_mask" in fx.files:       # audio_token_dropout case: replay the Bernoulli draw the reference made
        batch["frame_keep_mask"] = torch.from_numpy(fx["frame_keep_mask"])
    return cfg, fx, W, batch


@pytest.mark.parametrize("name", ["small_b2_2s", "small_b3_ragged", "h2048_b2_2s", "small_b2_1s_pad30", "qformer_b2_2s",
                                  "mosa_b2_2s", "moe_b2_2s", "dropout_b2_2s"])
def test_oracle_matches_reference_fixture(name):
    torch.set_num_threads(os.cpu_count())
    cfg, fx, W, batch = load_case(name)
    n_items = int(fx["num_items"])
    # log-mel + frame mask (a1)
    mel = po.log_mel(batch["waveform"], cfg)
    assert tuple(mel.shape) == tuple(fx["mel_shape"])
    assert np.abs(sub(mel, 8192) - fx["mel_sub"]).max() < 1e-5
    L = int(batch["sample_lengths"][0])
    mask = po.mel_attention_mask(batch["sample_lengths"], batch["waveform"].shape[1])
    assert np.array_equal(mask.numpy(), fx["mel_mask"])
    assert int(mask[0].sum()) == L // 160
    # token-count arithmetic (a2) -- integers, exact
    n_a = po.projector_output_length(po.encoder_output_length(L // 160), cfg.proj_k)
    if po.projector_kind(W["projector"]) == "qformer":
        n_a = po.qformer_output_length(po.encoder_output_length(L // 160))
    elif po.projector_kind(W["projector"]) == "mosa":
        n_a = po.mosa_output_length(po.encoder_output_length(L // 160))
    assert np.array_equal(fx["audio_token_counts"], np.full(len(fx["audio_token_counts"]), n_a))
    # forward pieces + loss + grads + optimiser (a3-a12)
    res = po.train_step(W, batch, cfg, lr=1e-3, max_grad_norm=1.0, num_items_in_batch=n_items)
    loss_mean, logits, parts = po.model_forward(W, batch, cfg, None, return_parts=True)
    assert tuple(parts["encoder_out"].shape) == tuple(fx["enc_shape"])
    assert np.abs(sub(parts["encoder_out"], 8192) - fx["enc_sub"]).max() < 2e-4
    assert np.abs(sub(parts["projector_out"], 8192) - fx["proj_sub"]).max() < 2e-4
    assert abs(float(res["loss"]) - float(fx["loss"])) < 2e-5
    assert abs(float(loss_mean) - float(fx["loss_mean_path"])) < 2e-5
    if "aux_loss" in fx.files:      # moe: load-balance + z-loss, part of the loss above (asr_modeling.py:528-531)
        assert float(fx["aux_loss"]) > 0 and abs(float(parts["aux_loss"]) - float(fx["aux_loss"])) < 1e-7
    lab_pos = torch.nn.functional.pad(batch["labels"], (0, 1), value=-100)[:, 1:] != -100
    assert np.abs(sub(logits[lab_pos], 8192) - fx["logits_lab_sub"]).max() < 2e-4
    # greedy ids: identical wherever the reference's top-1 margin exceeds fp32 noise
    top2 = logits.topk(2, -1).values
    sure = (top2[..., 0] - top2[..., 1]) > 1e-4
    assert np.array_equal(logits.argmax(-1)[sure].numpy(), fx["logits_argmax"][sure.numpy()])
    for k, g in res["grads"].items():
        ref = fx["grad_sub." + k]
        assert np.abs(sub(g) - ref).max() <= 1e-4 * np.abs(ref).max() + 1e-7
        # torch's fp32 CPU .norm() of a 5M-element tensor carries ~4e-4 relative error (the oracle sums in fp32 pairwise)
        # (gradients that are zero in exact arithmetic -- e.g. attention key biases -- are pure rounding noise: absolute floor)
        assert abs(float(g.norm()) - float(fx["grad_l2." + k])) < 1e-3 * float(fx["grad_l2." + k]) + 1e-7
        # (zero-gradient tensors move by lr * noise / (|noise| + eps): allow 1e-4 there)
        assert np.abs(sub(res["params"][k]) - fx["new_param_sub." + k]).max() < 1e-4
    assert abs(float(res["grad_norm"]) - float(fx["grad_norm"])) < 1e-3 * float(fx["grad_norm"])


def test_audio_token_dropout_mask_is_the_seeded_bernoulli_draw():
    """a4 (asr_modeling.py:458-479): the keep mask recorded from the unmodified reference (train mode, audio_token_dropout = 0.10) is
    `torch.bernoulli(full((B, S_e), 1 - p))` drawn first after `torch.manual_seed(seed)` -- the RNG contract the CUDA path keeps
    (HotPath.apply_frame_dropout draws the same way on its device) -- and it zeroes whole frames without rescaling."""
    fx = np.load(os.path.join(GOLD, "dropout_b2_2s.npz"))
    mask = torch.from_numpy(fx["frame_keep_mask"])
    p = float(fx["dropout_p"])
    assert abs(p - 0.10) < 1e-12 and set(np.unique(fx["frame_keep_mask"]).tolist()) == {0.0, 1.0}
    torch.manual_seed(int(fx["dropout_seed"]))
    assert torch.equal(mask, torch.bernoulli(torch.full(mask.shape, 1.0 - p)))
    assert tuple(mask.shape) == tuple(fx["enc_shape"][:2]) and 0 < int((mask == 0).sum()) < mask.numel() // 2


def test_oracle_unfrozen_lm_gradients_match_reference():
    """freeze_language_model: false (configs/experiments/embedded.yaml:19-33): every Qwen3 weight gradient of the oracle against
    the unmodified reference's autograd (tests/golden/unfrozen_b2_2s.npz)."""
    torch.set_num_threads(os.cpu_count())
    cfg, fx, W, batch = load_case("unfrozen_b2_2s")
    n_items = int(fx["num_items"])
    res = po.train_step(W, batch, cfg, lr=1e-3, max_grad_norm=1.0, num_items_in_batch=n_items, train_lm=True)
    assert abs(float(res["loss"]) - float(fx["loss"])) < 2e-5
    keys = [k[len("lm_grad_sub."):] for k in fx.files if k.startswith("lm_grad_sub.")]
    assert len(keys) == 2 + 11 * cfg.lm_layers and set(keys) == set(res["lm_grads"])
    for k in keys:
        g, ref = res["lm_grads"][k], fx["lm_grad_sub." + k]
        assert np.abs(sub(g, 2048) - ref).max() <= 1e-4 * np.abs(ref).max() + 1e-7, k
        assert abs(float(g.norm()) - float(fx["lm_grad_l2." + k])) < 1e-3 * float(fx["lm_grad_l2." + k]) + 1e-7, k
    for k, g in res["grads"].items():           # the projector gradients are unchanged by unfreezing the decoder
        ref = fx["grad_sub." + k]
        assert np.abs(sub(g) - ref).max() <= 1e-4 * np.abs(ref).max() + 1e-7
    # the reference clips over ALL trainable parameters
    total = (sum(float(g.double().pow(2).sum()) for g in res["lm_grads"].values())
             + sum(float(g.double().pow(2).sum()) for g in res["grads"].values())) ** 0.5
    assert abs(total - float(fx["grad_norm"])) < 1e-3 * float(fx["grad_norm"])


def test_oracle_full_size_fixture():
    """Full-depth (32 + 28 layers) model, 1 x 4 s clip: loss and logits against the reference."""
    torch.set_num_threads(os.cpu_count())
    cfg, fx, W, batch = load_case("full_b1_4s")
    loss, logits = po.model_forward(W, batch, cfg, int(fx["num_items"]))
    assert abs(float(loss) - float(fx["loss"])) < 5e-5
    lab_pos = torch.nn.functional.pad(batch["labels"], (0, 1), value=-100)[:, 1:] != -100
    assert np.abs(sub(logits[lab_pos], 8192) - fx["logits_lab_sub"]).max() < 5e-4


def test_oracle_full_size_260_token_fixture():
    """Full-depth model, 4 x 10 s clips, 260 labelled tokens (tests/golden/full_b4_10s.npz -- the fixture the GPU suite asserts the
    1e-3 loss bound on): the oracle's forward reproduces the unmodified reference's loss."""
    torch.set_num_threads(os.cpu_count())
    cfg, fx, W, batch = load_case("full_b4_10s")
    assert int(fx["num_items"]) == 260
    with torch.no_grad():
        loss, _ = po.model_forward(W, batch, cfg, int(fx["num_items"]))
    assert abs(float(loss) - float(fx["loss"])) < 5e-5


def test_mel_filter_bank_matches_hf():
    from transformers.audio_utils import mel_filter_bank
    ref = mel_filter_bank(201, 128, 0.0, 8000.0, 16000, norm="slaney", mel_scale="slaney")
    assert np.abs(po.mel_filter_bank() - ref).max() < 1e-12
    assert (np.abs(ref) > 0).sum(0).max() <= 16      # the CUDA kernel's per-filter tap budget (logmel.cu MAXW)


def test_oracle_greedy_ids_match_reference_generate():
    """The oracle's greedy decoding against ids produced by the unmodified reference's ASRModel.generate -> HF generate
    (tests/golden/generate_ids.npz): the chain CUDA path == oracle (tests/test_path_gpu.py) == reference is closed on both ends."""
    from oracle.make_integer_golden import GENERATE_CASE, generate_inputs
    torch.set_num_threads(os.cpu_count())
    fx = np.load(os.path.join(GOLD, "generate_ids.npz"))
    cfg, W, batch, prompt = generate_inputs()
    assert np.array_equal(prompt.numpy(), fx["prompt"])
    ob = dict(batch)
    ob["input_ids"] = prompt
    ids, margins = po.greedy_generate(W, ob, cfg, max_new_tokens=GENERATE_CASE["new_tokens"])
    assert float(margins.min()) > 1e-4          # every step is decisive at fp32 precision
    assert np.array_equal(ids.numpy(), fx["ids"])


def test_oracle_reproduces_full_size_reference_generate_ids():
    """tests/golden/generate_full_b8.npz (full-size model, 8 sequences x 16 free-running tokens from the UNMODIFIED reference's
    ASRModel.generate, sharpened LM with planted rows: oracle/make_generate_golden.py).  One teacher-forced fp32 oracle forward over
    prompt + reference ids reproduces every greedy decision (causal: position p's logits are step p's) with the recorded margins."""
    import torch.nn.functional as F
    from oracle.make_generate_golden import apply_planted, case_inputs
    torch.set_num_threads(os.cpu_count())
    fx = np.load(os.path.join(GOLD, "generate_full_b8.npz"))
    cfg, W, batch, prompt = case_inputs()
    assert np.array_equal(prompt.numpy(), fx["prompt"])
    apply_planted(W, fx["planted_tokens"], fx["planted_rows_bf16"])
    ids = torch.cat([prompt, torch.from_numpy(fx["ids"])], 1)
    with torch.no_grad():
        _, logits = po.model_forward(W, dict(batch, input_ids=ids, labels=None, attention_mask=None), cfg)
    S0, T = prompt.shape[1], fx["ids"].shape[1]
    step_logits = logits[:, S0 - 1: S0 - 1 + T]
    assert np.array_equal(step_logits.argmax(-1).numpy(), fx["ids"])
    top2 = step_logits.topk(2, -1).values
    assert np.abs((top2[..., 0] - top2[..., 1]).numpy() - fx["margins"]).max() < 1e-3 and float(fx["margins"].min()) >= 0.5
    assert bool(fx["ids_bf16_autocast_equal"])          # the reference's own bf16-autocast run produced the same ids


def test_oracle_reproduces_ragged_reference_generate_ids():
    """tests/golden/generate_ragged.npz: four clips of 1 / 2 / 1.5 / 2 s in one batch, LEFT-padded prompts, ids from the unmodified
    reference's ASRModel.generate -> HF generate (oracle/make_ragged_generate_golden.py).  The oracle's restatement of HF's left-padding
    semantics (rotary positions = cumsum(mask) - 1, padding keys masked) reproduces them, free-running."""
    from oracle.make_generate_golden import apply_planted
    from oracle.make_ragged_generate_golden import NEW_TOKENS, case_inputs
    fx = np.load(os.path.join(GOLD, "generate_ragged.npz"))
    cfg, W, batch = case_inputs(int(fx["seed"]))
    assert np.array_equal(batch["input_ids"].numpy(), fx["input_ids"]) and np.array_equal(batch["attention_mask"].numpy(), fx["attention_mask"])
    assert np.array_equal(batch["audio_token_counts"].numpy(), fx["audio_token_counts"]) and len(set(fx["audio_token_counts"].tolist())) >= 3
    assert (fx["attention_mask"] == 0).any() and not (np.diff(fx["attention_mask"], axis=1) < 0).any()      # ragged, left-padded
    apply_planted(W, fx["planted_tokens"], fx["planted_rows_bf16"])
    ids, margins = po.greedy_generate(W, batch, cfg, max_new_tokens=NEW_TOKENS, attention_mask=batch["attention_mask"])
    assert np.array_equal(ids.numpy(), fx["ids"]) and float(margins.min()) >= 0.5 and bool(fx["ids_bf16_autocast_equal"])
    # without the mask / position handling the padded sequences decode differently: the fixture really exercises it
    wrong, _ = po.greedy_generate(W, batch, cfg, max_new_tokens=2)
    assert not np.array_equal(wrong.numpy(), fx["ids"][:, :2])


def test_oracle_equals_recorded_reference_on_the_bench_parity_sample():
    """bench.py's CE-loss parity sample (full-size model, seed-1 weights, 1 x 30 s clip): the oracle's fp32 loss against the value the
    unmodified reference produced (tests/golden/reference_precision_gap.json), which also records the reference's own bf16-autocast
    loss -- the like-for-like yardstick bench.py reports next to the fp32 delta."""
    import json
    torch.set_num_threads(os.cpu_count())
    gap = json.load(open(os.path.join(GOLD, "reference_precision_gap.json")))
    c, s1 = gap["config"], gap["B1"]
    cfg = po.PathConfig(proj_hidden=c["proj_hidden"])
    W = po.init_weights(cfg, seed=c["weights_seed"])
    batch = po.synthetic_batch(cfg, s1["batch"], c["clip_seconds"], seed=s1["batch_seed"], response_len=c["response_len"])
    assert int((batch["labels"] != -100).sum()) == s1["num_items"]
    with torch.no_grad():
        loss, _ = po.model_forward(W, batch, cfg, s1["num_items"])
    assert abs(float(loss) - s1["ce_loss_reference_fp32"]) < 2e-5
    # the production recipe's own distance to fp32 on this sample is of the order of the 1e-3 budget
    assert 2e-4 < s1["gap"] < 3e-3 and gap["B4"]["gap"] < s1["gap"]
