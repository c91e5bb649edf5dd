"""Ragged-batch greedy ids from the UNMODIFIED reference: tests/golden/generate_ragged.npz.

Runs only in the build container (needs /root/reference).  Test infrastructure, not product code.

ASRModel.generate (tiny_audio/asr_modeling.py:562-646) with clips of DIFFERENT lengths in one batch: per-sample audio token counts
come from `audio_attention_mask` (:587-589), the prompts carry different numbers of `<audio>` placeholders and are therefore
LEFT-padded to a common length (what HF generate requires of decoder-only models), `attention_mask` marks the padding.  HF then
masks the padding keys and counts rotary positions from each sequence's first real token.

Small seeded model (2 + 2 layers) sharpened with planted tokens exactly like the full-size fixture (oracle/make_generate_golden.py:
every greedy step's fp32 top-1 margin >= MARGIN, all generated ids distinct and fed back as inputs).  Checks: oracle == reference,
and the reference under torch.autocast(bfloat16) gives the same ids.

usage:  python oracle/make_ragged_generate_golden.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import path_oracle as po  # noqa: E402

CLIP_SECONDS = (1.0, 2.0, 1.5, 2.0)
NEW_TOKENS = 10
MARGIN = 1.0
SEED = 207


def case_inputs(seed: int):
    """(cfg, W, batch) with left-padded prompts: input_ids / attention_mask [B, S0], waveform [B, L_max] zero-padded,
    sample_lengths, audio_token_counts [B]."""
    cfg = po.small_config(enc_layers=2, lm_layers=2)
    W = po.init_weights(cfg, seed=seed)
    prompts, waves, counts = [], [], []
    for i, sec in enumerate(CLIP_SECONDS):
        b = po.synthetic_batch(cfg, 1, sec, seed=seed * 10 + i, response_len=2)
        first = int((b["labels"][0] != -100).nonzero()[0])
        prompts.append(b["input_ids"][0, :first])
        waves.append(b["waveform"][0, : int(b["sample_lengths"][0])])
        counts.append(int(b["audio_token_counts"][0]))
    S0 = max(len(p) for p in prompts)
    pad_id = min(151643, cfg.vocab - 2)
    ids = torch.full((len(prompts), S0), pad_id, dtype=torch.int64)
    am = torch.zeros((len(prompts), S0), dtype=torch.int64)
    for i, p in enumerate(prompts):
        ids[i, S0 - len(p):] = p
        am[i, S0 - len(p):] = 1
    L = max(len(w) for w in waves)
    wave = torch.zeros(len(waves), L)
    for i, w in enumerate(waves):
        wave[i, : len(w)] = w
    batch = {"input_ids": ids, "attention_mask": am, "waveform": wave, "sample_lengths": torch.tensor([len(w) for w in waves]),
             "audio_token_counts": torch.tensor(counts, dtype=torch.int64)}
    return cfg, W, batch


def main():
    torch.set_num_threads(os.cpu_count())
    from oracle.make_golden import build_reference_model, load_reference
    from oracle.make_generate_golden import apply_planted, plant, to_bf16_bits
    seed = SEED
    cfg, W, batch = case_inputs(seed)
    case = dict(new_tokens=NEW_TOKENS, margin=MARGIN, plant_seed=seed + 1)
    planted = plant(cfg, W, batch, batch["input_ids"], case=case, attention_mask=batch["attention_mask"], lo_id=100, hi_id=cfg.vocab - 10)
    rows_bits = to_bf16_bits(W["lm"]["model.embed_tokens.weight"][torch.from_numpy(planted)])
    cfg, W, batch = case_inputs(seed)                       # fresh weights + the planted rows: what the tests rebuild
    apply_planted(W, planted, rows_bits)
    ids, margins = po.greedy_generate(W, batch, cfg, max_new_tokens=NEW_TOKENS, attention_mask=batch["attention_mask"])
    print("min margin", float(margins.min()), "distinct", len(set(ids.reshape(-1).tolist())))
    assert float(margins.min()) >= 0.5 * MARGIN
    mods = load_reference()
    ref = build_reference_model(cfg, W, mods, "mlp")
    ref.eval()
    B = ids.shape[0]
    clips = [batch["waveform"][b, : int(batch["sample_lengths"][b])].numpy() for b in range(B)]
    feats = ref.feature_extractor(clips, sampling_rate=16000, padding="longest", return_attention_mask=True, return_tensors="pt")
    # the reference derives the per-sample audio token counts from the frame mask (asr_modeling.py:587-589)
    enc_len = ref._compute_encoder_output_lengths(feats.attention_mask)
    assert torch.equal(ref.projector.get_output_length(enc_len).to(torch.long), batch["audio_token_counts"])
    kw = dict(input_ids=batch["input_ids"], input_features=feats.input_features, audio_attention_mask=feats.attention_mask,
              attention_mask=batch["attention_mask"], max_new_tokens=NEW_TOKENS)
    out = ref.generate(**kw)
    print("reference", out.tolist())
    print("oracle   ", ids.tolist())
    assert torch.equal(out, ids), "reference generate != oracle on the ragged batch"
    with torch.autocast("cpu", dtype=torch.bfloat16):
        out_bf16 = ref.generate(**kw)
    same = bool(torch.equal(out_bf16, ids))
    print("bf16 autocast equal:", same)
    # each sequence alone (no padding) must give its row: the padding really is invisible
    for b in range(B):
        n = int(batch["attention_mask"][b].sum())
        f1 = ref.feature_extractor([clips[b]], sampling_rate=16000, padding="longest", return_attention_mask=True, return_tensors="pt")
        o1 = ref.generate(input_ids=batch["input_ids"][b: b + 1, -n:], input_features=f1.input_features, audio_attention_mask=f1.attention_mask,
                          attention_mask=torch.ones(1, n, dtype=torch.int64), max_new_tokens=NEW_TOKENS)
        if clips[b].shape[0] == batch["waveform"].shape[1]:          # (a shorter clip alone sees less zero padding in the encoder: skip)
            assert torch.equal(o1, ids[b: b + 1]), b
    path = os.path.join(ROOT, "tests", "golden", "generate_ragged.npz")
    np.savez_compressed(path, seed=np.array(seed), input_ids=batch["input_ids"].numpy(), attention_mask=batch["attention_mask"].numpy(),
                        audio_token_counts=batch["audio_token_counts"].numpy(), ids=out.numpy(), margins=margins.numpy().astype(np.float32),
                        ids_bf16_autocast_equal=np.array(same), mel_mask_sum=feats.attention_mask.sum(-1).numpy(),
                        planted_tokens=planted, planted_rows_bf16=rows_bits, sample_lengths=batch["sample_lengths"].numpy())
    print("wrote", path)


if __name__ == "__main__":
    main()
