"""Full-size greedy-id fixture: tests/golden/generate_full_b8.npz, produced by the UNMODIFIED reference's ASRModel.generate.

Runs only in the build container (needs /root/reference).  Test infrastructure, not product code.

North star: greedy token ids bit-exact.  With random-initialised towers the decoder's logits are nearly flat (top-1 / top-2 gaps of
~0.1 against ~0.05 of bf16 noise), so bit-exact FREE-RUNNING ids over 16 tokens are only well defined for a "sharpened" LM
("scale lm_head / plant a dominant token").  This script plants them:

  * weights  = oracle.path_oracle.init_weights(FULL, seed)  -- 32 + 28 layers, V = 151 936, the architecture bench.py runs;
  * prompts  = the chat-template prompt of 8 x 4 s clips, each followed by ONE distinct extra token so that the eight
               sequences take different paths through the decoder;
  * planting = decoding step by step with the fp32 oracle: at every step, for every sequence, the row of a FRESH token of the
               tied embedding / lm_head table is replaced by alpha * u, u = the minimum-norm vector with u . h = 1 for that
               sequence's final hidden state h and u . h' = 0 for every hidden state decided before (so earlier decisions are
               untouched), alpha = (current top logit + 1.25 MARGIN): the fresh token wins by >= MARGIN.  The rows are small
               (|row| ~ 0.2 against 0.65 for a random row), so a planted token has no tendency to predict itself -- the 128
               generated ids are all different -- and each is fed back as an input, so every later hidden state depends on the
               earlier ones through all 28 layers, the KV cache and the attention kernels: a wrong cache row or position moves
               h by O(1) and flips ids downstream (a flip needs a ~20 % change of h; bf16 noise is ~1 %).
  * rows are rounded to bf16-representable values before the final check (stored as 2 bytes per element).

Final check, independent of the planting plan, on the FINAL weights: (1) fp32 oracle greedy ids with every margin >= MARGIN / 2,
(2) the unmodified reference's ASRModel.generate -> HF generate in fp32 gives the same ids, (3) the same under
torch.autocast(bfloat16) -- the reference's production recipe (recorded in the fixture as `ids_bf16_autocast_equal`).

usage:  python oracle/make_generate_golden.py
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import path_oracle as po  # noqa: E402

CASE = dict(weights_seed=21, batch_seed=33, batch=8, clip_s=4.0, new_tokens=16, margin=1.0, plant_seed=5)


def case_inputs():
    """(cfg, W, batch, prompt [B, S0]) -- seeded; the planted rows are NOT applied here (apply_planted does that)."""
    c = CASE
    cfg = po.PathConfig()
    W = po.init_weights(cfg, seed=c["weights_seed"])
    batch = po.synthetic_batch(cfg, c["batch"], c["clip_s"], seed=c["batch_seed"], response_len=2)
    first = int((batch["labels"][0] != -100).nonzero()[0])
    rng = np.random.default_rng(c["batch_seed"] + 1)
    extra = torch.from_numpy(rng.choice(np.arange(2000, 150000), size=c["batch"], replace=False).astype(np.int64))
    prompt = torch.cat([batch["input_ids"][:, :first], extra[:, None]], 1).contiguous()
    return cfg, W, batch, prompt


def apply_planted(W, token_ids: np.ndarray, rows_bf16_bits: np.ndarray):
    """Overwrite the planted rows of the tied embedding / lm_head table (bf16 bit patterns -> fp32)."""
    rows = torch.from_numpy(rows_bf16_bits.astype(np.int32) << 16).view(torch.float32)
    W["lm"]["model.embed_tokens.weight"][torch.from_numpy(token_ids.astype(np.int64))] = rows
    assert W["lm"]["lm_head.weight"].data_ptr() == W["lm"]["model.embed_tokens.weight"].data_ptr()


def to_bf16_bits(x: torch.Tensor) -> np.ndarray:
    return (x.to(torch.bfloat16).view(torch.int16).numpy().astype(np.int32) & 0xFFFF).astype(np.uint16)


@torch.no_grad()
def last_hidden(W, cfg, ids, packed, am=None):
    emb = F.embedding(ids, W["lm"]["model.embed_tokens.weight"])
    emb = po.scatter_audio(emb, ids, packed, cfg.audio_token_id)
    pos = (am.cumsum(-1) - 1).clamp(min=0) if am is not None else None
    return po.lm_forward(W["lm"], emb, cfg, attention_mask=am, position_ids=pos)[:, -1]


@torch.no_grad()
def plant(cfg, W, batch, prompt, case=None, attention_mask=None, lo_id=2000, hi_id=150000):
    """`attention_mask`: left-padded prompts (ragged batches, oracle/make_ragged_generate_golden.py)."""
    c = case or CASE
    am = attention_mask.clone() if attention_mask is not None else None
    E = W["lm"]["model.embed_tokens.weight"]
    mel = po.log_mel(batch["waveform"], cfg)
    audio = po.projector_forward(W["projector"], po.encoder_forward(W["encoder"], mel, cfg), cfg)
    packed = po.gather_audio_embeds(audio, batch["audio_token_counts"])
    rng = np.random.default_rng(c["plant_seed"])
    used = set(prompt.reshape(-1).tolist()) | {cfg.audio_token_id}
    planted, stats = [], []
    past = []                                     # every final hidden state decided so far (fp64 rows)
    ids = prompt.clone()
    B = ids.shape[0]
    for t in range(c["new_tokens"]):
        t0 = time.time()
        hid = last_hidden(W, cfg, ids, packed, am).double()        # [B, D]
        for b in range(B):
            logits = (E.double() @ hid[b])
            top2 = logits.topk(2)
            while True:
                r = int(rng.integers(lo_id, hi_id))
                if r not in used:
                    break
            used.add(r)
            # minimum-norm u with u.hid[b] = 1 and u.h = 0 for every hidden state decided earlier (previous steps, and the
            # sequences of this step that are already settled), so their logits for token r do not move
            cons = past + [hid[j] for j in range(b)]
            A = torch.stack(cons + [hid[b]]) if cons else hid[b][None]
            rhs = torch.zeros(A.shape[0], dtype=torch.float64)
            rhs[-1] = 1.0
            u = torch.linalg.lstsq(A, rhs[:, None]).solution[:, 0]
            alpha = float(top2.values[0]) + 1.25 * c["margin"]
            new_row = (alpha * u).float().to(torch.bfloat16).float()                        # bf16-representable replacement row
            E[r] = new_row
            planted.append(r)
            stats.append((alpha, float(new_row.norm()), float(u.norm() * hid[b].norm())))
            chk = (E.double() @ hid[b]).topk(2)
            assert int(chk.indices[0]) == r and float(chk.values[0] - chk.values[1]) >= c["margin"], (t, b, chk)
        # all sequences of this step are settled: decide
        logits = hid.float() @ E.t()
        nxt = logits.argmax(-1)
        past.extend(hid[b] for b in range(B))
        ids = torch.cat([ids, nxt[:, None]], 1)
        if am is not None:
            am = torch.cat([am, torch.ones_like(am[:, :1])], 1)
        st = stats[-B:]
        print(f"step {t}: ids {nxt.tolist()}  alpha {min(a for a, _, _ in st):.2f}..{max(a for a, _, _ in st):.2f}  |row| "
              f"{max(n for _, n, _ in st):.3f}  inflation {max(i for _, _, i in st):.2f}  ({time.time() - t0:.1f}s)", flush=True)
    return np.array(planted, dtype=np.int64)


def main():
    torch.set_num_threads(os.cpu_count())
    from oracle.make_golden import build_reference_model, load_reference
    c = CASE
    cfg, W, batch, prompt = case_inputs()
    planted = plant(cfg, W, batch, prompt)
    rows_bits = to_bf16_bits(W["lm"]["model.embed_tokens.weight"][torch.from_numpy(planted)])

    # ---- final check on a fresh copy of the weights (what the tests will rebuild) ----
    cfg, W, batch, prompt = case_inputs()
    apply_planted(W, planted, rows_bits)
    ob = dict(batch, input_ids=prompt)
    ids, margins = po.greedy_generate(W, ob, cfg, max_new_tokens=c["new_tokens"])
    print("oracle ids", ids.tolist(), "min margin", float(margins.min()))
    assert float(margins.min()) >= 0.5 * c["margin"]

    mods = load_reference()
    ref = build_reference_model(cfg, W, mods, "mlp")
    ref.eval()
    L = int(batch["sample_lengths"][0])
    feats = ref.feature_extractor([batch["waveform"][b, :L].numpy() for b in range(prompt.shape[0])], sampling_rate=16000,
                                  padding="longest", return_attention_mask=True, return_tensors="pt")
    kw = dict(input_ids=prompt, input_features=feats.input_features, audio_attention_mask=feats.attention_mask,
              attention_mask=torch.ones_like(prompt), max_new_tokens=c["new_tokens"])
    out = ref.generate(**kw)
    assert torch.equal(out, ids), "reference fp32 generate != oracle"
    with torch.autocast("cpu", dtype=torch.bfloat16):
        out_bf16 = ref.generate(**kw)
    same_bf16 = bool(torch.equal(out_bf16, ids))
    print("reference generate == oracle; under bf16 autocast equal:", same_bf16)
    # single-sequence run (B = 1): sample 0 alone must give row 0 (no cross-sample coupling)
    out1 = ref.generate(input_ids=prompt[:1], input_features=feats.input_features[:1], audio_attention_mask=feats.attention_mask[:1],
                        attention_mask=torch.ones_like(prompt[:1]), max_new_tokens=c["new_tokens"])
    assert torch.equal(out1, ids[:1])
    path = os.path.join(ROOT, "tests", "golden", "generate_full_b8.npz")
    np.savez_compressed(path, prompt=prompt.numpy(), ids=out.numpy(), margins=margins.numpy().astype(np.float32),
                        planted_tokens=planted, planted_rows_bf16=rows_bits, ids_bf16_autocast_equal=np.array(same_bf16),
                        case=np.array(str(c)))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
