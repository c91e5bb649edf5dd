"""CE loss of the UNMODIFIED reference in fp32 and under its production recipe (fp32 masters + bf16 autocast,
configs/training/production.yaml:49) on bench.py's two parity samples -- full-size model, seeded weights (seed 1), 30 s clips.
Runs only where /root/reference exists (needs ~12 GB of RAM, a minute of CPU).  Output: tests/golden/reference_precision_gap.json,
which bench.py reads to report the CUDA path's distance to the reference at LIKE-FOR-LIKE precision next to the fp32 delta.

usage:  python oracle/make_reference_precision_gap.py
"""
from __future__ import annotations

import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import path_oracle as po  # noqa: E402
from oracle.make_golden import build_reference_model, load_reference  # noqa: E402

SAMPLES = {"B1": dict(batch=1, batch_seed=0), "B4": dict(batch=4, batch_seed=1)}      # bench.py: cpu_sample_batch / parity_batch
COMMON = dict(weights_seed=1, proj_hidden=2048, clip_seconds=30.0, response_len=64)


def gradient_gap(mods):
    """Projector gradients of the reference, fp32 vs bf16 autocast, on the full-size 4 s fixture (tests/golden/full_b1_4s.npz's
    inputs): the yardstick for the CUDA path's 2.8-3.6e-2 against fp32 (tests/test_path_gpu.py)."""
    from oracle.make_golden import CASES
    _, B, clip_s, _, R, seed = CASES["full_b1_4s"]
    cfg = po.FULL
    W = po.init_weights(cfg, seed=seed)
    batch = po.synthetic_batch(cfg, B, clip_s, seed=seed, response_len=R)
    n_items = int((batch["labels"] != -100).sum())
    ref = build_reference_model(cfg, W, mods, "mlp")
    ref.train()
    L = int(batch["sample_lengths"][0])
    feats = ref.feature_extractor([batch["waveform"][b, :L].numpy() for b in range(B)], sampling_rate=16000, padding="longest",
                                  return_attention_mask=True, return_tensors="pt")
    rb = dict(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"], labels=batch["labels"],
              input_features=feats.input_features, audio_attention_mask=feats.attention_mask,
              audio_token_counts=batch["audio_token_counts"])

    def run(autocast):
        ref.zero_grad()
        with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
            loss = ref(**rb, num_items_in_batch=torch.tensor(n_items)).loss
        loss.backward()
        return float(loss.detach()), {k: p.grad.detach().clone() for k, p in ref.projector.named_parameters()}

    l32, g32 = run(False)
    l16, g16 = run(True)
    return {"ce_loss_reference_fp32": l32, "ce_loss_reference_bf16_autocast": l16,
            "grad_rel_err_bf16_vs_fp32": {k: float((g16[k].float() - g32[k]).norm() / g32[k].norm()) for k in g32}}


SMALL_CASES = ("small_b2_2s", "small_b3_ragged", "h2048_b2_2s", "small_b2_1s_pad30", "dropout_b2_2s")


def small_case_gaps(mods):
    """The reference's own fp32 and bf16-autocast CE on the small parity cases of tests/golden (same weights, same inputs): the
    like-for-like target the GPU parity tests assert 1e-3 against (tests/test_path_gpu.py)."""
    import numpy as np
    from oracle.make_golden import CASES, DROPOUT_SEED
    out = {}
    for name in SMALL_CASES:
        spec, B, clip_s, pad_s, R, seed = CASES[name]
        cfg = po.small_config(**{k: v for k, v in spec.items() if not k.startswith("_")})
        fx = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        W = po.init_weights(cfg, seed=seed)
        batch = po.synthetic_batch(cfg, B, clip_s, seed=seed, response_len=R, pad_to_seconds=pad_s)
        ref = build_reference_model(cfg, W, mods, "mlp")
        ref.train()
        p_drop = float(spec.get("_dropout", 0.0))
        ref.config.audio_token_dropout = p_drop
        L = int(batch["sample_lengths"][0])
        feats = ref.feature_extractor([batch["waveform"][b, :L].numpy() for b in range(B)], sampling_rate=16000,
                                      padding="max_length" if pad_s == 30.0 else "longest", return_attention_mask=True, return_tensors="pt")
        rb = dict(input_ids=torch.from_numpy(fx["input_ids"]), attention_mask=torch.from_numpy(fx["attention_mask"]),
                  labels=torch.from_numpy(fx["labels"]), input_features=feats.input_features, audio_attention_mask=feats.attention_mask,
                  audio_token_counts=torch.from_numpy(fx["audio_token_counts"]))
        n_items = int(fx["num_items"])
        with torch.no_grad():
            torch.manual_seed(DROPOUT_SEED)
            fp32 = float(ref(**rb, num_items_in_batch=torch.tensor(n_items)).loss)
            torch.manual_seed(DROPOUT_SEED)          # fp32 hidden states under autocast too (final LayerNorm): same Bernoulli draw
            with torch.autocast("cpu", dtype=torch.bfloat16):
                bf16 = float(ref(**rb, num_items_in_batch=torch.tensor(n_items)).loss)
        assert abs(fp32 - float(fx["loss"])) < 2e-5, (name, fp32, float(fx["loss"]))
        out[name] = dict(num_items=n_items, ce_loss_reference_fp32=fp32, ce_loss_reference_bf16_autocast=bf16, gap=abs(fp32 - bf16))
        print(name, out[name], flush=True)
    return out


def main():
    torch.set_num_threads(os.cpu_count())
    mods = load_reference()
    path = os.path.join(ROOT, "tests", "golden", "reference_precision_gap.json")
    if "--small-only" in sys.argv:       # add / refresh the small cases without redoing the full-size runs
        with open(path) as f:
            out = json.load(f)
        out["small_cases"] = small_case_gaps(mods)
        with open(path, "w") as f:
            json.dump(out, f, indent=1)
        return
    cfg = po.PathConfig(proj_hidden=COMMON["proj_hidden"])
    W = po.init_weights(cfg, seed=COMMON["weights_seed"])
    ref = build_reference_model(cfg, W, mods, "mlp")
    ref.train()
    out = {"config": dict(COMMON, model="GLM-ASR encoder 32L + Qwen3-0.6B 28L, MLP projector", autocast="torch.autocast('cpu', bfloat16)",
                          torch=torch.__version__)}
    for name, s in SAMPLES.items():
        batch = po.synthetic_batch(cfg, s["batch"], COMMON["clip_seconds"], seed=s["batch_seed"], response_len=COMMON["response_len"])
        n_items = int((batch["labels"] != -100).sum())
        L = int(batch["sample_lengths"][0])
        feats = ref.feature_extractor([batch["waveform"][b, :L].numpy() for b in range(s["batch"])], sampling_rate=16000,
                                      padding="longest", return_attention_mask=True, return_tensors="pt")
        rb = dict(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"], labels=batch["labels"],
                  input_features=feats.input_features, audio_attention_mask=feats.attention_mask,
                  audio_token_counts=batch["audio_token_counts"])
        with torch.no_grad():
            fp32 = float(ref(**rb, num_items_in_batch=torch.tensor(n_items)).loss)
            with torch.autocast("cpu", dtype=torch.bfloat16):
                bf16 = float(ref(**rb, num_items_in_batch=torch.tensor(n_items)).loss)
        out[name] = dict(s, num_items=n_items, ce_loss_reference_fp32=fp32, ce_loss_reference_bf16_autocast=bf16, gap=abs(fp32 - bf16))
        print(name, out[name], flush=True)
    out["full_b1_4s_gradients"] = gradient_gap(mods)
    print(out["full_b1_4s_gradients"], flush=True)
    out["small_cases"] = small_case_gaps(mods)
    with open(path, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
