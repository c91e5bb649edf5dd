"""Integer golden vectors from the UNMODIFIED reference (runs only where /root/reference exists): the token-count arithmetic the
collator and the processor rely on (scripts/train.py:335-338, tiny_audio/asr_processing.py:95-110) for every registered projector,
and the ragged gather of audio embeddings (tiny_audio/asr_modeling.py:27-44).  Output: tests/golden/integer_semantics.npz

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


def main():
    mods = load_reference()
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
