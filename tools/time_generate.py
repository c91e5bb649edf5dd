"""Greedy generate at full model size: per-token latency of the KV-cache decode path vs the cache-free path."""
import os
import sys
import time
import torch

sys.path.insert(0, ".")
from tiny_audio_b200.engine import PathDims
from tiny_audio_b200.synthetic import build_offline_model, synthetic_batch

dims = PathDims(proj_hidden=2048)
model = build_offline_model(dims, device=torch.device("cuda"), seed=1234)
hot = model._hot_path()
params = {k: p.detach().float().contiguous() for k, p in model.projector.state_dict().items()}
clip = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
for B in ((32,) if os.environ.get('TA_PROFILE_STEP') == '1' else (1, 8, 32)):
    host = synthetic_batch(dims, B, clip, seed=5, response_len=4)
    n_prompt = int((host["labels"][0] != -100).nonzero().min())
    prompt = host["input_ids"][:, :n_prompt].cuda()
    kw = dict(proj_params=params, waveform=host["input_features"].cuda(), audio_token_counts=host["audio_token_counts"].cuda())
    for use_cache, T in ((True, 64), (True, 8), (False, 8)):
        hot.greedy_generate(input_ids=prompt, max_new_tokens=2, use_cache=use_cache, **kw)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ids = hot.greedy_generate(input_ids=prompt, max_new_tokens=T, use_cache=use_cache, **kw)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"B={B:2d} prompt {n_prompt} tokens, {T} new tokens, use_cache={use_cache}: {dt * 1e3:8.1f} ms total (encoder + prefill + decode)", flush=True)
    # decode step alone
    S0 = n_prompt
    cache = hot.new_kv_cache(B, S0 + 70)
    pos = torch.full((1,), S0, device="cuda", dtype=torch.int32)
    logits = torch.empty(B, hot.lm.vocab_pad, device="cuda", dtype=torch.bfloat16)
    ids = torch.zeros(B, device="cuda", dtype=torch.int64)
    nxt = torch.empty_like(ids)
    cache[0].zero_(); cache[1].zero_()
    for _ in range(3):
        hot.decode_step(ids, pos, S0, cache, logits, nxt)
    torch.cuda.synchronize()
    if os.environ.get('TA_PROFILE_STEP') == '1':           # ncu --profile-from-start off: one decode step
        pos.fill_(S0 + 40)
        torch.cuda.profiler.start()
        hot.decode_step(ids, pos, S0 + 40, cache, logits, nxt)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    pos.fill_(S0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    e0.record()
    for i in range(n):
        hot.decode_step(ids, pos, S0 + i, cache, logits, nxt)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    wbytes = 2.0 * (dims.lm_layers * (dims.lm_dim * (dims.lm_heads + 2 * dims.lm_kv_heads) * dims.lm_head_dim + dims.lm_heads * dims.lm_head_dim * dims.lm_dim
                                      + 3 * dims.lm_dim * dims.lm_ffn) + hot.lm.vocab_pad * dims.lm_dim)
    kvbytes = 2.0 * 2 * dims.lm_layers * B * (S0 + n / 2) * dims.lm_kv_heads * dims.lm_head_dim
    print(f"B={B:2d} decode step: {ms:.3f} ms/token = {B / ms * 1e3:.0f} tokens/s; algorithmic bytes {1e-6 * (wbytes + kvbytes):.0f} MB "
          f"-> {(wbytes + kvbytes) / ms * 1e-6:.0f} GB/s", flush=True)
    # CUDA-graph replay of the same step (position lives on the device)
    g = torch.cuda.CUDAGraph()
    pos.fill_(S0)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        hot.decode_step(ids, pos, S0, cache, logits, nxt)
        s.synchronize()
        with torch.cuda.graph(g, stream=s):
            hot.decode_step(ids, pos, S0, cache, logits, nxt)
    torch.cuda.synchronize()
    pos.fill_(S0)
    e0.record()
    for i in range(n):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"B={B:2d} decode step (CUDA graph replay): {ms:.3f} ms/token = {B / ms * 1e3:.0f} tokens/s -> {(wbytes + kvbytes) / ms * 1e-6:.0f} GB/s; "
          f"pos after replays {int(pos)} (expected {S0 + n})", flush=True)
