"""Is a clip's result independent of the batch it is computed in?  Full-size model, one batch of N clips vs its two halves:
mel, encoder output, projector output and the loss are compared bit for bit (forward only)."""
import sys
import torch
sys.path.insert(0, ".")
from tiny_audio_b200.engine import HotPath, PathDims
from tiny_audio_b200.synthetic import build_offline_model, synthetic_batch

N, clip_s = int(sys.argv[1]) if len(sys.argv) > 1 else 8, float(sys.argv[2]) if len(sys.argv) > 2 else 30.0
dims = PathDims(proj_hidden=2048)
dev = torch.device("cuda")
model = build_offline_model(dims, device=dev, seed=1234)
hp = model._hot_path()
pmap = {n: p.data for n, p in model.projector.named_parameters()}
gb = synthetic_batch(dims, N, clip_s, seed=4242, response_len=64)
n_items = int((gb["labels"] != -100).sum())


def run(sl):
    b = {k: gb[k][sl] for k in ("input_features", "input_ids", "labels", "audio_token_counts")}
    grads = {n: torch.zeros_like(p) for n, p in pmap.items()}
    loss, parts = hp.forward_backward(input_ids=b["input_ids"].to(dev), labels=b["labels"], proj_params=pmap, waveform=b["input_features"].to(dev),
                                      audio_token_counts=b["audio_token_counts"].to(dev), num_items_in_batch=n_items, grads=grads, return_parts=True)
    torch.cuda.synchronize()
    return float(loss), {k: v.detach().clone() for k, v in parts.items() if torch.is_tensor(v)}


l_full, p_full = run(slice(0, N))
l_a, p_a = run(slice(0, N // 2))
l_b, p_b = run(slice(N // 2, N))
l_full2, _ = run(slice(0, N))
print(f"loss full {l_full:.7f} (again {l_full2:.7f})  halves {l_a + l_b:.7f}  delta {abs(l_full - l_a - l_b):.3e}  run-to-run {abs(l_full - l_full2):.3e}")
for k in p_full:
    f = p_full[k]
    if k not in p_a or f.shape[0] != N and f.dim() < 2:
        continue
    try:
        h = torch.cat([p_a[k], p_b[k]], 0)
    except Exception as e:
        print(k, "cat failed", e)
        continue
    if h.shape != f.shape:
        print(f"{k}: shapes {tuple(f.shape)} vs {tuple(h.shape)}")
        continue
    d = (f.float() - h.float()).abs().max()
    print(f"{k:16s} {tuple(f.shape)} bitwise equal {torch.equal(f, h)}  max |diff| {float(d):.3e}")

# ---- backward: projector gradients of the whole batch vs the sum over equal shards (every shard normalised by the global token count)
def grads_of(sl):
    b = {k: gb[k][sl] for k in ("input_features", "input_ids", "labels", "audio_token_counts")}
    grads = {n: torch.zeros_like(p) for n, p in pmap.items()}
    hp.forward_backward(input_ids=b["input_ids"].to(dev), labels=b["labels"], proj_params=pmap, waveform=b["input_features"].to(dev),
                        audio_token_counts=b["audio_token_counts"].to(dev), num_items_in_batch=n_items, grads=grads)
    torch.cuda.synchronize()
    return {n: g.double() for n, g in grads.items()}


g_full = grads_of(slice(0, N))
g_full2 = grads_of(slice(0, N))
print("run-to-run (whole batch):", {n: f"{float((g_full[n] - g_full2[n]).norm() / g_full[n].norm()):.2e}" for n in g_full})
for shards in (2, 4, 8):
    if N % shards:
        continue
    acc = None
    for i in range(shards):
        g = grads_of(slice(i * N // shards, (i + 1) * N // shards))
        acc = g if acc is None else {n: acc[n] + g[n] for n in g}
    print(f"{shards} shards of {N // shards}:", {n: f"{float((acc[n] - g_full[n]).norm() / g_full[n].norm()):.2e}" for n in acc})

# ---- where does the backward start to depend on the batch?  d(loss)/d(audio embeddings) out of the decoder, whole batch vs halves,
#      under different GEMM kernel selections
lib = __import__("tiny_audio_b200.lib", fromlist=["x"]).load()
audio_full = p_full["projector_out"].float().reshape(-1, dims.lm_dim).contiguous()
n_a = p_full["projector_out"].shape[1]


def d_audio_of(sl):
    ids = gb["input_ids"][sl].to(dev)
    a = p_full["projector_out"][sl].float().reshape(-1, dims.lm_dim).contiguous()
    _, da = hp.lm_loss_and_audio_grad(input_ids=ids, audio=a, n_a=n_a, labels=gb["labels"][sl], audio_token_counts=gb["audio_token_counts"][sl].to(dev),
                                      num_items_in_batch=n_items)
    torch.cuda.synchronize()
    return da.clone().double()


for name, setup in (("default", lambda: None), ("1-CTA GEMM kernels", lambda: lib.ta_gemm_set_cta_pair(0)), ("128-wide tiles", lambda: lib.ta_gemm_set_tile_n(128)),
                    ("attention bwd variant 1", lambda: lib.ta_attn_set_bwd_variant(1)), ("no ROWDOT fusion", lambda: lib.ta_lm_set_fused_attn_dsum(0)),
                    ("SwiGLU-bwd per-thread epilogue", lambda: lib.ta_gemm_set_swiglu_bwd_tma(0))):
    lib.ta_gemm_set_cta_pair(1); lib.ta_gemm_set_tile_n(0); lib.ta_attn_set_bwd_variant(2); lib.ta_lm_set_fused_attn_dsum(1); lib.ta_gemm_set_swiglu_bwd_tma(1)
    setup()
    full = d_audio_of(slice(0, N))
    halves = torch.cat([d_audio_of(slice(0, N // 2)), d_audio_of(slice(N // 2, N))], 0)
    eighths = torch.cat([d_audio_of(slice(i * N // 8, (i + 1) * N // 8)) for i in range(8)], 0) if N % 8 == 0 else halves
    print(f"[{name}] d_audio whole vs halves rel {float((full - halves).norm() / full.norm()):.2e}, vs eighths {float((full - eighths).norm() / full.norm()):.2e}")
lib.ta_gemm_set_cta_pair(1); lib.ta_gemm_set_tile_n(0); lib.ta_attn_set_bwd_variant(2); lib.ta_lm_set_fused_attn_dsum(1); lib.ta_gemm_set_swiglu_bwd_tma(1)
