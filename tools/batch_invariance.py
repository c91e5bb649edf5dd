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
