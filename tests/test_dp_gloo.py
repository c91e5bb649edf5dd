"""N > 1 host logic on CPU with the gloo backend (world_size 2): batch sharding, global num_items, and the
equivalence  SUM_r grad_r(loss_r / N_global)  ==  grad(full batch)  that the single all-reduce relies on.
The compute here is the oracle (tests may use it); the CUDA path is covered by -m gpu tests."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from oracle import path_oracle as po
    from tiny_audio_b200 import dp
    cfg = po.small_config(enc_layers=1, lm_layers=1)
    W = po.init_weights(cfg, seed=3)
    full = po.synthetic_batch(cfg, 4, 1.0, seed=3, response_len=6)
    full["labels"][1, -4:] = -100          # ragged label counts so that per-rank means would be WRONG
    keys = ("waveform", "input_ids", "labels", "attention_mask", "audio_token_counts", "sample_lengths")
    mine = dp.shard_batch({k: full[k] for k in keys}, rank, world)
    n_global = dp.global_num_items(mine["labels"])
    res = po.train_step(W, mine, cfg, num_items_in_batch=n_global, train_lm=True)
    # one flat buffer: projector gradients followed by every decoder gradient (the unfrozen recipe all-reduces both at once)
    flat = torch.cat([g.reshape(-1) for g in res["grads"].values()] + [res["lm_grads"][k].reshape(-1) for k in sorted(res["lm_grads"])])
    dp.allreduce_flat_(flat)
    loss = res["loss"].clone()
    dist.all_reduce(loss)
    if rank == 0:
        ref = po.train_step(W, {k: full[k] for k in keys}, cfg, num_items_in_batch=int((full["labels"] != -100).sum()), train_lm=True)
        flat_ref = torch.cat([g.reshape(-1) for g in ref["grads"].values()] + [ref["lm_grads"][k].reshape(-1) for k in sorted(ref["lm_grads"])])
        out.put((n_global, int((full["labels"] != -100).sum()), float(loss), float(ref["loss"]),
                 float((flat - flat_ref).norm() / flat_ref.norm())))
    dist.destroy_process_group()


def test_two_rank_gradient_sum_equals_full_batch():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    n_global, n_ref, loss, loss_ref, err = out.get()
    assert n_global == n_ref
    assert abs(loss - loss_ref) < 1e-5
    assert err < 1e-5


def test_shard_batch_requires_even_split():
    from tiny_audio_b200 import dp
    with pytest.raises(ValueError):
        dp.shard_batch({"x": torch.zeros(5, 2)}, 0, 2)
    s = dp.shard_batch({"x": torch.arange(8).view(8, 1)}, 1, 4)
    assert s["x"].view(-1).tolist() == [2, 3]
