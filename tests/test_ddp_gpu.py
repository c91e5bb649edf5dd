"""Two data-parallel ranks driving the CUDA path the two ways it is used:

  (a) the HF Trainer / accelerate way -- ASRModel wrapped in DistributedDataParallel (bucketed MEAN all-reduce during backward),
      the loss multiplied by the number of processes (HF:trainer.py average_tokens_across_devices), ClipAdamW(ddp_wrapped=True)
      which must NOT reduce again;
  (b) the plain torchrun way bench.py uses -- no wrapper, ClipAdamW(allreduce=True) does the one SUM all-reduce.

Both must reproduce the single-process step on the whole global batch (same seeded weights, ragged label counts so that per-rank
means would be wrong).  The two ranks share cuda:0 and talk over gloo (NCCL refuses two ranks on one device), so the test runs on
the one-GPU box the driver uses; device-resident labels exercise ta_label_rows on the way."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(cfg, W, dev):
    from tiny_audio_b200.engine import PathDims
    from tiny_audio_b200.synthetic import build_offline_model
    m = build_offline_model(PathDims.from_any(cfg.to_dict()), device=dev, enc_state=W["encoder"], lm_state=W["lm"], proj_state=W["projector"])
    m.train()
    return m


def _worker(rank, world, port, out):
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    from oracle import path_oracle as po          # seeded weights / batch only
    from tiny_audio_b200 import dp
    from tiny_audio_b200.optim import ClipAdamW
    cfg = po.small_config(enc_layers=1, lm_layers=1)
    W = po.init_weights(cfg, seed=3)
    full = po.synthetic_batch(cfg, 4, 1.0, seed=3, response_len=6)
    full["labels"][1, -4:] = -100          # ragged label counts
    keys = ("waveform", "input_ids", "labels", "audio_token_counts")
    mine = {k: v.to(dev) for k, v in dp.shard_batch({k: full[k] for k in keys}, rank, world).items()}
    n_global = dp.global_num_items(mine["labels"], device=torch.device("cpu"))

    def call(model, b, nib):
        return model(input_ids=b["input_ids"], input_features=b["waveform"], labels=b["labels"], audio_token_counts=b["audio_token_counts"],
                     num_items_in_batch=nib).loss

    res = {}
    # (a) DDP wrapper + Trainer's loss scaling; the optimiser is told not to reduce again
    m_a = _build(cfg, W, dev)
    ddp = DDP(m_a, device_ids=[0])
    opt_a = ClipAdamW([p for p in m_a.parameters() if p.requires_grad], lr=1e-3, max_grad_norm=1.0, ddp_wrapped=True)
    opt_a.zero_grad()
    loss = call(ddp, mine, n_global) * world
    loss.backward()
    opt_a.step()
    res["a"] = (opt_a.flat_grad.clone().cpu(), float(opt_a.grad_norm()), [p.detach().clone().cpu() for p in opt_a._params])
    # (b) no wrapper: ClipAdamW owns the SUM all-reduce
    m_b = _build(cfg, W, dev)
    opt_b = ClipAdamW([p for p in m_b.parameters() if p.requires_grad], lr=1e-3, max_grad_norm=1.0, allreduce=True)
    opt_b.zero_grad()
    call(m_b, mine, n_global).backward()
    opt_b.step()
    res["b"] = (opt_b.flat_grad.clone().cpu(), float(opt_b.grad_norm()), [p.detach().clone().cpu() for p in opt_b._params])
    if rank == 0:
        # single process, whole global batch
        m_s = _build(cfg, W, dev)
        opt_s = ClipAdamW([p for p in m_s.parameters() if p.requires_grad], lr=1e-3, max_grad_norm=1.0, allreduce=False)
        opt_s.zero_grad()
        fb = {k: full[k].to(dev) for k in keys}
        call(m_s, fb, n_global).backward()
        opt_s.step()
        res["single"] = (opt_s.flat_grad.clone().cpu(), float(opt_s.grad_norm()), [p.detach().clone().cpu() for p in opt_s._params])
        res["n_global"] = (n_global, int((full["labels"] != -100).sum()))
        torch.save(res, out)
    dist.barrier()
    dist.destroy_process_group()


def test_ddp_wrapped_and_plain_torchrun_steps_equal_the_single_process_step(cuda, tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "ddp.pt")
    mp.get_context("spawn")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    g_s, n_s, p_s = res["single"]
    assert n_s > 0 and res["n_global"][0] == res["n_global"][1] and res["n_global"][0] % 7 != 0     # global count, ragged
    for tag in ("a", "b"):
        g, n, ps = res[tag]
        rel = float((g - g_s).norm() / g_s.norm())
        print(f"[ddp {tag}] reduced-gradient rel err vs single process {rel:.2e}, clip norm {n:.5f} vs {n_s:.5f}")
        assert rel < 2e-3, (tag, rel)                     # shards vs whole batch: fp32 summation order + bf16 d(logits) rounding
        assert abs(n - n_s) < 2e-3 * n_s, (tag, n, n_s)   # a double reduction would show up as a factor of 2 here
        for a, b in zip(ps, p_s):
            assert float((a - b).abs().max()) < 2.5e-4    # lr = 1e-3: one AdamW step moves every element by <= ~1e-3
