"""torch.optim-compatible front for the fused clip + AdamW kernels, with the data-parallel gradient all-reduce.

Replaces, on the hot path, HF Trainer's `clip_grad_norm_(1.0)` + `torch.optim.AdamW(fused=True)`
(configs/training/production.yaml:5-9) and the DDP bucket all-reduce accelerate would add under torchrun
(the reference itself has no distributed code).  All trainable parameters' gradients live in ONE
flat fp32 buffer (`param.grad` are views of it), so a step is:

    [NCCL all_reduce(flat_grad, SUM)]  ->  ta_grad_sumsq  ->  ta_adamw_clip_step        (no host sync)

The loss is normalised by the GLOBAL number of label tokens (num_items_in_batch), so SUM over ranks yields the
gradient of the global token-mean CE -- the same quantity HF Trainer produces with
average_tokens_across_devices (HF:trainer.py:2141-2143, 2013-2018).

Who reduces the gradients is explicit (`allreduce`):
  * allreduce=True  -- this optimiser does the one SUM all-reduce (bench.py, plain torchrun loops: the model is NOT wrapped
    in DistributedDataParallel and the loss is NOT multiplied by the world size);
  * allreduce=False -- somebody else already reduced them: under HF Trainer / accelerate the model is DDP-wrapped (bucketed
    MEAN all-reduce during backward) and Trainer multiplies the loss by the number of processes, which together give the
    same SUM; reducing again here would scale the gradient by the world size before the clip;
  * allreduce=None (default) -- True iff a process group with more than one rank exists AND `ddp_wrapped` was not declared.
    Pass `ddp_wrapped=True` (or allreduce=False) whenever the model goes through DistributedDataParallel.

Optimiser state lives in flat buffers but is exposed through `self.state` in torch.optim.AdamW's layout (`step`, `exp_avg`,
`exp_avg_sq` per parameter), so `state_dict()` / `load_state_dict()` -- what HF Trainer writes to / reads from optimizer.pt --
round-trip the moments and the step count, and interchange with the reference's `adamw_torch_fused` checkpoints.
"""
from __future__ import annotations

from typing import Iterable

import torch

from . import lib as L


class ClipAdamW(torch.optim.Optimizer):
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, max_grad_norm: float = 1.0, process_group=None, allreduce=None,
                 ddp_wrapped: bool = False):
        params = list(params)
        if params and isinstance(params[0], dict):      # torch-style parameter groups (decoder lr / weight decay: train.py:384-437)
            groups = [dict(g, params=[p for p in g["params"] if p.requires_grad]) for g in params]
            groups = [g for g in groups if g["params"]]
            super().__init__(groups, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
            params = [p for g in groups for p in g["params"]]
        else:
            params = [p for p in params if p.requires_grad]
            super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.lib = L.load()
        self.max_grad_norm = max_grad_norm
        self.process_group = process_group
        self.allreduce = (not ddp_wrapped) if allreduce is None else bool(allreduce)
        self._params = params
        self._index = {id(p): i for i, p in enumerate(params)}
        for p in params:
            L.require_cuda(p)
            assert p.dtype == torch.float32 and p.is_contiguous(), "ClipAdamW needs contiguous fp32 parameters"
        n = sum(p.numel() for p in params)
        dev = params[0].device
        self.flat_grad = torch.zeros(n, device=dev, dtype=torch.float32)
        self.m = torch.zeros(n, device=dev, dtype=torch.float32)
        self.v = torch.zeros(n, device=dev, dtype=torch.float32)
        self.gnorm_sq = torch.zeros(1, device=dev, dtype=torch.float32)
        self.step_count = 0
        off = 0
        self._slices = []
        for p in params:
            k = p.numel()
            p.grad = self.flat_grad[off: off + k].view_as(p)      # autograd accumulates in place into the flat buffer
            self._slices.append((off, k))
            off += k
        self._bind_state()

    # ------------------------------------------------------------------ torch.optim.AdamW-compatible state
    def _bind_state(self):
        """self.state[p] = views of the flat moment buffers + the (global) step count, in torch.optim.AdamW's layout."""
        for p, (off, k) in zip(self._params, self._slices):
            self.state[p] = {"step": torch.tensor(float(self.step_count)), "exp_avg": self.m[off: off + k].view_as(p),
                             "exp_avg_sq": self.v[off: off + k].view_as(p)}

    def state_dict(self):
        for p in self._params:
            self.state[p]["step"] = torch.tensor(float(self.step_count))
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)          # fills self.state[p] with (copies of) the saved tensors
        steps = set()
        for p, (off, k) in zip(self._params, self._slices):
            st = self.state.get(p)
            if not st:
                continue
            self.m[off: off + k].copy_(st["exp_avg"].reshape(-1))
            self.v[off: off + k].copy_(st["exp_avg_sq"].reshape(-1))
            steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError(f"ClipAdamW keeps one step count for all parameters; the checkpoint holds {sorted(steps)}")
        self.step_count = steps.pop() if steps else 0
        self._bind_state()
        self.zero_grad()                             # re-bind the .grad views

    def zero_grad(self, set_to_none: bool = False):
        self.flat_grad.zero_()
        for p, (off, k) in zip(self._params, self._slices):
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * off:
                p.grad = self.flat_grad[off: off + k].view_as(p)

    def grad_norm(self) -> torch.Tensor:
        return self.gnorm_sq.sqrt()

    @torch.no_grad()
    def step(self, closure=None):
        for p, (off, k) in zip(self._params, self._slices):
            if p.grad is None:                                     # no gradient this step (zero_grad(set_to_none=True) + unused):
                self.flat_grad[off: off + k].zero_()               # the slice must not keep the previous step's values
            elif p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * off:
                self.flat_grad[off: off + k].copy_(p.grad.reshape(-1))      # a fresh .grad tensor (not our view) is folded in
        if self.allreduce and torch.distributed.is_available() and torch.distributed.is_initialized() and \
                torch.distributed.get_world_size(self.process_group) > 1:
            torch.distributed.all_reduce(self.flat_grad, op=torch.distributed.ReduceOp.SUM, group=self.process_group)
        st = L.stream_ptr()
        self.step_count += 1
        self.gnorm_sq.zero_()
        L.check(self.lib.ta_grad_sumsq(L.ptr(self.flat_grad), self.flat_grad.numel(), L.ptr(self.gnorm_sq), st))
        g = self.param_groups[0]
        for group in self.param_groups:
            for p in group["params"]:
                i = self._index[id(p)]
                off, k = self._slices[i]
                L.check(self.lib.ta_adamw_clip_step(
                    L.ptr(p), L.ptr(self.flat_grad[off: off + k]), L.ptr(self.m[off: off + k]), L.ptr(self.v[off: off + k]), k,
                    group["lr"], group["betas"][0], group["betas"][1], group["eps"], group["weight_decay"], self.step_count,
                    self.max_grad_norm, L.ptr(self.gnorm_sq), st))
                # the kernel writes through the raw pointer: tell autograd (and ASRModel's operand-refresh check, which keys on
                # the parameters' version counters) that the tensor changed
                torch.autograd.graph.increment_version(p)
        return None
