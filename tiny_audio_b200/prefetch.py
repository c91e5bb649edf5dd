"""Host -> device staging of collated batches on a side stream, double buffered.

The reference trains through HF Trainer, whose DataLoader (pin_memory, scripts/train.py:467-480 / TrainingArguments
`dataloader_pin_memory`) hands pinned host batches to accelerate, which copies each one to the GPU right before the step -- on the
compute stream, so a 61 MB waveform batch (32 x 30 s of fp32 samples) costs 1.2 ms of PCIe time in front of every step.  This
iterator issues the copy of batch i+1 on its own stream while batch i is being computed; the tensors it yields are device
tensors that `ASRModel.forward` takes as they are (labels stay on the host: the model builds its labelled-row list from them
without a device sync).

    for batch in DevicePrefetcher(dataloader, device="cuda"):
        loss = model(**batch).loss
        ...
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Sequence

import torch

DEVICE_KEYS = ("input_features", "input_ids", "attention_mask", "audio_token_counts", "audio_attention_mask")


class DevicePrefetcher:
    def __init__(self, batches: Iterable[Dict], device="cuda", device_keys: Sequence[str] = DEVICE_KEYS):
        self.it = iter(batches)
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise ValueError("DevicePrefetcher stages batches onto a CUDA device")
        self.keys = set(device_keys)
        self.stream = torch.cuda.Stream(self.dev)
        self.buffers = [dict(), dict()]            # persistent device buffers of the two slots (re-used while shapes repeat)
        self.free_ev: list = [None, None]          # compute stream is done with the slot (recorded when its successor is handed out)
        self.n = 0
        self.pending = None                        # (slot, batch, copy-done event)
        self.prev_slot: Optional[int] = None
        self._stage()

    def _stage(self):
        try:
            b = next(self.it)
        except StopIteration:
            self.pending = None
            return
        k = self.n & 1
        self.n += 1
        out = {}
        with torch.cuda.stream(self.stream):
            if self.free_ev[k] is not None:
                self.stream.wait_event(self.free_ev[k])      # the step that used this slot two batches ago has finished
            for key, v in b.items():
                if torch.is_tensor(v) and key in self.keys and not v.is_cuda:
                    buf = self.buffers[k].get(key)
                    if buf is None or buf.shape != v.shape or buf.dtype != v.dtype:
                        buf = torch.empty(v.shape, dtype=v.dtype, device=self.dev)
                        self.buffers[k][key] = buf
                    buf.copy_(v, non_blocking=True)           # asynchronous only from pinned memory
                    out[key] = buf
                else:
                    out[key] = v
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.pending = (k, out, ev)

    def __iter__(self):
        return self

    def __next__(self) -> Dict:
        if self.pending is None:
            raise StopIteration
        k, out, ev = self.pending
        cur = torch.cuda.current_stream(self.dev)
        if self.prev_slot is not None:             # everything the consumer enqueued for the previous batch is in front of this point
            e = torch.cuda.Event()
            e.record(cur)
            self.free_ev[self.prev_slot] = e
        cur.wait_event(ev)
        self.prev_slot = k
        self._stage()                              # next batch's copy runs under this batch's step
        return out
