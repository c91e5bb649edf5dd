"""LoRA adapters on the frozen Qwen3 decoder (BASELINE config 5; reference: tiny_audio/asr_modeling.py:289-301 ->
peft `LoraConfig(r, lora_alpha, target_modules, lora_dropout, bias="none")` + `get_peft_model`).  peft is not a dependency:
its semantics are restated -- y = W x + (alpha / r) * B(A(x)), A ~ kaiming-uniform(a = sqrt 5), B = 0, dropout 0 -- and the
arithmetic runs inside the CUDA decoder engine (csrc/engine.cu, augmented-K GEMMs).  Parameters are stored stacked over
layers (one tensor per projection kind); `peft_state_dict()` exports them under peft's adapter key names."""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn as nn

from .engine import LORA_PROJS

_ATTN = ("q_proj", "k_proj", "v_proj", "o_proj")


class LoraAdapters(nn.Module):
    def __init__(self, text_config, rank: int = 8, alpha: float = 32.0, target_modules=LORA_PROJS, dropout: float = 0.0):
        super().__init__()
        if dropout != 0.0:
            raise NotImplementedError("lora_dropout != 0 is not supported on the B200 path (the reference default is 0.0)")
        if rank > 8:
            raise NotImplementedError("LoRA rank > 8 needs a wider rank-block layout in engine.py (reference default: 8)")
        c = text_config
        D, Fd = c.hidden_size, c.intermediate_size
        hd = getattr(c, "head_dim", None) or D // c.num_attention_heads
        QD, KD = c.num_attention_heads * hd, c.num_key_value_heads * hd
        io = {"q_proj": (D, QD), "k_proj": (D, KD), "v_proj": (D, KD), "o_proj": (QD, D), "gate_proj": (D, Fd), "up_proj": (D, Fd),
              "down_proj": (Fd, D)}
        self.rank, self.alpha, self.scaling = rank, alpha, alpha / rank
        self.targets = [t for t in LORA_PROJS if t in set(target_modules)]
        Lyr = c.num_hidden_layers
        self.lora_A = nn.ParameterDict()
        self.lora_B = nn.ParameterDict()
        for t in self.targets:
            i, o = io[t]
            a = torch.empty(Lyr, rank, i)
            for l in range(Lyr):
                nn.init.kaiming_uniform_(a[l], a=math.sqrt(5))
            self.lora_A[t] = nn.Parameter(a)
            self.lora_B[t] = nn.Parameter(torch.zeros(Lyr, o, rank))

    def tensors(self):
        """(A dict, B dict) of the stacked parameters, in LORA_PROJS order."""
        return {t: self.lora_A[t] for t in self.targets}, {t: self.lora_B[t] for t in self.targets}

    def peft_state_dict(self) -> Dict[str, torch.Tensor]:
        """Adapter weights under peft's adapter_model.safetensors key names."""
        out = {}
        for t in self.targets:
            block = "self_attn" if t in _ATTN else "mlp"
            for l in range(self.lora_A[t].shape[0]):
                base = f"base_model.model.model.layers.{l}.{block}.{t}"
                out[f"{base}.lora_A.weight"] = self.lora_A[t][l].detach()
                out[f"{base}.lora_B.weight"] = self.lora_B[t][l].detach()
        return out

    def load_peft_state_dict(self, sd: Dict[str, torch.Tensor]):
        with torch.no_grad():
            for t in self.targets:
                block = "self_attn" if t in _ATTN else "mlp"
                for l in range(self.lora_A[t].shape[0]):
                    base = f"base_model.model.model.layers.{l}.{block}.{t}"
                    self.lora_A[t][l].copy_(sd[f"{base}.lora_A.weight"])
                    self.lora_B[t][l].copy_(sd[f"{base}.lora_B.weight"])
