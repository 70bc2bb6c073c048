"""Mirror of trainer_3m_fix/fmoe/transformer.py:12-75 (_Expert, FMoETransformerMLP)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .gates import NaiveGate
from .layers import FMoE, FMoELinear


class _Expert(nn.Module):
    def __init__(self, num_expert, d_model, d_hidden, activation, rank=0):
        super().__init__()
        self.htoh4 = FMoELinear(num_expert, d_model, d_hidden, bias=True, rank=rank)
        self.h4toh = FMoELinear(num_expert, d_hidden, d_model, bias=True, rank=rank)
        self.activation = activation


class FMoETransformerMLP(FMoE):
    def __init__(self, num_expert=32, d_model=1024, d_hidden=4096, world_size=1, mp_group=None,
                 activation=torch.nn.GELU(), gate=NaiveGate, top_k=2, expert_dp_comm="none", gate_hook=None):
        super().__init__(num_expert=num_expert, d_model=d_model, gate=gate, top_k=top_k, world_size=world_size,
                         mp_group=mp_group, gate_hook=gate_hook)
        self.experts = _Expert(num_expert, d_model, d_hidden, activation, rank=self.mp_rank)

    def forward(self, inp: torch.Tensor):
        original_shape = inp.shape
        inp = inp.reshape(-1, self.d_model)
        output = super().forward(inp)
        return output.reshape(original_shape)
