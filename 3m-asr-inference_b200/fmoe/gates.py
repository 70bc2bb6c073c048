"""Mirror of trainer_3m_fix/fmoe/gates.py:46-66 (NaiveGate). The parameter holder is identical (`gate` = nn.Linear
(d_model, num_expert * world_size)); forward runs the fused gate kernel instead of Linear + topk + softmax."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops


class NaiveGate(nn.Module):
    def __init__(self, d_model, num_expert, world_size, top_k=2):
        super().__init__()
        self.gate = nn.Linear(d_model, num_expert * world_size)
        self.top_k = top_k

    def invalidate_packed(self) -> None:
        self._wr_key = None
        self._wrp_key = None

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self.invalidate_packed()

    def router_params(self):
        """(Wr [d, E_total] fp32 contiguous, br [E_total]) in the layout the kernels take."""
        w = self.gate.weight
        key = (w.data_ptr(), w._version, str(w.device))
        if getattr(self, "_wr_key", None) is None or self._wr_key != key:
            self._wr = w.detach().float().t().contiguous()
            self._wr_key = key
        b = self.gate.bias
        return self._wr, (None if b is None else b.detach().float().contiguous())

    def router_packed(self):
        """bf16 hi/lo packing of the router for the tensor-core gate (None when E > 32)."""
        Wr, _ = self.router_params()
        if Wr.shape[1] > 32 or not Wr.is_cuda:
            return None
        if getattr(self, "_wrp_key", None) is None or self._wrp_key != self._wr_key:
            self._wrp = ops.pack_router(Wr)
            self._wrp_key = self._wr_key
        return self._wrp

    def forward(self, inp):
        """Returns (gate_top_k_idx [N * top_k] int64, gate_score [N, 1, top_k], None).
        The reference also returns the dense logits (gates.py:66); they are not materialised here."""
        Wr, br = self.router_params()
        x = inp.reshape(-1, inp.shape[-1]).contiguous()
        idx, score = ops.gate(x, None, Wr, br, top_k=self.top_k, gate_mode=ops.GATE_NAIVE)
        return idx.view(-1).long(), score.view(-1, 1, self.top_k).to(inp.dtype), None
