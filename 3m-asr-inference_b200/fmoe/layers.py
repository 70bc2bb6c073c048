"""Mirror of trainer_3m_fix/fmoe/layers.py: FMoELinear (:13-52, parameter holder) and FMoE (:105-210, commented out
upstream).  Parameters keep the reference names and shapes (`weight [E, out, in]`, `bias [E, out]`)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from .gates import NaiveGate


class FMoELinear(nn.Module):
    def __init__(self, num_expert: int, in_feat: int, out_feat: int, bias: bool = True, rank: int = 0):
        super().__init__()
        self.num_expert = num_expert
        self.in_feat = in_feat
        self.out_feat = out_feat
        self.rank = rank
        self.weight = nn.Parameter(torch.Tensor(num_expert, out_feat, in_feat))
        nn.init.xavier_uniform_(self.weight, gain=0.5)
        if bias:
            self.bias = nn.Parameter(torch.Tensor(num_expert, out_feat))
            self.bias.data.fill_(0.0)
        else:
            self.register_parameter("bias", None)

    def forward(self, inp, fwd_expert_count, capacity=-1):
        raise RuntimeError(
            "FMoELinear is a parameter holder: the two expert linears and the activation run as ONE fused grouped "
            "GEMM kernel, driven by the owning module (FMoE / LocalFmoeCatEmbedFeedForward).")

    def extra_repr(self) -> str:
        return "num_expert={}, in_features={}, out_features={}, bias={}, rank={}".format(
            self.num_expert, self.in_feat, self.out_feat, self.bias is not None, self.rank)


def activation_code(act) -> int:
    """Maps the reference's activation modules (trainer_3m_fix/utils/common.py:31-43) to the kernel's act_type."""
    name = type(act).__name__.lower() if not isinstance(act, str) else act.lower()
    if name in ("swish", "silu"):
        return ops.ACT_SILU
    if name == "relu":
        return ops.ACT_RELU
    if name == "gelu":
        return ops.ACT_GELU
    raise NotImplementedError(f"activation {name!r} is not supported by the fused expert kernel (silu/relu/gelu)")


class PackedExpertCache:
    """bf16 copies of the expert weights, refreshed when the parameters change (load_state_dict, .to(), optimiser step).
    The key is (data_ptr, _version, device) per parameter.  Writes through `param.data` (p.data.copy_ / fill_ / add_, the
    idiom of the reference's own FMoELinear init and BMUF sync, utils/fmoe_localComm_bmuf.py) do NOT bump `_version`:
    after such an update call `invalidate()` here, or `invalidate_packed_weights(model)` for a whole model."""

    def __init__(self):
        self._key = None
        self._packed = None

    def invalidate(self) -> None:
        self._key = None
        self._packed = None

    def get(self, w1: FMoELinear, w2: FMoELinear) -> ops.PackedExperts:
        ps = [w1.weight, w1.bias, w2.weight, w2.bias]
        key = tuple((None if p is None else (p.data_ptr(), p._version, str(p.device))) for p in ps)
        if key != self._key:
            self._packed = ops.pack_experts(w1.weight, w1.bias, w2.weight, w2.bias)
            self._key = key
        return self._packed


class FMoE(nn.Module):
    """FMoE(num_expert, d_model, world_size, mp_group, top_k, gate, expert, gate_hook): fmoe/layers.py:123-210.
    `num_expert` is per worker.  Child classes attach `self.experts` exposing two FMoELinear layers + an activation."""

    def __init__(self, num_expert=32, d_model=1024, world_size=1, mp_group=None, top_k=2, gate=NaiveGate, expert=None,
                 gate_hook=None):
        super().__init__()
        if mp_group is not None:
            raise NotImplementedError("mp_group (replicated-input model parallelism) is unused by 3M-ASR; not built")
        if expert is not None:
            raise NotImplementedError("arbitrary expert modules are not fusable; use the built-in two-layer expert")
        self.num_expert = num_expert
        self.d_model = d_model
        self.world_size = world_size
        self.mp_group = None
        self.mp_size = 1
        self.mp_rank = 0
        self.top_k = top_k
        self.gate = gate(d_model, num_expert, world_size, top_k)
        self.experts = None
        self.experts_fused = True
        self.gate_hook = gate_hook
        self._cache = PackedExpertCache()
        self.ep_group = None  # torch.distributed group of the expert-parallel workers (world_size > 1)
        self.ep_capacity = 8192  # tokens per rank per call the expert-parallel receive buffers are sized for

    def invalidate_packed(self) -> None:
        """Drop the bf16 / hi-lo packed copies of the expert and router weights (after an in-place `.data` update)."""
        self._cache.invalidate()
        if hasattr(self.gate, "invalidate_packed"):
            self.gate.invalidate_packed()

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self.invalidate_packed()

    def _expert_layers(self):
        e = self.experts
        for a, b in (("htoh4", "h4toh"), ("w_1", "w_2"), ("hid_proj", "mem_proj")):
            if hasattr(e, a):
                return getattr(e, a), getattr(e, b), e.activation
        raise RuntimeError("self.experts must expose (htoh4, h4toh) or (w_1, w_2) FMoELinear layers")

    def forward(self, inp):
        w1, w2, act = self._expert_layers()
        packed = self._cache.get(w1, w2)
        x = inp.reshape(-1, self.d_model).contiguous()
        Wr, br = self.gate.router_params()
        if self.world_size > 1 and x.dtype == torch.bfloat16:
            # expert parallelism over peer-mapped memory (the product path; context created collectively on first use)
            from .. import ep_p2p
            if getattr(self, "_ep_ctx", None) is None:
                self._ep_ctx = ep_p2p.EpContext.from_process_group(self.num_expert, self.d_model,
                                                                   max(self.ep_capacity, x.shape[0]) * self.top_k,
                                                                   group=self.ep_group)
            if x.shape[0] * self.top_k > self._ep_ctx.cap:
                raise RuntimeError(f"{x.shape[0]} tokens x top-{self.top_k} exceed the expert-parallel capacity "
                                   f"{self._ep_ctx.cap}; set a larger `ep_capacity` on every rank before the first forward")
            out = self._ep_ctx.forward(x, None, Wr, br, packed, top_k=self.top_k, gate_mode=ops.GATE_NAIVE,
                                       act_type=activation_code(act),
                                       Wr_packed=self.gate.router_packed() if hasattr(self.gate, "router_packed") else None)
            return out.reshape(inp.shape)
        if self.world_size > 1:
            from .. import ep   # fp32 / fp16 activations: the NCCL all-to-all formulation
            out = ep.ep_moe_layer(x, None, Wr, br, packed, num_local_expert=self.num_expert, group=self.ep_group,
                                  top_k=self.top_k, gate_mode=ops.GATE_NAIVE, act_type=activation_code(act))
            return out.reshape(inp.shape)
        res = ops.moe_layer(x, None, Wr, br, packed, top_k=self.top_k, gate_mode=ops.GATE_NAIVE,
                            act_type=activation_code(act), ff_scale=1.0, return_routing=self.gate_hook is not None,
                            Wr_packed=self.gate.router_packed() if hasattr(self.gate, "router_packed") else None)
        if self.gate_hook:
            self.gate_hook(res.idx.view(-1).long(), res.score.view(-1, 1, self.top_k), None)
        return res.out.reshape(inp.shape)


def invalidate_packed_weights(model: nn.Module) -> None:
    """Walks `model` and drops every cached packed copy (experts, routers, folded-LayerNorm routers).  Needed after
    weights were changed through `.data` (no autograd version bump); load_state_dict does it on its own."""
    for m in model.modules():
        if hasattr(m, "invalidate_packed"):
            m.invalidate_packed()
