"""3M-ASR checkpoints either side of the MoE path.

What the reference does with a checkpoint, and what replaces it here:

* `builder.py:132-134` -- `torch.load(path, map_location='cpu')` + `model.load_state_dict(...)`: the parameter names and
  shapes of the mirrors in `layer.py` / `fmoe/layers.py` are the reference's, so that call works unchanged.
* `model/conformer_fmoe_localComm_catEmbed_domain_acc_hier.py:259-273` (`load_state_dict_comm`) -- a WHOLE checkpoint
  holds `num_experts * world_size` experts in every `...experts...` tensor; rank r keeps rows
  `[r * num_experts, (r + 1) * num_experts)` and everything else whole  ->  `slice_experts`.
* same file `:236-257` (`state_dict_comm`) -- the inverse for saving: every rank's experts are placed into a zero tensor
  of the whole size and summed over the ranks  ->  `gather_experts` (same all-reduce formulation, any backend).
* `fmoe_expert_plugin.cpp` takes the expert weights as plugin fields in the checkpoint's own `[E, out, in]` layout; the
  kernels here read the same layout as their K-major tensor-core operand, so "packing" is one cast to bf16 (fp32
  rounded to the TF32 grid for `compute = tf32`) plus the hi/lo split of the router  ->  `pack_moe_layers`.

The router (`router_weights [idim + embed_dim, num_experts * world_size]`) is never sliced: every rank routes its own
tokens over ALL experts.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Mapping, Optional

import torch

EXPERT_KEY = "experts"   # the reference's test: `"experts" in key`  (..._hier.py:248, :268)


def slice_experts(whole_state: Mapping[str, torch.Tensor], rank: int, world_size: int,
                  num_experts: int) -> "OrderedDict[str, torch.Tensor]":
    """Rank `rank`'s view of a whole-model state dict (`load_state_dict_comm`, ..._hier.py:259-273): expert tensors
    keep rows [rank * num_experts, (rank + 1) * num_experts), everything else is passed through."""
    if world_size <= 1:
        return OrderedDict(whole_state)
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for k, v in whole_state.items():
        if EXPERT_KEY not in k:
            out[k] = v
            continue
        if v.size(0) != num_experts * world_size:
            raise ValueError(f"{k}: leading dimension {v.size(0)} != num_experts * world_size = "
                             f"{num_experts} * {world_size}")
        out[k] = v[rank * num_experts:(rank + 1) * num_experts]
    return out


def gather_experts(local_state: Mapping[str, torch.Tensor], rank: int, world_size: int, num_experts: int,
                   group=None) -> "OrderedDict[str, torch.Tensor]":
    """Whole-model state dict from every rank's local one (`state_dict_comm`, ..._hier.py:236-257).  Collective: every
    rank of `group` calls it with the same keys in the same order."""
    if world_size <= 1:
        return OrderedDict(local_state)
    import torch.distributed as dist
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for k, v in local_state.items():
        if EXPERT_KEY not in k:
            out[k] = v
            continue
        if v.size(0) != num_experts:
            raise ValueError(f"{k}: leading dimension {v.size(0)} != num_experts = {num_experts}")
        whole = v.new_zeros((num_experts * world_size,) + tuple(v.shape[1:]))
        whole[rank * num_experts:(rank + 1) * num_experts] = v
        dist.all_reduce(whole, group=group)
        out[k] = whole
    return out


def load_state_dict_comm(model: torch.nn.Module, whole_state: Mapping[str, torch.Tensor], rank: int, world_size: int,
                         num_experts: int, strict: bool = True):
    """`Net.load_state_dict_comm` for any module built from the mirrors (..._hier.py:259-273)."""
    return model.load_state_dict(slice_experts(whole_state, rank, world_size, num_experts), strict=strict)


def moe_layer_prefixes(state: Mapping[str, torch.Tensor]):
    """Prefixes of the MoE feed-forward modules found in a state dict, in checkpoint order: every key
    `<prefix>router_weights` whose siblings `<prefix>experts.w_1.weight` / `w_2.weight` exist."""
    out = []
    for k in state:
        if k.endswith("router_weights"):
            p = k[:-len("router_weights")]
            if p + "experts.w_1.weight" in state and p + "experts.w_2.weight" in state:
                out.append(p)
    return out


def pack_moe_layers(state: Mapping[str, torch.Tensor], device, rank: int = 0, world_size: int = 1,
                    num_experts: Optional[int] = None, compute: str = "bf16") -> Dict[str, dict]:
    """One-time device-side packing of every MoE layer of a (whole) checkpoint for direct C-ABI / `ops.moe_layer` use:
    {prefix: {"experts": ops.PackedExperts, "Wr": fp32 [R, E], "Wr_packed": bf16 hi/lo or None, "br": fp32 [E] or None}}.
    With world_size > 1 the experts are rank `rank`'s slice, the router stays whole.  Needs the CUDA library."""
    from . import ops
    dev = torch.device(device)
    out: Dict[str, dict] = {}
    for p in moe_layer_prefixes(state):
        W1, W2 = state[p + "experts.w_1.weight"], state[p + "experts.w_2.weight"]
        b1, b2 = state.get(p + "experts.w_1.bias"), state.get(p + "experts.w_2.bias")
        if world_size > 1:
            n = num_experts if num_experts is not None else W1.size(0) // world_size
            sl = slice(rank * n, (rank + 1) * n)
            W1, W2 = W1[sl], W2[sl]
            b1 = None if b1 is None else b1[sl]
            b2 = None if b2 is None else b2[sl]
        H, D = W1.shape[1], W1.shape[2]
        E = W1.shape[0]
        zeros = lambda n: torch.zeros(E, n, device=dev)
        W1d, W2d = W1.to(dev), W2.to(dev)
        b1d = zeros(H) if b1 is None else b1.to(dev)
        b2d = zeros(D) if b2 is None else b2.to(dev)
        if compute == "tf32":
            experts = ops.fp32_experts(W1d, b1d, W2d, b2d)
        elif compute == "bf16":
            experts = ops.pack_experts(W1d, b1d, W2d, b2d)
        else:
            raise ValueError("compute must be 'bf16' or 'tf32'")
        Wr = state[p + "router_weights"].to(dev).float().contiguous()
        br = state.get(p + "router_bias")
        out[p] = {"experts": experts, "Wr": Wr, "Wr_packed": ops.pack_router(Wr) if Wr.shape[1] <= 32 else None,
                  "br": None if br is None else br.to(dev).float().contiguous()}
    return out
