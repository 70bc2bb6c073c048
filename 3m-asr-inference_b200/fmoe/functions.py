"""Mirror of trainer_3m_fix/fmoe/functions.py (forward halves): moe_prepare_forward :13-52, MOEScatter :55-86,
MOEGather :168-199, local-expert linear :135-152.  Device-resident: unlike the reference, counts are NOT copied to the
CPU (no host synchronisation) unless the caller asks with `.cpu()`."""
from __future__ import annotations

import torch

from .. import ops


def moe_prepare_forward(gate, num_expert, world_size=1, comm=None):
    """gate: [N * top_k] integer tensor of target experts. Returns (pos, local_expert_count, global_expert_count,
    fwd_expert_count, fwd_batch_size) with the reference's meaning; `pos` is the STABLE sort permutation."""
    if world_size != 1:
        raise NotImplementedError("use b200 ep.ep_moe_layer for world_size > 1")
    idx = gate.reshape(-1).to(torch.int32).contiguous()
    n = idx.numel()
    dummy = torch.empty(n, 8, dtype=torch.bfloat16, device=idx.device)
    d = ops.dispatch(dummy, idx.view(n, 1), num_expert)
    counts = d.counts.long()
    # pos = mapping^-1 restricted to routed entries
    valid = d.mapping >= 0
    pos = torch.empty(int(valid.sum()), dtype=torch.long, device=idx.device)
    pos[d.mapping[valid].long()] = torch.nonzero(valid).flatten()
    return pos, counts, counts, counts, int(counts.sum())


def moe_scatter(inp, idx, num_expert):
    """MOEScatter.forward for world_size 1: returns the dispatch record (xbuf = inp rows in expert order, bf16)."""
    return ops.dispatch(inp.contiguous(), idx.to(torch.int32).contiguous(), num_expert)


def moe_gather(ybuf, mapping, top_k=1):
    """MOEGather.forward for world_size 1: out[i] = ybuf[mapping[i]] (un-weighted)."""
    if top_k != 1:
        raise ValueError("moe_gather returns one row per entry; pass the flattened [N * top_k] mapping with top_k=1")
    return ops.combine(ybuf, mapping, None, None, ff_scale=1.0, top_k=1)
