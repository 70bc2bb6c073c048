"""Mirror of trainer_3m_fix/fmoe/functions.py (forward halves; inference only): `moe_prepare_forward` :13-52 and the
four autograd Functions `MOEScatter` :55-104, `MOELinear` :107-132, `MOEbiasLinear` :135-165, `MOEGather` :168-210, with
the reference's names, argument order and `.apply` call convention.  Where the reference hands the work to the
un-vendored `fmoe_cuda` extension, these call one C-ABI entry point each (include/b200moe.h, "staged operator surface"):

    fmoe_cuda.local_scatter(inp, pos)        -> b200moe_combine as a plain row gather     out[i] = inp[pos[i]]
    fmoe_cuda.local_gather(buf, pos)         -> b200moe_scatter_rows                       out[pos[i]] = buf[i]
    fmoe_cuda.forward(buf, weight, counts)   -> b200moe_expert_linear (tcgen05 grouped GEMM, bf16 operands)
    fmoe_cuda.expert_exchange / global_scatter / global_gather -> torch.distributed all-to-all (NCCL over NVLink; the
        fused peer-memory exchange of ep_p2p.py is the product path for whole layers and has no staged form)

Differences from the reference, all on the side of fewer host round trips:
  * `moe_prepare_forward` returns DEVICE tensors (the reference `.cpu()`s three of them) and does not synchronise for
    world_size == 1; `pos` is the STABLE argsort (the reference's torch.sort is unstable: any order inside an expert).
  * count arguments may be device or host tensors, int32 or int64.
  * backward is not implemented (this repository is the inference path); calling it raises.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
from torch.autograd import Function

from .. import ops


def _i32(t: torch.Tensor, device) -> torch.Tensor:
    return t.to(device=device, dtype=torch.int32, non_blocking=True).contiguous()


def exchange_expert_counts(local_expert_count: torch.Tensor, num_expert: int, world_size: int, comm=None) -> torch.Tensor:
    """fmoe_cuda.expert_exchange (functions.py:37-40): entry [j * num_expert + e] of the result is how many rows rank j
    sends to this rank's local expert e.  One all-to-all of `num_expert` counts per peer."""
    import torch.distributed as dist
    out = torch.empty_like(local_expert_count)
    dist.all_to_all_single(out, local_expert_count.contiguous(), group=comm)
    return out


def moe_prepare_forward(gate, num_expert, world_size=1, comm=None):
    """gate: 1-d integer tensor, target (global) expert of every entry; num_expert: experts per worker.
    Returns (pos, local_expert_count, global_expert_count, fwd_expert_count, fwd_batch_size) with the reference's meaning
    (functions.py:13-52).  Counts are int64 device tensors.  world_size == 1: no host synchronisation, and
    fwd_batch_size = gate.numel() (exact: the reference's gates route every entry).  world_size > 1: the size of the
    receive buffer has to reach the host, as in the reference (one `.item()`)."""
    idx = gate.reshape(-1).to(torch.int32).contiguous()
    n_global = num_expert * world_size
    p = ops.prepare(idx, n_global, top_k=1)
    pos = p.pos.long()
    local_expert_count = p.counts.long()
    if world_size > 1:
        global_expert_count = exchange_expert_counts(local_expert_count, num_expert, world_size, comm)
        fwd_expert_count = global_expert_count.view(world_size, num_expert).sum(dim=0)
        fwd_batch_size = int(fwd_expert_count.sum().item())
    else:
        global_expert_count = local_expert_count
        fwd_expert_count = local_expert_count
        fwd_batch_size = idx.numel()
    return pos, local_expert_count, global_expert_count, fwd_expert_count, fwd_batch_size


def _split_sizes(local_expert_count, global_expert_count, world_size) -> Tuple[list, list]:
    """Rows sent to / received from each rank (host lists: torch.distributed wants them on the host, and so did the
    reference, whose counts already are CPU tensors here)."""
    both = torch.stack([local_expert_count.reshape(world_size, -1).sum(dim=1),
                        global_expert_count.reshape(world_size, -1).sum(dim=1)]).cpu()
    return both[0].tolist(), both[1].tolist()


def _recv_to_expert_major(global_expert_count, num_expert, world_size, device) -> torch.Tensor:
    """all_to_all_single delivers rows ordered [source rank][local expert]; fmoe_cuda.global_scatter's receive buffer is
    ordered [local expert][source rank] (its loop nest: experts outside, ranks inside).  Returns, for every row of the
    expert-major buffer, the row of the rank-major one."""
    c = global_expert_count.reshape(world_size, num_expert).to(device=device, dtype=torch.long)
    start = (torch.cumsum(c.reshape(-1), 0) - c.reshape(-1)).view(world_size, num_expert)   # rank-major segment starts
    seg_len = c.t().reshape(-1)                                                             # expert-major segment order
    seg_start = start.t().reshape(-1)
    total = int(seg_len.sum())
    seg = torch.repeat_interleave(torch.arange(seg_len.numel(), device=device), seg_len, output_size=total)
    first = torch.cumsum(seg_len, 0) - seg_len
    return seg_start[seg] + (torch.arange(total, device=device) - first[seg])


def _gather_rows(inp: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    """out[i] = inp[index[i]] through b200moe_combine (un-weighted, no residual)."""
    return ops.combine(inp.contiguous(), _i32(index, inp.device), None, None, ff_scale=1.0, top_k=1)


def _no_backward(name):
    raise NotImplementedError(f"{name}.backward: this repository holds the inference path of fast_moe only")


class MOEScatter(Function):
    """Rows [batch] -> expert-contiguous rows; world_size > 1 also exchanges them (functions.py:55-86)."""

    @staticmethod
    def forward(ctx, inp, pos, local_expert_count, global_expert_count, fwd_batch_size, world_size):
        local_input_buf = _gather_rows(inp, pos)                                    # fmoe_cuda.local_scatter :72
        if world_size > 1:                                                          # fmoe_cuda.global_scatter :74-80
            import torch.distributed as dist
            send, recv = _split_sizes(local_expert_count, global_expert_count, world_size)
            rank_major = torch.empty(sum(recv), inp.shape[1], dtype=inp.dtype, device=inp.device)
            dist.all_to_all_single(rank_major, local_input_buf[:sum(send)].contiguous(), recv, send)
            num_expert = global_expert_count.numel() // world_size
            order = _recv_to_expert_major(global_expert_count, num_expert, world_size, inp.device)
            return _gather_rows(rank_major, order) if rank_major.shape[0] else rank_major
        return local_input_buf

    @staticmethod
    def backward(ctx, *grads):
        _no_backward("MOEScatter")


class MOEGather(Function):
    """Expert-contiguous rows -> [batch] order; world_size > 1 first returns them to their sources (functions.py:168-199)."""

    @staticmethod
    def forward(ctx, global_output_buf, pos, local_expert_count, global_expert_count, local_batch_size, world_size):
        if world_size > 1:                                                          # fmoe_cuda.global_gather :185-191
            import torch.distributed as dist
            send, recv = _split_sizes(local_expert_count, global_expert_count, world_size)
            num_expert = global_expert_count.numel() // world_size
            dev = global_output_buf.device
            order = _recv_to_expert_major(global_expert_count, num_expert, world_size, dev)
            rank_major = ops.scatter_rows(global_output_buf.contiguous(), _i32(order, dev), sum(recv)) \
                if global_output_buf.shape[0] else global_output_buf
            local_output_buf = torch.empty(sum(send), global_output_buf.shape[1], dtype=global_output_buf.dtype, device=dev)
            dist.all_to_all_single(local_output_buf, rank_major, send, recv)
        else:
            local_output_buf = global_output_buf
        n = min(local_output_buf.shape[0], pos.numel())
        return ops.scatter_rows(local_output_buf[:n].contiguous(), _i32(pos.reshape(-1)[:n], local_output_buf.device),
                                local_batch_size)                                   # fmoe_cuda.local_gather :194

    @staticmethod
    def backward(ctx, *grads):
        _no_backward("MOEGather")


# bf16 copies of the weights handed to MOELinear / MOEbiasLinear, keyed like fmoe/layers.py:PackedExpertCache
_PACKED: Dict[tuple, torch.Tensor] = {}


def _packed_weight(weight: torch.Tensor) -> torch.Tensor:
    if weight.dtype == torch.bfloat16:
        return weight.detach().contiguous()
    key = (weight.data_ptr(), weight._version, str(weight.device), tuple(weight.shape))
    w = _PACKED.get(key)
    if w is None:
        if len(_PACKED) > 64:
            _PACKED.clear()
        w = ops.pack_bf16(weight.detach().contiguous())
        _PACKED[key] = w
    return w


def invalidate_packed() -> None:
    """Drop the cached bf16 weight copies (after an in-place `.data` update, which bumps no version counter)."""
    _PACKED.clear()


def _offsets(fwd_expert_count: torch.Tensor, device) -> torch.Tensor:
    c = fwd_expert_count.to(device=device, dtype=torch.int64, non_blocking=True)
    off = torch.zeros(c.numel() + 1, dtype=torch.int32, device=device)
    off[1:] = torch.cumsum(c, 0).to(torch.int32)
    return off


def _grouped_linear(global_input_buf, weight, bias, fwd_expert_count, capacity):
    if capacity is not None and capacity > 0:
        raise NotImplementedError("token dropping by capacity is training-time only (capacity = -1 at inference)")
    x = global_input_buf.contiguous()
    xb = x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16)
    b = None if bias is None else bias.detach().float().contiguous()
    out = ops.expert_linear(xb, _offsets(fwd_expert_count, x.device), _packed_weight(weight), b, act_type=ops.ACT_NONE)
    return out if out.dtype == x.dtype else out.to(x.dtype)


class MOELinear(Function):
    """One linear per expert over expert-contiguous rows, no bias (functions.py:107-121); weight [E, out, in]."""

    @staticmethod
    def forward(ctx, global_input_buf, weight, fwd_expert_count, capacity=-1, training=False):
        return _grouped_linear(global_input_buf, weight, None, fwd_expert_count, capacity)

    @staticmethod
    def backward(ctx, *grads):
        _no_backward("MOELinear")


class MOEbiasLinear(Function):
    """MOELinear with bias [E, out] (functions.py:135-152).  The reference appends the bias as an extra weight column and a
    column of ones to the input; here it is added in the GEMM's epilogue."""

    @staticmethod
    def forward(ctx, global_input_buf, weight, bias, fwd_expert_count, capacity=-1, training=False):
        return _grouped_linear(global_input_buf, weight, bias, fwd_expert_count, capacity)

    @staticmethod
    def backward(ctx, *grads):
        _no_backward("MOEbiasLinear")


# ---- round-1 helpers, kept: the fused stage calls with the dispatch record -------------------------------------------
def moe_scatter(inp, idx, num_expert):
    """MOEScatter.forward for world_size 1 from the expert ids: the dispatch record (xbuf = inp rows in expert order, bf16)."""
    return ops.dispatch(inp.contiguous(), idx.to(torch.int32).contiguous(), num_expert)


def moe_gather(ybuf, mapping, top_k=1):
    """MOEGather.forward for world_size 1 from the mapping: out[i] = ybuf[mapping[i]] (un-weighted)."""
    if top_k != 1:
        raise ValueError("moe_gather returns one row per entry; pass the flattened [N * top_k] mapping with top_k=1")
    return ops.combine(ybuf, mapping, None, None, ff_scale=1.0, top_k=1)
