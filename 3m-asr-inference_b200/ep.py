"""Expert parallelism: experts partitioned contiguously over the ranks of one NVLink/NVSwitch domain, tokens exchanged
with all-to-all (NCCL over NVLink 5 through torch.distributed).

Reference behaviour being reproduced (the only multi-GPU mode of the hot path):
  expert e lives on rank e // num_local_expert, `num_expert` is per worker     trainer_3m_fix/model/..._hier.py:259-273
  counts all-to-all (fmoe_cuda.expert_exchange)                                trainer_3m_fix/fmoe/functions.py:37-44
  rows all-to-all-v (fmoe_cuda.global_scatter / global_gather)                 trainer_3m_fix/fmoe/functions.py:74-80,185-191
  received rows are processed expert-major, then returned and combined locally  functions.py:43-44,194

Data path per layer and rank (all tensors stay on the device; the only host synchronisation is reading the two
split-size vectors, which the reference also does with .cpu() at functions.py:48-50):
  gate over all E_total experts -> stable dispatch by GLOBAL expert id (== sorted by destination rank, then local
  expert) -> all_to_all(counts [W, E_local]) -> all_to_all_v(bf16 rows, 1 KiB per token at D = 512) -> stable dispatch of
  the received rows by LOCAL expert id -> one grouped-GEMM FFN over the local experts -> inverse permutation ->
  all_to_all_v back -> combine (x gate score, x ff_scale, + residual) in token order.
W = 1 degenerates to the single-GPU fused layer.

The arithmetic is pluggable only so that the host-side logic (split sizes, permutations, ordering) can be exercised on
CPU with the gloo backend in the tests; the product backend is the CUDA one and nothing else is ever selected implicitly.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from . import ops


class CudaBackend:
    """The product path: every stage is a call into libb200moe.so."""

    @staticmethod
    def gate(x, embed, Wr, br, x_len, top_k, gate_mode, seq_len, Wr_packed=None):
        E = Wr.shape[1]
        Demb = 0 if embed is None else embed.shape[-1]
        if Wr_packed is not None and ops.gate_tc_usable(x.dtype, x.shape[-1], Demb, E, top_k):
            return ops.gate_tc(x, embed, Wr_packed, E, br, x_len, top_k=top_k, gate_mode=gate_mode, seq_len=seq_len)
        return ops.gate(x, embed, Wr, br, x_len, top_k=top_k, gate_mode=gate_mode, seq_len=seq_len)

    @staticmethod
    def dispatch(x, idx, num_expert, hidden=0):
        d = ops.dispatch(x, idx, num_expert, hidden=hidden)
        return d.counts, d.offsets, d.mapping, d.xbuf

    @staticmethod
    def expert_ffn(xbuf, offsets, experts, act_type, out_dtype):
        return ops.expert_ffn(xbuf, offsets, experts, act_type=act_type, out_dtype=out_dtype)

    @staticmethod
    def combine(ybuf, mapping, score, residual, ff_scale, top_k):
        return ops.combine(ybuf, mapping, score, residual, ff_scale=ff_scale, top_k=top_k)


def split_sizes(send_counts: torch.Tensor, recv_counts: torch.Tensor):
    """[W, E_local] count matrices -> (rows sent to each rank, rows received from each rank) as Python lists."""
    both = torch.stack([send_counts.sum(dim=1), recv_counts.sum(dim=1)]).cpu()
    return both[0].tolist(), both[1].tolist()


def received_expert_ids(recv_counts: torch.Tensor, total: int) -> torch.Tensor:
    """Local expert id of every received row. Rows arrive ordered by source rank, then by local expert."""
    W, E_local = recv_counts.shape
    ids = torch.arange(E_local, device=recv_counts.device, dtype=torch.int32).repeat(W)
    return torch.repeat_interleave(ids, recv_counts.reshape(-1).long(), output_size=total)


def ep_moe_layer(x: torch.Tensor, embed: Optional[torch.Tensor], Wr: torch.Tensor, br: Optional[torch.Tensor], experts,
                 *, num_local_expert: int, group=None, top_k: int = 1, gate_mode: int = ops.GATE_3M,
                 act_type: int = ops.ACT_SILU, ff_scale: float = 1.0, residual: Optional[torch.Tensor] = None,
                 keep_expert_output: bool = False, x_len: Optional[torch.Tensor] = None, seq_len: Optional[int] = None,
                 out: Optional[torch.Tensor] = None, Wr_packed: Optional[torch.Tensor] = None, backend=None,
                 return_routing: bool = False):
    """x [S, D] local tokens; experts = this rank's `num_local_expert` experts (PackedExperts for the CUDA backend).
    Wr [R, num_local_expert * W] is replicated.  Returns out [S, D] (and routing when asked)."""
    be = backend or CudaBackend
    W = dist.get_world_size(group) if dist.is_initialized() else 1
    E_local = num_local_expert
    E_total = E_local * W
    if Wr.shape[1] != E_total:
        raise ValueError(f"router has {Wr.shape[1]} experts, expected {E_local} x {W}")
    S, D = x.shape

    idx, score = be.gate(x, embed, Wr, br, x_len, top_k, gate_mode, seq_len, Wr_packed)
    counts, offsets, mapping, xbuf = be.dispatch(x, idx, E_total)
    n_valid_local = None

    if W == 1:
        recv_rows = xbuf
        recv_counts = counts.view(1, E_local)
        total_recv = None
    else:
        send_counts = counts.view(W, E_local).contiguous()
        recv_counts = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts, group=group)       # functions.py:37-40
        send_splits, recv_splits = split_sizes(send_counts, recv_counts)
        total_recv = int(sum(recv_splits))
        n_valid_local = int(sum(send_splits))
        recv_rows = torch.empty(total_recv, D, dtype=xbuf.dtype, device=xbuf.device)
        dist.all_to_all_single(recv_rows, xbuf[:n_valid_local], recv_splits, send_splits, group=group)  # :74-80

    if W == 1:
        ybuf = be.expert_ffn(xbuf, offsets, experts, act_type, x.dtype)
        y_sorted = ybuf
    else:
        # received rows: [source rank][local expert] -> expert-major (stable, so source-rank order is kept inside an expert)
        idx_recv = received_expert_ids(recv_counts, total_recv).view(-1, 1)
        _c2, offsets2, mapping2, xbuf2 = be.dispatch(recv_rows, idx_recv, E_local)
        ybuf2 = be.expert_ffn(xbuf2, offsets2, experts, act_type, xbuf.dtype)
        y_recv_order = be.combine(ybuf2, mapping2, None, None, 1.0, 1)       # inverse permutation
        y_sorted = torch.empty(S * top_k, D, dtype=y_recv_order.dtype, device=x.device)
        dist.all_to_all_single(y_sorted[:n_valid_local], y_recv_order, send_splits, recv_splits, group=group)  # :185-191
        if y_sorted.dtype != x.dtype:
            y_sorted = y_sorted.to(x.dtype)

    res = be.combine(y_sorted, mapping, None if keep_expert_output else score, residual, ff_scale, top_k)
    if out is not None:
        out.copy_(res.view(out.shape))
        res = out
    if return_routing:
        return res, idx, score, counts, mapping
    return res
