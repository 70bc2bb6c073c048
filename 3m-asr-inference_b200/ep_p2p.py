"""Expert parallelism over peer-mapped GPU memory (the product multi-GPU path; `ep.py` keeps the NCCL formulation).

Reference behaviour being reproduced: experts partitioned contiguously over the GPUs of one node, `num_expert` per
worker (trainer_3m_fix/model/..._hier.py:259-273), tokens exchanged around the expert computation
(trainer_3m_fix/fmoe/functions.py:37-50 counts, :74-80 rows out, :185-191 rows back).  Here the exchange is done by the
kernels themselves: the dispatch kernel tells every rank its counts and stores rows to their final place in the owner
GPU's receive buffer, the expert-FFN kernel stores results on the source GPU, flags with system-scope release/acquire
signal arrival.  torch.distributed is used ONCE, at set-up, to pass the 64-byte CUDA IPC handles around; there is no
collective and no host synchronisation per layer (see include/b200moe.h, "expert parallelism over peer-mapped memory").

Folded combine: pass `out=ctx.out_buffer(slot, S)` (slots 0 / 1 alternately, the same slot on every rank) with
`residual=x` (or None) and top-1 routing, and the owners write the finished output rows straight into that buffer; with
`wait=False` the wait for them is left to the next `forward` on the context (or `ctx.wait()`).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from . import _lib, ops


class _RawDeviceMemory:
    """__cuda_array_interface__ over memory this package allocated with cudaMalloc (the symmetric buffer)."""

    def __init__(self, ptr: int, rows: int, cols: int, owner):
        self._owner = owner   # keeps the context (and with it the allocation) alive
        self.__cuda_array_interface__ = {"shape": (rows, cols), "typestr": "<i2", "data": (ptr, False), "version": 3,
                                         "strides": None}


def _device_view(ptr: int, rows: int, cols: int, device, owner) -> torch.Tensor:
    with torch.cuda.device(device):
        t = torch.as_tensor(_RawDeviceMemory(ptr, rows, cols, owner), device=device)
    return t.view(torch.bfloat16)


class EpContext:
    """One rank's view of the symmetric buffers of all ranks."""

    def __init__(self, rank: int, world: int, num_local_expert: int, d_model: int, cap: int, buffers: List[int],
                 device, timeout_ms: int = 2000, owned_local: Optional[int] = None, opened: Optional[List[int]] = None):
        lib = _lib.load()
        arr = (C.c_void_p * world)(*buffers)
        self._ctx = lib.b200moe_ep_create(rank, world, num_local_expert, d_model, cap, arr, timeout_ms)
        if not self._ctx:
            raise RuntimeError("b200moe_ep_create failed: " + lib.b200moe_last_error().decode())
        self.rank, self.world, self.E_local, self.D, self.cap = rank, world, num_local_expert, d_model, cap
        self.device = torch.device(device)
        self._owned_local = owned_local      # buffer this object cudaMalloc'ed (freed on close)
        self._opened = opened or []          # IPC mappings this object opened
        self._ws = {}
        self._out = {}

    # ---- construction -------------------------------------------------------------------------------------------------
    @staticmethod
    def buffer_bytes(world: int, num_local_expert: int, d_model: int, cap: int) -> int:
        n = int(_lib.load().b200moe_ep_buffer_bytes(world, num_local_expert, d_model, cap))
        if n == 0:
            raise ValueError("unsupported expert-parallel configuration")
        return n

    @classmethod
    def from_process_group(cls, num_local_expert: int, d_model: int, cap: int, group=None, timeout_ms: int = 2000):
        """One process per GPU (torch.cuda.current_device()).  Collective: every rank of `group` must call it."""
        import torch.distributed as dist
        lib = _lib.load()
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        dev = torch.device("cuda", torch.cuda.current_device())
        # the buffers are SYMMETRIC: every rank addresses its peers' buffers with its own layout, so the layout parameters
        # must be identical everywhere (a per-rank capacity silently sends rows to the wrong place)
        shapes = [None] * world
        dist.all_gather_object(shapes, (num_local_expert, d_model, cap), group=group)
        if any(sh != shapes[0] for sh in shapes):
            raise ValueError(f"EpContext: (num_local_expert, d_model, cap) differ across ranks: {shapes}")
        nbytes = cls.buffer_bytes(world, num_local_expert, d_model, cap)
        local = C.c_void_p()
        _lib.check(lib.b200moe_ep_alloc(nbytes, C.byref(local)), "b200moe_ep_alloc")
        handle = C.create_string_buffer(64)
        _lib.check(lib.b200moe_ep_ipc_export(local, handle), "b200moe_ep_ipc_export")
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        bufs, opened = [], []
        for r in range(world):
            if r == rank:
                bufs.append(local.value)
                continue
            peer = C.c_void_p()
            _lib.check(lib.b200moe_ep_ipc_open(C.create_string_buffer(handles[r], 64), C.byref(peer)),
                       "b200moe_ep_ipc_open")
            bufs.append(peer.value)
            opened.append(peer.value)
        ctx = cls(rank, world, num_local_expert, d_model, cap, bufs, dev, timeout_ms, owned_local=local.value,
                  opened=opened)
        dist.barrier(group)   # nobody pushes before every mapping exists
        return ctx

    @classmethod
    def simulate(cls, world: int, num_local_expert: int, d_model: int, cap: int, device, timeout_ms: int = 500):
        """All ranks inside ONE process on ONE GPU (tests): returns `world` contexts over `world` local buffers.  Drive
        them stage by stage (forward(..., stages=1) on every rank, then stages=2, then stages=4): kernels of different
        ranks wait on one another, so they must be queued in dependency order on the single device."""
        nbytes = cls.buffer_bytes(world, num_local_expert, d_model, cap)
        keep = [torch.zeros(nbytes, dtype=torch.uint8, device=device) for _ in range(world)]
        ptrs = [t.data_ptr() for t in keep]
        out = []
        for r in range(world):
            c = cls(r, world, num_local_expert, d_model, cap, ptrs, device, timeout_ms)
            c._keep = keep
            out.append(c)
        return out

    def close(self) -> None:
        lib = _lib.load()
        if self._ctx:
            lib.b200moe_ep_destroy(self._ctx)
            self._ctx = None
        for p in self._opened:
            lib.b200moe_ep_ipc_close(p)
        self._opened = []
        if self._owned_local:
            lib.b200moe_ep_free(self._owned_local)
            self._owned_local = None

    # ---- per layer ------------------------------------------------------------------------------------------------------
    def workspace(self, hidden: int, block: bool = False) -> torch.Tensor:
        key = (hidden, ops._stream(), block)
        ws = self._ws.get(key)
        if ws is None:
            lib = _lib.load()
            n = int(lib.b200moe_ep_block_workspace_bytes(self._ctx, hidden) if block
                    else lib.b200moe_ep_workspace_bytes(self._ctx, hidden))
            ws = torch.empty(max(n, 1), dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        return ws

    def out_buffer(self, slot: int, rows: int) -> torch.Tensor:
        """[rows, D] bf16 view of output slot 0 / 1 inside this rank's symmetric buffer (reachable by the peers)."""
        if rows > self.cap:
            raise ValueError(f"{rows} rows exceed the context capacity {self.cap}")
        base = self._out.get(slot)
        if base is None:
            ptr, nbytes = C.c_void_p(), C.c_size_t()
            _lib.check(_lib.load().b200moe_ep_out_buffer(self._ctx, slot, C.byref(ptr), C.byref(nbytes)),
                       "b200moe_ep_out_buffer")
            base = _device_view(ptr.value, self.cap, self.D, self.device, self)
            self._out[slot] = base
        return base[:rows]

    def wait(self) -> None:
        """Orders the current stream behind the owners' writes of the last folded `forward(..., wait=False)`."""
        _lib.check(_lib.load().b200moe_ep_wait(self._ctx, ops._stream()), "b200moe_ep_wait")

    def status(self) -> int:
        st = C.c_int(0)
        _lib.check(_lib.load().b200moe_ep_status(self._ctx, C.byref(st)), "b200moe_ep_status")
        return st.value

    def forward(self, x: torch.Tensor, embed: Optional[torch.Tensor], Wr: Optional[torch.Tensor],
                br: Optional[torch.Tensor], experts: "ops.PackedExperts", *, residual: Optional[torch.Tensor] = None,
                x_len: Optional[torch.Tensor] = None, seq_len: Optional[int] = None, top_k: int = 1,
                gate_mode: int = ops.GATE_3M, act_type: int = ops.ACT_SILU, ff_scale: float = 1.0,
                keep_expert_output: bool = False, out: Optional[torch.Tensor] = None,
                Wr_packed: Optional[torch.Tensor] = None, return_routing: bool = False, stages: int = 15,
                routing_bufs=None, norm_ff=None, norm_final=None, eps: float = 1e-12, Wr_packed_ln=None,
                wait: bool = True, out_slot: Optional[int] = None):
        """x [S, D] bf16: this rank's tokens.  experts: this rank's `num_local_expert` experts.  Wr [R, E_total].
        norm_ff / norm_final = (gamma, beta) fp32 [D]: the block's LayerNorms either side of the layer (ops.moe_layer)."""
        block = norm_ff is not None or norm_final is not None
        norms = [t for pair in (norm_ff, norm_final) if pair is not None for t in pair]
        if block and stages != 15:
            raise ValueError("the block call runs all stages at once")
        if not wait and stages != 15:
            raise ValueError("wait=False goes with the one-call form")
        dev = ops._need_cuda(x, embed, Wr, br, residual, x_len, experts.W1, experts.b1, experts.W2, experts.b2, out,
                             Wr_packed, Wr_packed_ln, *norms)
        if x.dtype != torch.bfloat16:
            raise TypeError("the expert-parallel path takes bf16 activations")
        S, D = x.shape
        T = seq_len if seq_len is not None else S
        B = S // T if T else 0
        E_local, H, _ = experts.W1.shape
        if E_local != self.E_local:
            raise ValueError(f"got {E_local} local experts, the context was built for {self.E_local}")
        E_total = E_local * self.world
        Demb = 0 if embed is None else embed.shape[-1]
        out_ptr = None
        if out_slot is not None:
            # output slot of the symmetric buffer (folded combine).  The pointer is passed even for a rank without tokens
            # (an empty view has no data pointer): every rank has to decide for or against folding the same way.
            out = self.out_buffer(out_slot, S)
            out_ptr = self._out[out_slot].data_ptr()
        elif out is None:
            out = torch.empty_like(x)
        idx = score = counts = mapping = None
        if return_routing:
            if routing_bufs is not None:
                idx, score, counts, mapping = routing_bufs
            else:
                idx = torch.empty(S, top_k, dtype=torch.int32, device=dev)
                score = torch.empty(S, top_k, dtype=torch.float32, device=dev)
                counts = torch.empty(E_total, dtype=torch.int32, device=dev)
                mapping = torch.empty(S * top_k, dtype=torch.int32, device=dev)
        ws = self.workspace(H, block)
        p = ops._ptr
        a = _lib.LayerArgs(
            x=p(x), embed=p(embed), residual=p(residual), out=out_ptr if out_ptr is not None else p(out), x_len=p(x_len),
            Wr=p(Wr), Wr_packed=p(Wr_packed),
            br=p(br), W1=p(experts.W1), b1=p(experts.b1), W2=p(experts.W2), b2=p(experts.b2), B=B, T=T, D=D, Demb=Demb,
            E=E_total, H=H, top_k=top_k, gate_mode=gate_mode, act_type=act_type, dtype=ops.dtype_code(x),
            keep_expert_output=int(keep_expert_output), ff_scale=float(ff_scale), idx_out=p(idx), score_out=p(score),
            counts_out=p(counts), mapping_out=p(mapping))
        if block:
            b = _lib.BlockArgs(layer=a, norm_ff_gamma=p(norm_ff[0]) if norm_ff else None,
                               norm_ff_beta=p(norm_ff[1]) if norm_ff else None,
                               norm_final_gamma=p(norm_final[0]) if norm_final else None,
                               norm_final_beta=p(norm_final[1]) if norm_final else None, eps=float(eps),
                               Wr_packed_ln=p(Wr_packed_ln) if norm_ff else None)
            _lib.check(_lib.load().b200moe_ep_block_forward(self._ctx, C.byref(b), p(ws), ws.numel(), ops._stream()),
                       "b200moe_ep_block_forward")
        elif not wait:
            _lib.check(_lib.load().b200moe_ep_forward_deferred(self._ctx, C.byref(a), p(ws), ws.numel(), ops._stream()),
                       "b200moe_ep_forward_deferred")
        else:
            _lib.check(_lib.load().b200moe_ep_forward_stages(self._ctx, C.byref(a), p(ws), ws.numel(), stages,
                                                             ops._stream()), "b200moe_ep_forward")
        if return_routing:
            return out, idx, score, counts, mapping
        return out
