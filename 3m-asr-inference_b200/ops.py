"""PyTorch-facing wrappers over the C ABI (device memory and streams come from torch; the arithmetic does not).

Each function mirrors one stage of the reference path (see include/b200moe.h for file:line citations):
  gate        -> router MatMul + SoftmaxTopKPluginDynamic / NaiveGate.forward
  dispatch    -> moe_prepare_forward + MOEScatter.forward / ScatterMapping(+Copy) kernels
  expert_ffn  -> MOEbiasLinear x2 with the activation / the per-expert cuBLAS loop of FMoEExpertPlugin::enqueue
  combine     -> MOEGather.forward + x gate score (+ x ff_scale + residual)
  moe_layer   -> all of the above behind one call (what the modules in fmoe/ and layer.py use)
Inputs must be CUDA tensors; there is deliberately no CPU implementation.
"""
from __future__ import annotations

from typing import Dict, NamedTuple, Optional, Tuple

import torch

from . import _lib
from ._lib import (ACT_GELU, ACT_NONE, ACT_RELU, ACT_SILU, BF16, COMPUTE_BF16, COMPUTE_TF32, F16, F32, GATE_3M,  # noqa: F401
                   GATE_NAIVE)

_DTYPE_CODE = {torch.float32: F32, torch.float16: F16, torch.bfloat16: BF16}


def dtype_code(t: torch.Tensor) -> int:
    try:
        return _DTYPE_CODE[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported activation dtype {t.dtype}; use float32, float16 or bfloat16") from None


def _need_cuda(*tensors) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("b200moe ops need CUDA tensors: this package has no CPU fallback")
        if not t.is_contiguous():
            raise ValueError("b200moe ops need contiguous tensors")
        dev = t.device if dev is None else dev
        if t.device != dev:
            raise ValueError("all tensors must live on the same device")
    return dev


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def launch_count() -> int:
    return int(_lib.load().b200moe_launch_count())


def config(key: str, value: int) -> None:
    """Run-time tunables ("route", "pdl", "pdl_trig", "prefetch", "ln_fuse"); see include/b200moe.h."""
    _lib.check(_lib.load().b200moe_config(key.encode(), int(value)), "b200moe_config")


def status(clear: bool = True) -> int:
    """Device status word of the kernels launched so far (synchronous read): 0 ok, 1 = a fused gate + dispatch launch
    gave up waiting for its grid (include/b200moe.h, b200moe_status)."""
    import ctypes
    st = ctypes.c_int(0)
    _lib.check(_lib.load().b200moe_status(ctypes.byref(st), int(clear)), "b200moe_status")
    return st.value


def profile_enable(on: bool) -> None:
    _lib.check(_lib.load().b200moe_profile_enable(int(on)), "b200moe_profile_enable")


def profile_read():
    """-> ({stage: total ms}, {stage: calls}) since the last read; stages: gate, dispatch, expert_ffn, combine."""
    import ctypes
    ms = (ctypes.c_float * 4)()
    n = (ctypes.c_int * 4)()
    _lib.check(_lib.load().b200moe_profile_read(ms, n), "b200moe_profile_read")
    names = ("gate", "dispatch", "expert_ffn", "combine")
    return {k: float(ms[i]) for i, k in enumerate(names)}, {k: int(n[i]) for i, k in enumerate(names)}


# ---- workspace cache ------------------------------------------------------------------------------------------------
_WS: Dict[Tuple, torch.Tensor] = {}


def workspace_bytes(S: int, E: int, D: int, H: int, top_k: int) -> int:
    return int(_lib.load().b200moe_workspace_bytes(S, E, D, H, top_k))


def get_workspace(device, S: int, E: int, D: int, H: int, top_k: int, block: bool = False) -> torch.Tensor:
    """One cached scratch buffer per (device, stream, shape). Sized by the library, owned by torch's allocator.
    block=True: sized for b200moe_block_forward (the layer's workspace + the normalised input)."""
    key = (str(device), _stream(), S, E, D, H, top_k, block)
    ws = _WS.get(key)
    if ws is None:
        n = (int(_lib.load().b200moe_block_workspace_bytes(S, E, D, H, top_k)) if block
             else workspace_bytes(S, E, D, H, top_k))
        ws = torch.empty(max(n, 1), dtype=torch.uint8, device=device)
        _WS[key] = ws
    return ws


def clear_workspaces() -> None:
    _WS.clear()


# ---- weights ----------------------------------------------------------------------------------------------------------
def pack_bf16(w: torch.Tensor) -> torch.Tensor:
    """Casts expert weights ([E, out, in], the reference's FMoELinear layout) to bf16 on the device, once."""
    _need_cuda(w)
    if w.dtype == torch.bfloat16:
        return w
    out = torch.empty(w.shape, dtype=torch.bfloat16, device=w.device)
    lib = _lib.load()
    _lib.check(lib.b200moe_pack_bf16(_ptr(w), dtype_code(w), _ptr(out), w.numel(), _stream()), "b200moe_pack_bf16")
    return out


class PackedExperts(NamedTuple):
    W1: torch.Tensor  # [E, H, D] bf16
    b1: Optional[torch.Tensor]  # [E, H] fp32
    W2: torch.Tensor  # [E, D, H] bf16
    b2: Optional[torch.Tensor]  # [E, D] fp32


def pack_experts(W1, b1, W2, b2) -> PackedExperts:
    def f32(b):
        return None if b is None else b.detach().float().contiguous()
    return PackedExperts(pack_bf16(W1.detach().contiguous()), f32(b1), pack_bf16(W2.detach().contiguous()), f32(b2))


def fp32_experts(W1, b1, W2, b2) -> PackedExperts:
    """The reference's fp32 FMoELinear weights ([E, out, in]) rounded once to TF32 values: for compute=COMPUTE_TF32."""
    def f32(t):
        return None if t is None else t.detach().float().contiguous()

    def tf32(t):
        _need_cuda(t)
        src = t.detach().float().contiguous()
        dst = torch.empty_like(src)
        _lib.check(_lib.load().b200moe_pack_tf32(_ptr(src), _ptr(dst), src.numel(), _stream()), "b200moe_pack_tf32")
        return dst
    return PackedExperts(tf32(W1), f32(b1), tf32(W2), f32(b2))


def pack_router(Wr: torch.Tensor) -> torch.Tensor:
    """fp32 router [R, E] (E <= 32) -> bf16 hi/lo K-major [64, R] for the tensor-core gate."""
    _need_cuda(Wr)
    R, E = Wr.shape
    if Wr.dtype != torch.float32:
        raise TypeError("router weights must be fp32")
    out = torch.empty(64, R, dtype=torch.bfloat16, device=Wr.device)
    _lib.check(_lib.load().b200moe_pack_router(_ptr(Wr), R, E, _ptr(out), _stream()), "b200moe_pack_router")
    return out


def pack_router_ln(Wr: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor) -> torch.Tensor:
    """Router packed for the folded norm_ff (b200moe_block_args.Wr_packed_ln): the last D = len(gamma) rows of the fp32
    router [R, E] (the x part of the concat) scaled by gamma, bf16 hi/lo K-major, plus c1 = gamma^T Wr_x, c0 = beta^T Wr_x."""
    _need_cuda(Wr, gamma, beta)
    R, E = Wr.shape
    D = gamma.numel()
    if Wr.dtype != torch.float32 or gamma.dtype != torch.float32 or beta.dtype != torch.float32:
        raise TypeError("router weights, gamma and beta must be fp32")
    lib = _lib.load()
    out = torch.empty(int(lib.b200moe_router_ln_pack_bytes(R)), dtype=torch.uint8, device=Wr.device)
    _lib.check(lib.b200moe_pack_router_ln(_ptr(Wr), R, E, D, _ptr(gamma), _ptr(beta), _ptr(out), _stream()),
               "b200moe_pack_router_ln")
    return out


def gate_tc_usable(x_dtype, D: int, Demb: int, E: int, top_k: int) -> bool:
    return x_dtype == torch.bfloat16 and E <= 32 and top_k <= 8 and D % 64 == 0 and Demb % 64 == 0


# ---- stages ---------------------------------------------------------------------------------------------------------------
def gate(x: torch.Tensor, embed: Optional[torch.Tensor], Wr: torch.Tensor, br: Optional[torch.Tensor] = None,
         x_len: Optional[torch.Tensor] = None, *, top_k: int = 1, gate_mode: int = GATE_3M,
         seq_len: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """x [S, D] (or [B, T, D]), embed likewise or None, Wr [Demb + D, E] fp32. Returns idx int32, score fp32 [S, top_k]."""
    dev = _need_cuda(x, embed, Wr, br, x_len)
    if x.dim() == 3:
        B, T, D = x.shape
    else:
        S, D = x.shape
        T = seq_len if seq_len is not None else S
        B = S // T if T else 0
    S = B * T
    Demb = 0 if embed is None else embed.shape[-1]
    E = Wr.shape[1]
    if Wr.dtype != torch.float32 or Wr.shape[0] != D + Demb:
        raise ValueError(f"Wr must be fp32 [{D + Demb}, E], got {tuple(Wr.shape)} {Wr.dtype}")
    idx = torch.empty(S, top_k, dtype=torch.int32, device=dev)
    score = torch.empty(S, top_k, dtype=torch.float32, device=dev)
    lib = _lib.load()
    _lib.check(lib.b200moe_gate(_ptr(x), _ptr(embed), _ptr(Wr), _ptr(br), _ptr(x_len), B, T, D, Demb, E, top_k,
                                gate_mode, dtype_code(x), _ptr(idx), _ptr(score), _stream()), "b200moe_gate")
    return idx, score


def gate_tc(x: torch.Tensor, embed: Optional[torch.Tensor], Wr_packed: torch.Tensor, num_expert: int,
            br: Optional[torch.Tensor] = None, x_len: Optional[torch.Tensor] = None, *, top_k: int = 1,
            gate_mode: int = GATE_3M, seq_len: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Tensor-core gate (bf16 activations, E <= 32). Same outputs as gate()."""
    dev = _need_cuda(x, embed, Wr_packed, br, x_len)
    if x.dim() == 3:
        B, T, D = x.shape
    else:
        S, D = x.shape
        T = seq_len if seq_len is not None else S
        B = S // T if T else 0
    S = B * T
    Demb = 0 if embed is None else embed.shape[-1]
    idx = torch.empty(S, top_k, dtype=torch.int32, device=dev)
    score = torch.empty(S, top_k, dtype=torch.float32, device=dev)
    _lib.check(_lib.load().b200moe_gate_tc(_ptr(x), _ptr(embed), _ptr(Wr_packed), _ptr(br), _ptr(x_len), B, T, D, Demb,
                                           num_expert, top_k, gate_mode, _ptr(idx), _ptr(score), _stream()),
               "b200moe_gate_tc")
    return idx, score


def softmax_topk(logits: torch.Tensor, mask: Optional[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    """SoftmaxTopKPluginDynamic.enqueue: logits [B, T, E], mask [B] int32 -> value [B, T, 1] (logits dtype), idx [B, T, 1]."""
    dev = _need_cuda(logits, mask)
    B, T, E = logits.shape
    value = torch.empty(B, T, 1, dtype=logits.dtype, device=dev)
    idx = torch.empty(B, T, 1, dtype=torch.int32, device=dev)
    lib = _lib.load()
    _lib.check(lib.b200moe_softmax_topk_enqueue(_ptr(logits), _ptr(mask), B, T, E, dtype_code(logits), _ptr(value),
                                                _ptr(idx), _stream()), "b200moe_softmax_topk_enqueue")
    return value, idx


class Dispatched(NamedTuple):
    counts: torch.Tensor   # [E] int32
    offsets: torch.Tensor  # [E + 1] int32
    mapping: torch.Tensor  # [S * top_k] int32
    xbuf: torch.Tensor     # [S * top_k, D] bf16
    ws: torch.Tensor


def dispatch(x: torch.Tensor, idx: torch.Tensor, num_expert: int, *, hidden: int = 0) -> Dispatched:
    """x [S, D]; idx [S, top_k] (or [S * top_k]) int32. `hidden` sizes the workspace for a following expert_ffn."""
    dev = _need_cuda(x, idx)
    S, D = x.shape
    top_k = idx.numel() // S if S else 1
    if idx.dtype != torch.int32:
        raise TypeError("idx must be int32")
    counts = torch.empty(num_expert, dtype=torch.int32, device=dev)
    offsets = torch.empty(num_expert + 1, dtype=torch.int32, device=dev)
    mapping = torch.empty(S * top_k, dtype=torch.int32, device=dev)
    xbuf = torch.empty(S * top_k, D, dtype=torch.bfloat16, device=dev)
    ws = get_workspace(dev, S, num_expert, D, hidden, top_k)
    lib = _lib.load()
    _lib.check(lib.b200moe_dispatch(_ptr(x), _ptr(idx), S, D, num_expert, top_k, dtype_code(x), _ptr(counts),
                                    _ptr(offsets), _ptr(mapping), _ptr(xbuf), _ptr(ws), _stream()),
               "b200moe_dispatch")
    return Dispatched(counts, offsets, mapping, xbuf, ws)


def expert_ffn(xbuf: torch.Tensor, offsets: torch.Tensor, experts: PackedExperts, *, act_type: int = ACT_SILU,
               out_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    dev = _need_cuda(xbuf, offsets, experts.W1, experts.b1, experts.W2, experts.b2)
    n_rows, D = xbuf.shape
    E, H, D2 = experts.W1.shape
    if D2 != D or xbuf.dtype != torch.bfloat16 or experts.W1.dtype != torch.bfloat16:
        raise ValueError("xbuf must be bf16 [rows, D] and the weights bf16 [E, H, D] / [E, D, H]")
    ybuf = torch.empty(n_rows, D, dtype=out_dtype, device=dev)
    ws = get_workspace(dev, n_rows, E, D, H, 1)
    lib = _lib.load()
    _lib.check(lib.b200moe_expert_ffn(_ptr(xbuf), _ptr(offsets), n_rows, _ptr(experts.W1), _ptr(experts.b1),
                                      _ptr(experts.W2), _ptr(experts.b2), E, D, H, act_type, _DTYPE_CODE[out_dtype],
                                      _ptr(ybuf), _ptr(ws), _stream()), "b200moe_expert_ffn")
    return ybuf


def combine(ybuf: torch.Tensor, mapping: torch.Tensor, score: Optional[torch.Tensor], residual: Optional[torch.Tensor],
            *, ff_scale: float = 1.0, top_k: int = 1) -> torch.Tensor:
    dev = _need_cuda(ybuf, mapping, score, residual)
    D = ybuf.shape[1]
    S = mapping.numel() // top_k
    out = torch.empty(S, D, dtype=ybuf.dtype, device=dev)
    lib = _lib.load()
    _lib.check(lib.b200moe_combine(_ptr(ybuf), _ptr(mapping), _ptr(score), _ptr(residual), float(ff_scale), S, D,
                                   top_k, dtype_code(ybuf), _ptr(out), _stream()), "b200moe_combine")
    return out


class Prepared(NamedTuple):
    counts: torch.Tensor   # [E] int32
    offsets: torch.Tensor  # [E + 1] int32
    mapping: torch.Tensor  # [n] int32: row of entry i in expert order (-1 = not routed)
    pos: torch.Tensor      # [n] int32: entry held by row r (the reference's `pos`: stable argsort of the expert ids)


def prepare(idx: torch.Tensor, num_expert: int, *, top_k: int = 1) -> Prepared:
    """Routing tables only (moe_prepare_forward, trainer_3m_fix/fmoe/functions.py:13-52): no row copy, no host sync.
    idx: [n] int32 target experts, n = S * top_k entries."""
    dev = _need_cuda(idx)
    if idx.dtype != torch.int32:
        raise TypeError("idx must be int32")
    n = idx.numel()
    if n % top_k:
        raise ValueError("idx.numel() must be a multiple of top_k")
    S = n // top_k
    counts = torch.empty(num_expert, dtype=torch.int32, device=dev)
    offsets = torch.empty(num_expert + 1, dtype=torch.int32, device=dev)
    mapping = torch.empty(n, dtype=torch.int32, device=dev)
    pos = torch.empty(n, dtype=torch.int32, device=dev)
    ws = get_workspace(dev, S, num_expert, 8, 0, top_k)
    _lib.check(_lib.load().b200moe_prepare(_ptr(idx), S, num_expert, top_k, _ptr(counts), _ptr(offsets), _ptr(mapping),
                                           _ptr(pos), _ptr(ws), _stream()), "b200moe_prepare")
    return Prepared(counts, offsets, mapping, pos)


def scatter_rows(inp: torch.Tensor, index: torch.Tensor, n_out: int) -> torch.Tensor:
    """out[index[i]] = inp[i] (fmoe_cuda.local_gather, functions.py:194); rows no index points at are zero."""
    dev = _need_cuda(inp, index)
    if index.dtype != torch.int32:
        raise TypeError("index must be int32")
    n, D = inp.shape
    out = torch.zeros(n_out, D, dtype=inp.dtype, device=dev)
    _lib.check(_lib.load().b200moe_scatter_rows(_ptr(inp), _ptr(index), n, n_out, D, dtype_code(inp), _ptr(out),
                                                _stream()), "b200moe_scatter_rows")
    return out


def expert_linear(xbuf: torch.Tensor, offsets: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
                  *, act_type: int = _lib.ACT_NONE) -> torch.Tensor:
    """One grouped linear over expert-contiguous rows (fmoe_cuda.forward: MOELinear / MOEbiasLinear.forward,
    functions.py:107-152): out[r] = act(xbuf[r] . weight[e]^T + bias[e]).  xbuf bf16 [rows, K], weight bf16 [E, N, K],
    bias fp32 [E, N] or None -> bf16 [rows, N]."""
    dev = _need_cuda(xbuf, offsets, weight, bias)
    n_rows, K = xbuf.shape
    E, N, K2 = weight.shape
    if K2 != K or xbuf.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16:
        raise ValueError("xbuf must be bf16 [rows, K] and weight bf16 [E, N, K]")
    if bias is not None and (bias.dtype != torch.float32 or tuple(bias.shape) != (E, N)):
        raise ValueError("bias must be fp32 [E, N]")
    if offsets.dtype != torch.int32 or offsets.numel() != E + 1:
        raise ValueError("offsets must be int32 [E + 1]")
    out = torch.empty(n_rows, N, dtype=torch.bfloat16, device=dev)
    ws = get_workspace(dev, n_rows, E, K, 0, 1)
    _lib.check(_lib.load().b200moe_expert_linear(_ptr(xbuf), _ptr(offsets), n_rows, _ptr(weight), _ptr(bias), E, K, N,
                                                 int(act_type), _ptr(out), _ptr(ws), _stream()),
               "b200moe_expert_linear")
    return out


class LayerOut(NamedTuple):
    out: torch.Tensor
    idx: Optional[torch.Tensor]
    score: Optional[torch.Tensor]
    counts: Optional[torch.Tensor]
    mapping: Optional[torch.Tensor]


def moe_layer(x: torch.Tensor, embed: Optional[torch.Tensor], Wr: torch.Tensor, br: Optional[torch.Tensor],
              experts: PackedExperts, *, residual: Optional[torch.Tensor] = None, x_len: Optional[torch.Tensor] = None,
              seq_len: Optional[int] = None, top_k: int = 1, gate_mode: int = GATE_3M, act_type: int = ACT_SILU,
              ff_scale: float = 1.0, keep_expert_output: bool = False, out: Optional[torch.Tensor] = None,
              return_routing: bool = False, ws: Optional[torch.Tensor] = None,
              Wr_packed: Optional[torch.Tensor] = None, compute: int = COMPUTE_BF16,
              norm_ff: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
              norm_final: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, eps: float = 1e-12,
              Wr_packed_ln: Optional[torch.Tensor] = None) -> LayerOut:
    """The fused layer: out = (residual) + ff_scale * sum_k score_k * FFN_{e_k}(x).  x [S, D] or [B, T, D].
    compute=COMPUTE_TF32: fp32 activations, `experts` from fp32_experts() (tensor cores in TF32, fp32 intermediates).
    norm_ff / norm_final = (gamma, beta) fp32 [D]: the Conformer block's LayerNorms either side of the layer
    (fmoe_transformer.py:144-166):  out = norm_final(residual + ff_scale * MoE(norm_ff(x), embed)).
    Wr_packed_ln = pack_router_ln(Wr, *norm_ff): lets norm_ff be folded into the fused gate + dispatch kernel."""
    block = norm_ff is not None or norm_final is not None
    if Wr_packed_ln is not None:
        R = (0 if embed is None else embed.shape[-1]) + x.shape[-1]
        if norm_ff is None or Wr_packed_ln.dtype != torch.uint8 or \
                Wr_packed_ln.numel() != int(_lib.load().b200moe_router_ln_pack_bytes(R)):
            raise ValueError("Wr_packed_ln must come from pack_router_ln(Wr, *norm_ff) for this router")
    norms = [t for pair in (norm_ff, norm_final) if pair is not None for t in pair]
    for t in norms:
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise TypeError("LayerNorm gamma / beta must be contiguous fp32 tensors")
    dev = _need_cuda(x, embed, Wr, br, residual, x_len, experts.W1, experts.b1, experts.W2, experts.b2, out,
                     Wr_packed, Wr_packed_ln, *norms)
    shape = x.shape
    if x.dim() == 3:
        B, T, D = x.shape
    else:
        S, D = x.shape
        T = seq_len if seq_len is not None else S
        B = S // T if T else 0
    S = B * T
    E, H, _ = experts.W1.shape
    Demb = 0 if embed is None else embed.shape[-1]
    if out is None:
        out = torch.empty(shape, dtype=x.dtype, device=dev)
    if ws is None:
        ws = get_workspace(dev, S, E, D, H, top_k, block)
    idx = score = counts = mapping = None
    if return_routing:
        idx = torch.empty(S, top_k, dtype=torch.int32, device=dev)
        score = torch.empty(S, top_k, dtype=torch.float32, device=dev)
        counts = torch.empty(E, dtype=torch.int32, device=dev)
        mapping = torch.empty(S * top_k, dtype=torch.int32, device=dev)
    a = _lib.LayerArgs(
        x=_ptr(x), embed=_ptr(embed), residual=_ptr(residual), out=_ptr(out), x_len=_ptr(x_len),
        Wr=_ptr(Wr), Wr_packed=_ptr(Wr_packed), br=_ptr(br), W1=_ptr(experts.W1), b1=_ptr(experts.b1), W2=_ptr(experts.W2), b2=_ptr(experts.b2),
        B=B, T=T, D=D, Demb=Demb, E=E, H=H, top_k=top_k, gate_mode=gate_mode, act_type=act_type,
        dtype=dtype_code(x), keep_expert_output=int(keep_expert_output), ff_scale=float(ff_scale),
        idx_out=_ptr(idx), score_out=_ptr(score), counts_out=_ptr(counts), mapping_out=_ptr(mapping),
        compute=int(compute))
    if compute == COMPUTE_TF32 and (x.dtype != torch.float32 or experts.W1.dtype != torch.float32):
        raise TypeError("TF32 compute takes fp32 activations and fp32 expert weights (ops.fp32_experts)")
    if compute == COMPUTE_BF16 and experts.W1.dtype != torch.bfloat16:
        raise TypeError("bf16 compute takes bf16-packed expert weights (ops.pack_experts)")
    lib = _lib.load()
    import ctypes
    if block:
        b = _lib.BlockArgs(layer=a, norm_ff_gamma=_ptr(norm_ff[0]) if norm_ff else None,
                           norm_ff_beta=_ptr(norm_ff[1]) if norm_ff else None,
                           norm_final_gamma=_ptr(norm_final[0]) if norm_final else None,
                           norm_final_beta=_ptr(norm_final[1]) if norm_final else None, eps=float(eps),
                           Wr_packed_ln=_ptr(Wr_packed_ln) if norm_ff else None)
        _lib.check(lib.b200moe_block_forward(ctypes.byref(b), _ptr(ws), ws.numel(), _stream()),
                   "b200moe_block_forward")
    else:
        _lib.check(lib.b200moe_forward(ctypes.byref(a), _ptr(ws), ws.numel(), _stream()), "b200moe_forward")
    return LayerOut(out, idx, score, counts, mapping)


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-12,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Row LayerNorm over the last dimension (LayerNormPluginDynamic's job, in the activation dtype); out may be x."""
    dev = _need_cuda(x, gamma, beta, out)
    D = x.shape[-1]
    S = x.numel() // D if D else 0
    if not x.is_contiguous() or gamma.dtype != torch.float32 or beta.dtype != torch.float32:
        raise TypeError("layernorm takes a contiguous input and fp32 gamma / beta")
    if out is None:
        out = torch.empty_like(x)
    _lib.check(_lib.load().b200moe_layernorm(_ptr(x), _ptr(gamma), _ptr(beta), float(eps), S, D, dtype_code(x),
                                             _ptr(out), _stream()), "b200moe_layernorm")
    return out


# ---- the other encoder plugins ------------------------------------------------------------------------------
def att_masked_softmax(scores: torch.Tensor, mask: Optional[torch.Tensor], scale: float) -> torch.Tensor:
    """AttMaskedSoftmaxPluginDynamic: scores [B, N, S, ld], mask [B] int32 valid keys -> softmax(scale * scores), 0 behind."""
    _need_cuda(scores, mask)
    B, N, S, ld = scores.shape
    out = torch.empty_like(scores)
    _lib.check(_lib.load().b200moe_att_masked_softmax(_ptr(scores), _ptr(mask), float(scale), B, N, S, ld,
                                                      dtype_code(scores), _ptr(out), _stream()), "b200moe_att_masked_softmax")
    return out


def glu(x: torch.Tensor) -> torch.Tensor:
    """GluPluginDynamic over dim 1: x [M, 2C, N] -> [M, C, N]."""
    _need_cuda(x)
    M, C2, N = x.shape
    if C2 % 2:
        raise ValueError("GLU needs an even number of channels")
    y = torch.empty(M, C2 // 2, N, dtype=x.dtype, device=x.device)
    _lib.check(_lib.load().b200moe_glu(_ptr(x), M, C2 // 2, N, dtype_code(x), _ptr(y), _stream()), "b200moe_glu")
    return y


def masked_fill(x: torch.Tensor, mask: torch.Tensor, fill: float = 0.0) -> torch.Tensor:
    """MaskedFillPluginDynamic: x [B, dim, T], mask [B] int32 valid lengths; t >= mask[b] -> fill."""
    _need_cuda(x, mask)
    B, dim, T = x.shape
    out = torch.empty_like(x)
    _lib.check(_lib.load().b200moe_masked_fill(_ptr(x), _ptr(mask), float(fill), B, dim, T, dtype_code(x), _ptr(out),
                                               _stream()), "b200moe_masked_fill")
    return out


def rel_pos_encoding(x: torch.Tensor, pe: torch.Tensor, scale: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """RelPositionalEncodingPluginDynamic: x [B, T, D], pe [max_len, D] -> (x * scale, pe[:T])."""
    _need_cuda(x, pe)
    B, T, D = x.shape
    if pe.dtype != x.dtype or pe.shape[-1] != D or pe.shape[-2] < T:
        raise ValueError("pe must be [max_len >= T, D] in the activation dtype")
    out = torch.empty_like(x)
    pos = torch.empty(T, D, dtype=x.dtype, device=x.device)
    _lib.check(_lib.load().b200moe_rel_pos_encoding(_ptr(x), _ptr(pe), float(scale), B, T, D, dtype_code(x), _ptr(out),
                                                    _ptr(pos), _stream()), "b200moe_rel_pos_encoding")
    return out, pos


# ---- torch.library registration ---------------------------------------------------------------------------------------------
# The layer as a registered PyTorch operator, `torch.ops.b200moe.fmoe_forward`: the op the reference's FMoE.forward /
# LocalFmoeCatEmbedFeedForward.forward boil down to (trainer_3m_fix/fmoe/layers.py:186-210, positionwise_feed_forward.py:
# 209-265), so that it can sit in exported / compiled graphs like any ATen op.  Inference only (no autograd formula).
@torch.library.custom_op("b200moe::fmoe_forward", mutates_args=(), device_types="cuda")
def fmoe_forward(x: torch.Tensor, embed: Optional[torch.Tensor], router_weight: torch.Tensor,
                 router_bias: Optional[torch.Tensor], w1: torch.Tensor, b1: Optional[torch.Tensor], w2: torch.Tensor,
                 b2: Optional[torch.Tensor], residual: Optional[torch.Tensor], top_k: int, gate_mode: int, act_type: int,
                 ff_scale: float) -> torch.Tensor:
    """x [S, D] (or [B, T, D]); w1 [E, H, D] / w2 [E, D, H] bf16-packed (pack_experts); router_weight [Demb + D, E] fp32."""
    Wrp = pack_router(router_weight) if gate_tc_usable(x.dtype, x.shape[-1], 0 if embed is None else embed.shape[-1],
                                                       router_weight.shape[1], top_k) else None
    return moe_layer(x.contiguous(), None if embed is None else embed.contiguous(), router_weight, router_bias,
                     PackedExperts(w1, b1, w2, b2), residual=residual, top_k=top_k, gate_mode=gate_mode,
                     act_type=act_type, ff_scale=ff_scale, Wr_packed=Wrp).out


@fmoe_forward.register_fake
def _fmoe_forward_fake(x, embed, router_weight, router_bias, w1, b1, w2, b2, residual, top_k, gate_mode, act_type,
                       ff_scale):
    return torch.empty_like(x)
