"""ctypes binding of libb200moe.so (the C ABI in include/b200moe.h).

There is no fallback: if the library has not been built, importing anything that needs it raises, and every compute
call returns an error code that is turned into a RuntimeError carrying b200moe_last_error().
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200moe.so")

F32, F16, BF16 = 0, 1, 2
ACT_SILU, ACT_RELU, ACT_GELU, ACT_NONE = 0, 1, 2, 3   # ACT_NONE: b200moe_expert_linear only
GATE_3M, GATE_NAIVE = 0, 1
COMPUTE_BF16, COMPUTE_TF32 = 0, 1

_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t


class LayerArgs(C.Structure):
    """struct b200moe_layer_args"""
    _fields_ = [
        ("x", _vp), ("embed", _vp), ("residual", _vp), ("out", _vp), ("x_len", _vp),
        ("Wr", _vp), ("Wr_packed", _vp), ("br", _vp),
        ("W1", _vp), ("b1", _vp), ("W2", _vp), ("b2", _vp),
        ("B", _i), ("T", _i), ("D", _i), ("Demb", _i), ("E", _i), ("H", _i), ("top_k", _i),
        ("gate_mode", _i), ("act_type", _i), ("dtype", _i),
        ("keep_expert_output", _i), ("ff_scale", _f),
        ("idx_out", _vp), ("score_out", _vp), ("counts_out", _vp), ("mapping_out", _vp),
        ("compute", _i),
    ]


class BlockArgs(C.Structure):
    """struct b200moe_block_args"""
    _fields_ = [("layer", LayerArgs), ("norm_ff_gamma", _vp), ("norm_ff_beta", _vp), ("norm_final_gamma", _vp),
                ("norm_final_beta", _vp), ("eps", _f), ("Wr_packed_ln", _vp)]


# name -> (restype, argtypes); every symbol include/b200moe.h declares
SIGNATURES = {
    "b200moe_last_error": (C.c_char_p, []),
    "b200moe_status": (_i, [C.POINTER(C.c_int), _i]),
    "b200moe_version": (_i, []),
    "b200moe_device_supported": (_i, [_i]),
    "b200moe_launch_count": (C.c_ulonglong, []),
    "b200moe_debug_ffn_trace": (_i, [_vp, _i]),
    "b200moe_debug_route_trace": (_i, [_vp]),
    "b200moe_debug_timeline": (_i, [_vp, _i]),
    "b200moe_debug_timeline_kind": (_i, [_i]),
    "b200moe_config": (_i, [C.c_char_p, _i]),
    "b200moe_profile_enable": (_i, [_i]),
    "b200moe_profile_read": (_i, [C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "b200moe_pack_bf16": (_i, [_vp, _i, _vp, _sz, _vp]),
    "b200moe_pack_tf32": (_i, [_vp, _vp, _sz, _vp]),
    "b200moe_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "b200moe_gate": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "b200moe_router_pack_bytes": (_sz, [_i]),
    "b200moe_pack_router": (_i, [_vp, _i, _i, _vp, _vp]),
    "b200moe_gate_tc": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "b200moe_dispatch": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200moe_expert_ffn": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "b200moe_combine": (_i, [_vp, _vp, _vp, _vp, _f, _i, _i, _i, _i, _vp, _vp]),
    "b200moe_prepare": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200moe_scatter_rows": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "b200moe_expert_linear": (_i, [_vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "b200moe_forward": (_i, [C.POINTER(LayerArgs), _vp, _sz, _vp]),
    "b200moe_ep_buffer_bytes": (_sz, [_i, _i, _i, _i]),
    "b200moe_ep_alloc": (_i, [_sz, C.POINTER(_vp)]),
    "b200moe_ep_free": (_i, [_vp]),
    "b200moe_ep_ipc_export": (_i, [_vp, _vp]),
    "b200moe_ep_ipc_open": (_i, [_vp, C.POINTER(_vp)]),
    "b200moe_ep_ipc_close": (_i, [_vp]),
    "b200moe_ep_create": (_vp, [_i, _i, _i, _i, _i, C.POINTER(_vp), _i]),
    "b200moe_ep_destroy": (None, [_vp]),
    "b200moe_ep_workspace_bytes": (_sz, [_vp, _i]),
    "b200moe_ep_forward": (_i, [_vp, C.POINTER(LayerArgs), _vp, _sz, _vp]),
    "b200moe_ep_forward_stages": (_i, [_vp, C.POINTER(LayerArgs), _vp, _sz, _i, _vp]),
    "b200moe_ep_forward_deferred": (_i, [_vp, C.POINTER(LayerArgs), _vp, _sz, _vp]),
    "b200moe_ep_wait": (_i, [_vp, _vp]),
    "b200moe_ep_out_buffer": (_i, [_vp, _i, C.POINTER(_vp), C.POINTER(_sz)]),
    "b200moe_ep_status": (_i, [_vp, C.POINTER(_i)]),
    "b200moe_router_ln_pack_bytes": (_sz, [_i]),
    "b200moe_pack_router_ln": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "b200moe_block_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "b200moe_block_forward": (_i, [C.POINTER(BlockArgs), _vp, _sz, _vp]),
    "b200moe_ep_block_workspace_bytes": (_sz, [_vp, _i]),
    "b200moe_ep_block_forward": (_i, [_vp, C.POINTER(BlockArgs), _vp, _sz, _vp]),
    "b200moe_layernorm": (_i, [_vp, _vp, _vp, _f, _i, _i, _i, _vp, _vp]),
    "b200moe_att_masked_softmax": (_i, [_vp, _vp, _f, _i, _i, _i, _i, _i, _vp, _vp]),
    "b200moe_glu": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "b200moe_masked_fill": (_i, [_vp, _vp, _f, _i, _i, _i, _i, _vp, _vp]),
    "b200moe_rel_pos_encoding": (_i, [_vp, _vp, _f, _i, _i, _i, _i, _vp, _vp, _vp]),
    "b200moe_plugin_create": (_vp, [_i, _i, _i, _i, _i]),
    "b200moe_plugin_clone": (_vp, [_vp]),
    "b200moe_plugin_serialization_size": (_sz, [_vp]),
    "b200moe_plugin_serialize": (_i, [_vp, _vp]),
    "b200moe_plugin_deserialize": (_vp, [_vp, _sz]),
    "b200moe_plugin_destroy": (None, [_vp]),
    "b200moe_plugin_workspace_bytes": (_sz, [_vp, _i]),
    "b200moe_plugin_invalidate": (_i, [_vp]),
    "b200moe_plugin_enqueue": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _sz, _vp]),
    "b200moe_softmax_topk_enqueue": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
}

_lib = None


def load() -> C.CDLL:
    """Loads the shared library (once) and attaches the prototypes. Raises if it is missing: there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
            f"`make -C {os.path.join(_HERE, 'csrc')}`. This package has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        if os.environ.get("B200MOE_AB_OLD_LIB") and not hasattr(lib, name):
            continue             # (tools/gpu_ab_libs.sh only: an older build of the library without the newest symbols)
        fn = getattr(lib, name)  # AttributeError here == header and library out of sync
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().b200moe_last_error()
        raise RuntimeError(f"{what} failed with status {status}: {msg.decode() if msg else '?'}")
