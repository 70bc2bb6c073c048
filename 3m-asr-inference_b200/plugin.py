"""Python mirror of the TensorRT plugin surface that builder.py / infer.py drive through network_helper:

  FMoEExpertPluginDynamic   v1  TRTAPI++/plugin/fmoe_expert_plugin/fmoe_expert_plugin.{h,cpp}
  SoftmaxTopKPluginDynamic  v1  TRTAPI++/plugin/softmax_topk_plugin/softmax_topk_plugin.cpp
  AttMaskedSoftmaxPluginDynamic, GluPluginDynamic, MaskedFillPluginDynamic, RelPositionalEncodingPluginDynamic,
  LayerNormPluginDynamic  v1   TRTAPI++/plugin/{att_masked_softmax,glu,masked_fill,rel_positional_encoding,layer_norm}_plugin

Same plugin names / versions, same creator field names (`data_type`, `num_expert`, `idim`, `hidden_units`, optional
`act_type`), same six inputs in the same order, same single un-weighted output, same 32-byte serialisation
(fmoe_expert_plugin.cpp:288-304).  TensorRT itself is absent from this image; INTEGRATION.md shows the
IPluginV2DynamicExt shim a maintainer adds to register these under TensorRT.
"""
from __future__ import annotations

import ctypes
import struct
from typing import Dict, Optional

import torch

from . import _lib, ops

FMOE_EXPERT_NAME = "FMoEExpertPluginDynamic"
FMOE_EXPERT_VERSION = "1"
SOFTMAX_TOPK_NAME = "SoftmaxTopKPluginDynamic"
SOFTMAX_TOPK_VERSION = "1"

_TORCH_DTYPE = {0: torch.float32, 1: torch.float16, 2: torch.bfloat16}


class FMoEExpertPlugin:
    """Owns a b200moe_plugin handle. enqueue(inputs=[input, gate_idx, w1_w, w1_b, w2_w, w2_b]) -> output."""

    def __init__(self, handle: int):
        if not handle:
            raise RuntimeError("plugin creation failed: " + (_lib.load().b200moe_last_error() or b"").decode())
        self._h = handle
        self._ws: Optional[torch.Tensor] = None

    # -- IPluginV2 identity ------------------------------------------------------------------------------------------
    def get_plugin_type(self) -> str:
        return FMOE_EXPERT_NAME

    def get_plugin_version(self) -> str:
        return FMOE_EXPERT_VERSION

    def get_nb_outputs(self) -> int:
        return 1

    # -- lifetime ----------------------------------------------------------------------------------------------------
    def clone(self) -> "FMoEExpertPlugin":
        return FMoEExpertPlugin(_lib.load().b200moe_plugin_clone(self._h))

    def destroy(self) -> None:
        if self._h:
            _lib.load().b200moe_plugin_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    # -- serialisation --------------------------------------------------------------------------------------------------
    def get_serialization_size(self) -> int:
        return int(_lib.load().b200moe_plugin_serialization_size(self._h))

    def serialize(self) -> bytes:
        buf = ctypes.create_string_buffer(self.get_serialization_size())
        _lib.check(_lib.load().b200moe_plugin_serialize(self._h, buf), "b200moe_plugin_serialize")
        return buf.raw

    @property
    def fields(self) -> Dict[str, int]:
        v = struct.unpack("8i", self.serialize())
        return {"data_type": v[0], "num_expert": v[1], "idim": v[2], "hidden_units": v[3], "act_type": v[4]}

    def invalidate(self) -> None:
        """Weights were updated in place behind the same pointers: the next enqueue re-packs them."""
        _lib.check(_lib.load().b200moe_plugin_invalidate(self._h), "b200moe_plugin_invalidate")

    # -- execution ------------------------------------------------------------------------------------------------------
    def get_workspace_size(self, S: int) -> int:
        return int(_lib.load().b200moe_plugin_workspace_bytes(self._h, S))

    def enqueue(self, inputs, workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
        x, gate_idx, w1w, w1b, w2w, w2b = inputs
        f = self.fields
        dt = _TORCH_DTYPE[f["data_type"]]
        for t in (x, w1w, w1b, w2w, w2b):
            if t.dtype != dt:
                raise TypeError(f"plugin data_type is {dt}, got {t.dtype}")  # supportsFormatCombination
            if not t.is_cuda or not t.is_contiguous():
                raise RuntimeError("plugin inputs must be contiguous CUDA tensors")
        if gate_idx.dtype != torch.int32:
            raise TypeError("gate_idx must be int32")
        S = x.numel() // f["idim"]
        need = self.get_workspace_size(S)
        if workspace is None:
            if self._ws is None or self._ws.numel() < need or self._ws.device != x.device:
                self._ws = torch.empty(need, dtype=torch.uint8, device=x.device)
            workspace = self._ws
        out = torch.empty_like(x)
        st = _lib.load().b200moe_plugin_enqueue(
            self._h, x.data_ptr(), gate_idx.data_ptr(), w1w.data_ptr(), w1b.data_ptr(), w2w.data_ptr(),
            w2b.data_ptr(), S, out.data_ptr(), workspace.data_ptr(), workspace.numel(),
            torch.cuda.current_stream().cuda_stream)
        _lib.check(st, "b200moe_plugin_enqueue")
        return out


class FMoEExpertPluginCreator:
    """create_plugin(name, fields) with the reference's field names; returns None on a bad data_type like the
    reference creator does (fmoe_expert_plugin.cpp:360-363)."""
    name = FMOE_EXPERT_NAME
    version = FMOE_EXPERT_VERSION
    field_names = ("data_type", "num_expert", "idim", "hidden_units", "act_type")

    def create_plugin(self, name: str, fields: Dict[str, int]) -> Optional[FMoEExpertPlugin]:
        h = _lib.load().b200moe_plugin_create(int(fields.get("data_type", -1)), int(fields.get("num_expert", 0)),
                                              int(fields.get("idim", 0)), int(fields.get("hidden_units", 0)),
                                              int(fields.get("act_type", 0)))
        return FMoEExpertPlugin(h) if h else None

    def deserialize_plugin(self, name: str, data: bytes) -> Optional[FMoEExpertPlugin]:
        h = _lib.load().b200moe_plugin_deserialize(data, len(data))
        return FMoEExpertPlugin(h) if h else None


class SoftmaxTopKPlugin:
    """enqueue([logits [B,T,E], mask [B] int32]) -> (value [B,T,1], idx [B,T,1] int32); top-1 only."""

    def __init__(self, data_type: int = 0):
        if data_type not in _TORCH_DTYPE:
            raise ValueError("invalid data_type")
        self.data_type = data_type

    def get_plugin_type(self) -> str:
        return SOFTMAX_TOPK_NAME

    def get_plugin_version(self) -> str:
        return SOFTMAX_TOPK_VERSION

    def enqueue(self, inputs):
        logits, mask = inputs
        return ops.softmax_topk(logits.contiguous(), None if mask is None else mask.reshape(-1).contiguous())


# ---- the other encoder plugins: same names, versions and creator fields as the reference's ----------------------
class _SimplePlugin:
    """One-output plugin over a C-ABI entry point; `fields` as the reference's creator parses them."""
    name = ""
    field_names = ()

    def __init__(self, fields: Dict[str, float]):
        dt = int(fields.get("data_type", 0))
        if dt not in _TORCH_DTYPE:
            raise ValueError("invalid data_type")
        self.fields = dict(fields)
        self.fields["data_type"] = dt

    def get_plugin_type(self) -> str:
        return self.name

    def get_plugin_version(self) -> str:
        return "1"

    def _check(self, t: torch.Tensor) -> torch.Tensor:
        if t.dtype != _TORCH_DTYPE[self.fields["data_type"]]:
            raise TypeError(f"plugin data_type is {_TORCH_DTYPE[self.fields['data_type']]}, got {t.dtype}")
        return t.contiguous()


class AttMaskedSoftmaxPlugin(_SimplePlugin):
    """AttMaskedSoftmaxPluginDynamic v1 (att_masked_softmax_plugin.cpp:82-108,157-178): fields data_type, scale; inputs
    scores [B, N, S, ld] (the reference reads them as [batch, seq_len, dim] = [B, N * S, ld]) + mask [B] int32."""
    name = "AttMaskedSoftmaxPluginDynamic"
    field_names = ("data_type", "scale")

    def enqueue(self, inputs):
        x, mask = inputs
        x = self._check(x)
        x4 = x if x.dim() == 4 else x.view(x.shape[0], 1, x.shape[1], x.shape[2])
        out = ops.att_masked_softmax(x4, None if mask is None else mask.reshape(-1).contiguous(),
                                     float(self.fields.get("scale", 1.0)))
        return out.view(x.shape)


class GluPlugin(_SimplePlugin):
    """GluPluginDynamic v1 (glu_plugin.cpp:187-200): fields data_type, axis_dim; the input is split in two along axis_dim."""
    name = "GluPluginDynamic"
    field_names = ("data_type", "axis_dim")

    def enqueue(self, inputs):
        x = self._check(inputs[0])
        ax = int(self.fields.get("axis_dim", 1)) % x.dim()
        M = 1
        for d in x.shape[:ax]:
            M *= d
        N = 1
        for d in x.shape[ax + 1:]:
            N *= d
        out = ops.glu(x.view(M, x.shape[ax], N))
        return out.view(*x.shape[:ax], x.shape[ax] // 2, *x.shape[ax + 1:])


class MaskedFillPlugin(_SimplePlugin):
    """MaskedFillPluginDynamic v1 (masked_fill_plugin.cpp:180-190): fields data_type, fill; inputs x [B, dim, T] (4-d
    [B, dim, 1, T] in the convolution module) + valid lengths [B] int32."""
    name = "MaskedFillPluginDynamic"
    field_names = ("data_type", "fill")

    def enqueue(self, inputs):
        x, mask = inputs
        x = self._check(x)
        out = ops.masked_fill(x.view(x.shape[0], -1, x.shape[-1]), mask.reshape(-1).contiguous(),
                              float(self.fields.get("fill", 0.0)))
        return out.view(x.shape)


class RelPositionalEncodingPlugin(_SimplePlugin):
    """RelPositionalEncodingPluginDynamic v1 (rel_positional_encoding_plugin.cpp:282-306): fields data_type, scale,
    max_len, dim, streaming (streaming = 0 here); the plugin owns the sinusoid table `pe` [max_len, dim]
    (positional_encoding.py:39-48); enqueue([x [B, T, dim]]) -> (x * scale, pe[:T])."""
    name = "RelPositionalEncodingPluginDynamic"
    field_names = ("data_type", "scale", "max_len", "dim", "streaming")

    def __init__(self, fields):
        super().__init__(fields)
        if int(self.fields.get("streaming", 0)) != 0:
            raise NotImplementedError("streaming positional encoding (frame offsets) is outside the offline path")
        import math
        d, n = int(self.fields["dim"]), int(self.fields.get("max_len", 5000))
        pe = torch.zeros(n, d)
        pos = torch.arange(0, n, dtype=torch.float32).unsqueeze(1)
        div = torch.exp(torch.arange(0, d, 2, dtype=torch.float32) * -(math.log(10000.0) / d))
        pe[:, 0::2] = torch.sin(pos * div)
        pe[:, 1::2] = torch.cos(pos * div)
        self._pe_cpu, self._pe = pe, None

    def enqueue(self, inputs):
        x = self._check(inputs[0])
        if self._pe is None or self._pe.device != x.device:
            self._pe = self._pe_cpu.to(device=x.device, dtype=x.dtype)
        return ops.rel_pos_encoding(x, self._pe, float(self.fields.get("scale", 1.0)))


class LayerNormPlugin(_SimplePlugin):
    """LayerNormPluginDynamic v1 (layer_norm_plugin.cpp:171-183): fields data_type, eps, dim; inputs x, gamma, beta."""
    name = "LayerNormPluginDynamic"
    field_names = ("data_type", "eps", "dim")

    def enqueue(self, inputs):
        x, gamma, beta = inputs
        return ops.layernorm(self._check(x), gamma.float().contiguous(), beta.float().contiguous(),
                             float(self.fields.get("eps", 1e-12)))


_SIMPLE = {c.name: c for c in (AttMaskedSoftmaxPlugin, GluPlugin, MaskedFillPlugin, RelPositionalEncodingPlugin,
                               LayerNormPlugin)}


class PluginRegistry:
    """get_plugin_creator(name, version, namespace) as used at positionwise_feed_forward.py:182,233, convolution.py:90,
    positional_encoding.py:104, network_helper (AttMaskedSoftmax / Glu / LayerNorm)."""

    def get_plugin_creator(self, name: str, version: str = "1", namespace: str = ""):
        if (name, version, namespace) == (FMOE_EXPERT_NAME, FMOE_EXPERT_VERSION, ""):
            return FMoEExpertPluginCreator()
        if (name, version, namespace) == (SOFTMAX_TOPK_NAME, SOFTMAX_TOPK_VERSION, ""):
            return type("SoftmaxTopKCreator", (), {
                "create_plugin": staticmethod(lambda n, fields: SoftmaxTopKPlugin(int(fields.get("data_type", 0))))})()
        if version == "1" and namespace == "" and name in _SIMPLE:
            cls = _SIMPLE[name]
            return type(name + "Creator", (), {"name": name, "field_names": cls.field_names,
                                               "create_plugin": staticmethod(lambda n, fields, c=cls: c(fields))})()
        return None
