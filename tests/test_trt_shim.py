"""The TensorRT plugin shim (3m-asr-inference_b200/csrc/trt/b200moe_trt_plugins.cpp) driven
through a stand-in for the TensorRT 8 plugin API (tests/trt_stub/NvInfer.h): registry lookup by the reference's plugin
names, creators with the reference's field names, serialisation byte-for-byte as the reference writes it, and -- on the
GPU -- enqueue with the reference's input order against the oracle.  TensorRT itself is not in the image."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden, rel_l2

STUB = os.path.join(ROOT, "tests", "trt_stub", "libb200moe_trt_stub.so")
K_FLOAT, K_HALF, K_INT32 = 0, 1, 3   # nvinfer1::DataType


@pytest.fixture(scope="module")
def shim():
    if not os.path.exists(STUB):
        subprocess.run(["make", "-C", os.path.join(ROOT, "3m-asr-inference_b200", "csrc"), "trt-stub"], check=True)
    lib = C.CDLL(STUB)
    lib.shim_create.restype = C.c_void_p
    lib.shim_create.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int), C.c_char_p, C.c_float]
    lib.shim_deserialize.restype = C.c_void_p
    lib.shim_deserialize.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t]
    lib.shim_clone.restype = C.c_void_p
    lib.shim_clone.argtypes = [C.c_void_p]
    lib.shim_destroy.argtypes = [C.c_void_p]
    for f in ("shim_type", "shim_version", "shim_get_namespace"):
        getattr(lib, f).restype = C.c_char_p
        getattr(lib, f).argtypes = [C.c_void_p]
    lib.shim_set_namespace.argtypes = [C.c_void_p, C.c_char_p]
    lib.shim_nb_outputs.argtypes = [C.c_void_p]
    lib.shim_serialization_size.restype = C.c_size_t
    lib.shim_serialization_size.argtypes = [C.c_void_p]
    lib.shim_serialize.argtypes = [C.c_void_p, C.c_void_p]
    lib.shim_supports.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int]
    lib.shim_output_dims.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(C.c_int)]
    lib.shim_workspace.restype = C.c_size_t
    lib.shim_workspace.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int]
    lib.shim_enqueue.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_int,
                                 C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p]
    lib.shim_field_names.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    return lib


def create(shim, plugin, fields, ffield=None):
    names = (C.c_char_p * len(fields))(*[k.encode() for k in fields])
    vals = (C.c_int * len(fields))(*fields.values())
    fname, fval = (ffield[0].encode(), ffield[1]) if ffield else (None, 0.0)
    return shim.shim_create(plugin.encode(), len(fields), names, vals, fname, fval)


def serialise(shim, p):
    n = shim.shim_serialization_size(p)
    buf = C.create_string_buffer(n)
    shim.shim_serialize(p, buf)
    return buf.raw


def ints(*v):
    return (C.c_int * len(v))(*v)


FMOE_FIELDS = {"data_type": 0, "num_expert": 32, "idim": 512, "hidden_units": 1024, "act_type": 0}


def test_registry_and_creator_fields(shim):
    assert shim.shim_num_creators() == 3
    buf = C.create_string_buffer(256)
    # the field names builder.py passes (positionwise_feed_forward.py:233-247; network_helper.py:87,445-470)
    assert shim.shim_field_names(b"FMoEExpertPluginDynamic", buf, 256) == 5
    assert buf.value.decode().split(",") == ["data_type", "num_expert", "idim", "hidden_units", "act_type"]
    assert shim.shim_field_names(b"SoftmaxTopKPluginDynamic", buf, 256) == 1 and buf.value == b"data_type"
    assert shim.shim_field_names(b"LayerNormPluginDynamic", buf, 256) == 3
    assert buf.value.decode().split(",") == ["data_type", "dim", "eps"]
    assert shim.shim_field_names(b"NoSuchPlugin", buf, 256) == -1


def test_fmoe_plugin_host_surface(shim):
    p = create(shim, "FMoEExpertPluginDynamic", FMOE_FIELDS)
    assert p
    assert shim.shim_type(p) == b"FMoEExpertPluginDynamic" and shim.shim_version(p) == b"1"
    assert shim.shim_nb_outputs(p) == 1
    # 32 bytes: the five fields + three zero ints (fmoe_expert_plugin.cpp:288-304)
    raw = serialise(shim, p)
    assert raw == struct.pack("<8i", 0, 32, 512, 1024, 0, 0, 0, 0)
    q = shim.shim_deserialize(b"FMoEExpertPluginDynamic", raw, len(raw))
    assert q and serialise(shim, q) == raw
    shim.shim_set_namespace(q, b"ns")
    c = shim.shim_clone(q)
    assert serialise(shim, c) == raw and shim.shim_get_namespace(c) == b"ns"
    # input, gate_idx (int32), four weight tensors, output: all `data_type` except gate_idx, linear format
    types = ints(K_FLOAT, K_INT32, K_FLOAT, K_FLOAT, K_FLOAT, K_FLOAT, K_FLOAT)
    assert all(shim.shim_supports(p, pos, types, 6, 1) == 1 for pos in range(7))
    assert shim.shim_supports(p, 1, ints(K_FLOAT, K_FLOAT, K_FLOAT, K_FLOAT, K_FLOAT, K_FLOAT, K_FLOAT), 6, 1) == 0
    assert shim.shim_supports(p, 0, ints(K_HALF, K_INT32, K_FLOAT, K_FLOAT, K_FLOAT, K_FLOAT, K_FLOAT), 6, 1) == 0
    assert shim.shim_supports(p, 0, types, 5, 1) == 0
    out = ints(0, 0, 0)
    assert shim.shim_output_dims(p, 0, ints(4, 50, 512), 3, 6, out) == 3 and list(out) == [4, 50, 512]
    assert shim.shim_workspace(p, ints(4, 50, 512), 3, 6, 1) > 4 * 50 * (512 + 1024) * 2
    assert not create(shim, "FMoEExpertPluginDynamic", dict(FMOE_FIELDS, data_type=7))   # the reference: invalid type_id
    assert not create(shim, "FMoEExpertPluginDynamic", dict(FMOE_FIELDS, idim=500))      # not a multiple of 128
    for h in (p, q, c):
        shim.shim_destroy(h)


def test_softmax_topk_and_layernorm_host_surface(shim):
    p = create(shim, "SoftmaxTopKPluginDynamic", {"data_type": 1})
    assert shim.shim_type(p) == b"SoftmaxTopKPluginDynamic" and shim.shim_nb_outputs(p) == 2
    raw = serialise(shim, p)
    assert raw == struct.pack("<6i", 1, -1, 1, 0, 0, 0)                   # softmax_topk_plugin.cpp: type, axis_dim, k, 3 x 0
    q = shim.shim_deserialize(b"SoftmaxTopKPluginDynamic", raw, len(raw))
    assert serialise(shim, q) == raw
    types = ints(K_HALF, K_INT32, K_HALF, K_INT32)                        # logits, mask, value, idx
    assert all(shim.shim_supports(p, pos, types, 2, 2) == 1 for pos in range(4))
    assert shim.shim_supports(p, 2, ints(K_HALF, K_INT32, K_FLOAT, K_INT32), 2, 2) == 0
    out = ints(0, 0, 0)
    assert shim.shim_output_dims(p, 1, ints(4, 50, 32), 3, 2, out) == 3 and list(out) == [4, 50, 1]
    ln = create(shim, "LayerNormPluginDynamic", {"data_type": 0, "dim": 512}, ("eps", 1e-12))
    assert shim.shim_type(ln) == b"LayerNormPluginDynamic" and shim.shim_nb_outputs(ln) == 1
    raw = serialise(shim, ln)
    assert raw == struct.pack("<iQf", 0, 512, 1e-12)                      # int32 data_type, size_t dim, float eps
    ln2 = shim.shim_deserialize(b"LayerNormPluginDynamic", raw, len(raw))
    assert serialise(shim, ln2) == raw
    assert all(shim.shim_supports(ln, pos, ints(K_FLOAT, K_FLOAT, K_FLOAT, K_FLOAT), 3, 1) == 1 for pos in range(4))
    for h in (p, q, ln, ln2):
        shim.shim_destroy(h)


def ptrs(*tensors):
    return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


@pytest.mark.gpu
def test_fmoe_plugin_enqueue_matches_oracle(shim, ops, oracle, synth):
    """infer.py's call: six inputs in the reference's order, fp32 (`data_type` 0), one un-weighted output."""
    E, D, H, B, T = 32, 512, 1024, 4, 50
    w = synth.make_weights(311, E, D, H, 512, random_bias=True)
    x, emb = synth.make_activations(312, B * T, D, 512, w)
    ref = oracle.moe_forward(x, emb, w.Wr, None, w.W1, w.b1, w.W2, w.b2, keep_expert_output=True)
    p = create(shim, "FMoEExpertPluginDynamic", FMOE_FIELDS)
    xd = x.cuda().view(B, T, D).contiguous()
    gate_idx = ref["idx"].view(B, T).to(torch.int32).cuda()
    W1, b1, W2, b2 = w.W1.cuda(), w.b1.cuda(), w.W2.cuda(), w.b2.cuda()
    out = torch.empty_like(xd)
    dims = ints(B, T, D)
    ws = torch.empty(shim.shim_workspace(p, dims, 3, 6, 1), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(2):   # the second call re-uses the packed weights
        assert shim.shim_enqueue(p, dims, 3, 6, ptrs(xd, gate_idx, W1, b1, W2, b2), 1, ptrs(out), ws.data_ptr(),
                                 stream) == 0
    torch.cuda.synchronize()
    assert rel_l2(out.cpu().view(B * T, D), ref["moe"]) <= 1e-2      # un-weighted expert output in token order
    shim.shim_destroy(p)


@pytest.mark.gpu
def test_softmax_topk_and_layernorm_enqueue(shim, ops, oracle, synth):
    B, T, E, D = 3, 40, 32, 512
    torch.manual_seed(5)
    logits = torch.randn(B, T, E)
    mask = torch.tensor([40, 13, 0], dtype=torch.int32)
    p = create(shim, "SoftmaxTopKPluginDynamic", {"data_type": 0})
    ld, md = logits.cuda(), mask.cuda()
    value = torch.empty(B, T, 1, device="cuda")
    idx = torch.empty(B, T, 1, dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    assert shim.shim_enqueue(p, ints(B, T, E), 3, 2, ptrs(ld, md), 2, ptrs(value, idx), None, stream) == 0
    torch.cuda.synchronize()
    prob = torch.softmax(logits, -1)
    valid = torch.arange(T)[None, :] < mask[:, None].long()
    assert torch.equal(idx.cpu().view(B, T)[valid].long(), prob.argmax(-1)[valid])
    torch.testing.assert_close(value.cpu().view(B, T)[valid], prob.max(-1).values[valid], rtol=1e-5, atol=1e-7)
    assert bool((idx.cpu().view(B, T)[~valid] == -1).all())
    shim.shim_destroy(p)

    x = torch.randn(B, T, D) * 2 + 0.5
    gamma, beta = 1 + 0.2 * torch.randn(D), 0.1 * torch.randn(D)
    ln = create(shim, "LayerNormPluginDynamic", {"data_type": 0, "dim": D}, ("eps", 1e-12))
    xd, gd, bd = x.cuda(), gamma.cuda(), beta.cuda()
    y = torch.empty_like(xd)
    assert shim.shim_enqueue(ln, ints(B, T, D), 3, 3, ptrs(xd, gd, bd), 1, ptrs(y), None, stream) == 0
    torch.cuda.synchronize()
    torch.testing.assert_close(y.cpu(), oracle.layer_norm(x, gamma, beta, 1e-12), rtol=2e-5, atol=2e-6)
    shim.shim_destroy(ln)
