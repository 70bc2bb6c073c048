"""The C-ABI library loads and exports every symbol include/b200moe.h declares; host-only entry points behave.
No compute call is made here (no GPU in the CPU test tier)."""
import ctypes
import os
import re
import struct

import pytest

from conftest import ROOT, pkg

HEADER = os.path.join(ROOT, "include", "b200moe.h")


@pytest.fixture(scope="module")
def lib():
    return pkg("_lib").load()


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200moe_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(lib):
    names = declared_symbols()
    assert len(names) >= 20
    sig = pkg("_lib").SIGNATURES
    for n in names:
        assert hasattr(lib, n), f"{n} declared in b200moe.h but not exported by libb200moe.so"
        assert n in sig, f"{n} declared in b200moe.h but has no ctypes prototype in _lib.py"
    assert sorted(sig) == names


def test_version_and_workspace(lib):
    assert lib.b200moe_version() == 100
    small = lib.b200moe_workspace_bytes(50, 32, 512, 1024, 1)
    big = lib.b200moe_workspace_bytes(3200, 32, 512, 1024, 1)
    assert 0 < small < big
    # must hold the bf16 scatter buffer and the bf16 hidden buffer at least
    assert big >= 3200 * (512 + 1024) * 2
    assert lib.b200moe_workspace_bytes(-1, 32, 512, 1024, 1) == 0


def test_plugin_fields_and_32_byte_serialisation(lib):
    # reference: data_type, num_expert, idim, hidden_units, act_type + 3 zero ints (fmoe_expert_plugin.cpp:288-304)
    h = lib.b200moe_plugin_create(0, 32, 512, 1024, 0)
    assert h
    assert lib.b200moe_plugin_serialization_size(h) == 32
    buf = ctypes.create_string_buffer(32)
    assert lib.b200moe_plugin_serialize(h, buf) == 0
    assert struct.unpack("8i", buf.raw) == (0, 32, 512, 1024, 0, 0, 0, 0)
    h2 = lib.b200moe_plugin_deserialize(buf.raw, 32)
    assert h2
    buf2 = ctypes.create_string_buffer(32)
    lib.b200moe_plugin_serialize(h2, buf2)
    assert buf2.raw == buf.raw
    h3 = lib.b200moe_plugin_clone(h)
    assert h3 and lib.b200moe_plugin_workspace_bytes(h3, 206) == lib.b200moe_plugin_workspace_bytes(h, 206)
    for x in (h, h2, h3):
        lib.b200moe_plugin_destroy(x)


def test_creator_rejects_bad_type_like_the_reference(lib):
    # fmoe_expert_plugin.cpp:360-363: type_id outside the supported set -> nullptr
    assert not lib.b200moe_plugin_create(5, 32, 512, 1024, 0)
    assert b"invalid type_id" in lib.b200moe_last_error()
    assert not lib.b200moe_plugin_deserialize(b"\0" * 8, 8)


def test_argument_errors_do_not_touch_the_gpu(lib):
    L = pkg("_lib")
    assert lib.b200moe_forward(None, None, 0, None) == -1
    a = L.LayerArgs(B=1, T=4, D=100, H=256, E=4, top_k=1)
    assert lib.b200moe_forward(ctypes.byref(a), None, 0, None) == -1
    assert b"multiples of 128" in lib.b200moe_last_error()
    a = L.LayerArgs(B=1, T=4, D=128, H=256, E=4, top_k=2, gate_mode=0)
    assert lib.b200moe_forward(ctypes.byref(a), None, 0, None) == -1
    assert b"top-1" in lib.b200moe_last_error()
    assert lib.b200moe_gate(None, None, None, None, None, 1, 4, 128, 0, 300, 1, 0, 2, None, None, None) == -1


def test_python_wrappers_refuse_cpu_tensors():
    import torch
    ops = pkg("ops")
    x = torch.zeros(4, 128)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.gate(x, None, torch.zeros(128, 4))
