"""The C restatement (oracle/moe_oracle.c) against the PyTorch oracle, which is itself pinned on the reference's vectors."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden, rel_l2

LIB = os.path.join(ROOT, "oracle", "libmoe_oracle.so")


@pytest.fixture(scope="module")
def clib():
    if not os.path.exists(LIB):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True)
    return C.CDLL(LIB)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _np(t):
    return None if t is None else np.ascontiguousarray(t.numpy().astype(np.float32))


def test_c_layer_matches_python_oracle_and_golden(clib, oracle):
    g = load_golden("case_3m_top1.npz")
    S, D = g["x"].shape
    Demb, E, H = g["embed"].shape[1], g["W1"].shape[0], g["W1"].shape[1]
    x, emb, Wr = _np(g["x"]), _np(g["embed"]), _np(g["Wr"])
    W1, b1, W2, b2 = _np(g["W1"]), _np(g["b1"]), _np(g["W2"]), _np(g["b2"])
    idx = np.zeros(S, np.int32); val = np.zeros(S, np.float32); counts = np.zeros(E, np.int32)
    mapping = np.zeros(S, np.int32); out = np.zeros((S, D), np.float32)
    clib.oracle_moe_forward_3m(_p(x), _p(emb), _p(Wr), None, _p(W1), _p(b1), _p(W2), _p(b2), _p(x), C.c_float(0.5),
                               S, D, Demb, E, H, 0, _p(idx), _p(val), _p(counts), _p(mapping), _p(out))
    assert np.array_equal(idx, g["gate_idx"].numpy())
    assert np.array_equal(counts, g["expert_count"].numpy())
    np.testing.assert_allclose(val, g["gate_value"].numpy(), rtol=1e-5)
    assert rel_l2(torch.from_numpy(out), g["final"]) < 1e-6
    r = oracle.moe_forward(g["x"], g["embed"], g["Wr"], None, g["W1"], g["b1"], g["W2"], g["b2"], residual=g["x"],
                           ff_scale=0.5)
    assert np.array_equal(mapping, r["mapping"].view(-1).numpy())


def test_c_naive_gate_and_prepare(clib, oracle, synth):
    E, D, S, k = 12, 48, 200, 3
    w = synth.make_weights(31, E, D, 64, 0, router_bias=True)
    x, _ = synth.make_activations(32, S, D, 0, w, top_k=k)
    idx = np.zeros((S, k), np.int32); score = np.zeros((S, k), np.float32)
    clib.oracle_gate_naive(_p(_np(x)), _p(_np(w.Wr)), _p(_np(w.br)), S, D, E, k, _p(idx), _p(score))
    ref_idx, ref_score, _ = oracle.gate_naive(x, w.Wr, w.br, k)
    assert np.array_equal(idx, ref_idx.numpy())
    np.testing.assert_allclose(score, ref_score.numpy(), rtol=1e-5, atol=1e-7)
    flat = idx.reshape(-1).copy()
    flat[::7] = -1
    counts = np.zeros(E, np.int32); offsets = np.zeros(E + 1, np.int32)
    mapping = np.zeros(S * k, np.int32); pos = np.zeros(S * k, np.int32)
    clib.oracle_prepare.restype = C.c_int
    nv = clib.oracle_prepare(_p(flat), S * k, E, _p(counts), _p(offsets), _p(mapping), _p(pos))
    p = oracle.prepare(torch.from_numpy(flat), E)
    assert nv == p["pos"].numel()
    assert np.array_equal(counts, p["counts"].numpy()) and np.array_equal(offsets, p["offsets"].numpy())
    assert np.array_equal(mapping, p["mapping"].numpy()) and np.array_equal(pos[:nv], p["pos"].numpy())


def test_c_block_matches_golden(clib):
    g = load_golden("case_block_3m.npz")
    S, D = g["x"].shape
    Demb, E, H = g["embed"].shape[1], g["W1"].shape[0], g["W1"].shape[1]
    a = {k: _np(g[k]) for k in ("x", "embed", "Wr", "W1", "b1", "W2", "b2", "ff_gamma", "ff_beta", "final_gamma",
                                 "final_beta")}
    idx = np.zeros(S, np.int32); val = np.zeros(S, np.float32); counts = np.zeros(E, np.int32)
    mapping = np.zeros(S, np.int32); out = np.zeros((S, D), np.float32)
    clib.oracle_moe_block_forward_3m(_p(a["x"]), _p(a["embed"]), _p(a["Wr"]), None, _p(a["W1"]), _p(a["b1"]),
                                     _p(a["W2"]), _p(a["b2"]), _p(a["ff_gamma"]), _p(a["ff_beta"]),
                                     _p(a["final_gamma"]), _p(a["final_beta"]), C.c_float(float(g["eps"])),
                                     C.c_float(float(g["ff_scale"])), S, D, Demb, E, H, 0, _p(idx), _p(val),
                                     _p(counts), _p(mapping), _p(out))
    assert np.array_equal(idx, g["gate_idx"].numpy())
    assert np.array_equal(counts, g["expert_count"].numpy())
    assert rel_l2(torch.from_numpy(out), g["out"]) < 1e-6
    xn = np.zeros((S, D), np.float32)
    clib.oracle_layer_norm(_p(a["x"]), _p(a["ff_gamma"]), _p(a["ff_beta"]), C.c_float(float(g["eps"])), S, D, _p(xn))
    np.testing.assert_allclose(xn, g["xn"].numpy(), rtol=1e-5, atol=1e-6)
