"""Shared test plumbing.  `-m "not gpu"`: oracle vs golden vectors, host logic, C-ABI symbol checks (no compute).
`-m gpu`: the parity tests proper -- the CUDA path through the C ABI against the oracle on identical seeded inputs."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG_NAME = "3m-asr-inference_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (sm_100a); run on the B200 box")


def pkg(sub: str = ""):
    """The package directory name is not a Python identifier, so it is imported by string."""
    return importlib.import_module(PKG_NAME + (("." + sub) if sub else ""))


@pytest.fixture(scope="session")
def oracle():
    return importlib.import_module("oracle.moe_oracle")


@pytest.fixture(scope="session")
def synth():
    return pkg("synth")


@pytest.fixture(scope="session")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return pkg("ops")


def from_bits(a: np.ndarray) -> torch.Tensor:
    """uint16 bf16 bit patterns -> fp32 tensor (inverse of tests/golden/make_golden.py:bits)."""
    return torch.from_numpy((a.astype(np.uint32) << 16).view(np.float32).copy())


def load_golden(name: str):
    z = np.load(os.path.join(ROOT, "tests", "golden", name))
    out = {}
    for k in z.files:
        if k.endswith("_bf16"):
            out[k[:-5]] = from_bits(z[k])
        else:
            v = z[k]
            out[k] = torch.from_numpy(v.copy()) if v.ndim else v.item()
    return out


def load_golden_repo_dims(synth):
    """case_3m_repo_dims.npz (E=32, D=512, H=1024): activations, router and the reference's outputs come from the file,
    the 128 MiB of expert weights are re-generated from the recorded seed and checked against the recorded checksums."""
    g = load_golden("case_3m_repo_dims.npz")
    w = synth.make_weights(int(g["weight_seed"]), 32, 512, 1024, 512, random_bias=True)
    assert torch.equal(w.Wr, g["Wr"]), "seeded weight stream differs from the one the fixture was made with"
    for name in ("W1", "W2", "b1", "b2"):
        assert abs(float(getattr(w, name).double().sum()) - float(g[name + "_sum"])) < 1e-9, name
    probe = torch.arange(0, 32 * 1024 * 512, 1048573)[:16]
    assert torch.equal(w.W1.reshape(-1)[probe], g["W1_probe"]) and torch.equal(w.W2.reshape(-1)[probe], g["W2_probe"])
    return g, w


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.double().cpu()
    b = b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
