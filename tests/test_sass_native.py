"""The built library's contraction kernels are Blackwell-native code: tcgen05 MMAs (UTCHMMA, `.2CTA` for the pair
variants), TMA tensor loads (UTMALDG), TMEM reads (LDTM) and mbarrier traffic in the SASS of the kernels on the hot path,
and no warp-level mma.sync (HMMA) anywhere.  Reads the disassembly only -- no GPU needed (cuobjdump ships with the
toolkit on the build container and on the GPU box)."""
import os
import re
import shutil
import subprocess

import pytest

from conftest import ROOT

LIB = os.path.join(ROOT, "3m-asr-inference_b200", "libb200moe.so")


@pytest.fixture(scope="module")
def opcodes():
    if shutil.which("cuobjdump") is None or shutil.which("c++filt") is None:
        pytest.skip("cuobjdump / c++filt not on PATH")
    if not os.path.exists(LIB):
        pytest.fail("libb200moe.so is not built (python -c 'import __graft_entry__ as g; g.build()')")
    out = subprocess.run(["python", os.path.join(ROOT, "tools", "sass_opcodes.py"), LIB], capture_output=True, text=True,
                         check=True).stdout
    per, name = {}, None
    for line in out.splitlines():
        if line and not line.startswith(" ") and ("kernel" in line or line == "whole library"):
            name = line.strip()
        elif line.startswith("    ") and name:
            per[name] = {k: int(v) for k, v in re.findall(r"([A-Z][A-Z0-9_.]*) x(\d+)", line)}
            name = None
    return per


def _kernel(per, needle):
    hits = [v for k, v in per.items() if needle in k]
    assert hits, f"no kernel matching {needle!r} in the SASS summary: {sorted(per)[:8]} ..."
    return hits[0]


def test_expert_kernel_is_tcgen05_tma(opcodes):
    k = _kernel(opcodes, "ffn_kernel<__nv_bfloat16, 0, 1, false>")   # product variant, single-CTA tiles
    assert k.get("UTCHMMA", 0) >= 8 and k.get("UTMALDG.3D", 0) >= 8 and k.get("LDTM", 0) >= 1
    assert k.get("UTCBAR", 0) >= 1                                       # tcgen05.commit -> mbarrier
    pair = _kernel(opcodes, "ffn_kernel<__nv_bfloat16, 0, 2, false>")   # cta_group::2 pairs
    assert pair.get("UTCHMMA.2CTA", 0) >= 8 and pair.get("UTMALDG.3D.2CTA", 0) >= 8


def test_router_kernels_are_tcgen05_tma(opcodes):
    for needle in ("route_kernel<false, false>", "route_kernel<true, false>", "gate_tc_kernel"):
        k = _kernel(opcodes, needle)
        assert k.get("UTCHMMA", 0) >= 4 and k.get("LDTM", 0) >= 1, (needle, k)
        assert any(op.startswith("UTMALDG") for op in k), (needle, k)


def test_no_warp_level_mma_anywhere(opcodes):
    whole = opcodes["whole library"]
    assert not any(op.startswith("HMMA") for op in whole), whole
    assert whole.get("UTCHMMA", 0) + whole.get("UTCHMMA.2CTA", 0) >= 100
