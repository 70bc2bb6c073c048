"""the reference's other encoder plugins as sm_100a kernels, against plain torch on the same
inputs.  Tolerances: fp32 1e-6 relative (expf vs exp), 16-bit types one rounding of the output."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 2e-6, torch.float16: 2e-3, torch.bfloat16: 1.6e-2}


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("B,N,S,ld", [(1, 8, 50, 50), (3, 4, 25, 25), (2, 2, 7, 300), (64, 8, 50, 50), (2, 1, 3, 1024)])
def test_att_masked_softmax(ops, dtype, B, N, S, ld):
    g = torch.Generator().manual_seed(B * 1000 + ld)
    x = (torch.randn(B, N, S, ld, generator=g) * 3).to(dtype)
    mask = torch.randint(1, ld + 1, (B,), generator=g, dtype=torch.int32)
    mask[0] = ld
    if B > 1:
        mask[1] = 0                                                     # no valid key at all: zeros, not NaN
    scale = 1.0 / math.sqrt(64)
    out = ops.att_masked_softmax(x.cuda(), mask.cuda(), scale).float().cpu()
    keep = torch.arange(ld)[None, :] < mask[:, None]
    ref = torch.softmax((x.float() * scale).masked_fill(~keep[:, None, None, :], float("-inf")), -1)
    ref = torch.nan_to_num(ref, nan=0.0)
    assert torch.equal(out[~keep[:, None, None, :].expand_as(out)], torch.zeros(int((~keep).sum()) * N * S))
    torch.testing.assert_close(out, ref, rtol=TOL[dtype], atol=TOL[dtype] * 0.05)
    out2 = ops.att_masked_softmax(x.cuda(), None, scale).float().cpu()  # no mask input = all keys valid
    torch.testing.assert_close(out2, torch.softmax(x.float() * scale, -1), rtol=TOL[dtype], atol=TOL[dtype] * 0.05)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("M,C,N", [(1, 512, 50), (64, 512, 50), (3, 5, 7), (2, 1024, 1)])
def test_glu(ops, dtype, M, C, N):
    g = torch.Generator().manual_seed(M + C + N)
    x = torch.randn(M, 2 * C, N, generator=g).to(dtype)
    out = ops.glu(x.cuda()).float().cpu()
    ref = torch.nn.functional.glu(x.float(), dim=1)
    torch.testing.assert_close(out, ref, rtol=TOL[dtype], atol=TOL[dtype] * 0.1)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_masked_fill_is_bit_exact(ops, dtype):
    g = torch.Generator().manual_seed(9)
    B, dim, T = 5, 37, 53
    x = torch.randn(B, dim, T, generator=g).to(dtype)
    mask = torch.tensor([53, 0, 17, 52, 1], dtype=torch.int32)
    for fill in (0.0, -1.5):
        out = ops.masked_fill(x.cuda(), mask.cuda(), fill).cpu()
        keep = (torch.arange(T)[None, :] < mask[:, None])[:, None, :]
        ref = torch.where(keep, x, torch.tensor(fill, dtype=dtype))
        assert torch.equal(out, ref)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_rel_pos_encoding(ops, dtype):
    enc = __import__("conftest").pkg("encoder")
    pe = enc.RelPositionalEncoding(512).pe[0].to(dtype)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(3, 50, 512, generator=g).to(dtype)
    out, pos = ops.rel_pos_encoding(x.cuda(), pe.cuda(), math.sqrt(512))
    assert torch.equal(pos.cpu(), pe[:50])
    torch.testing.assert_close(out.float().cpu(), x.float() * math.sqrt(512), rtol=TOL[dtype], atol=0)
