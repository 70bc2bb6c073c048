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


def test_plugin_registry_serves_the_encoder_plugins(ops):
    """The reference's plugin names / creator fields (TRTAPI++/plugin/*/..._plugin.cpp) through the registry mirror."""
    plugin = __import__("conftest").pkg("plugin")
    reg = plugin.PluginRegistry()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 6, 40, generator=g).cuda()
    lens = torch.tensor([40, 13], dtype=torch.int32).cuda()
    mf = reg.get_plugin_creator("MaskedFillPluginDynamic", "1", "").create_plugin("m", {"data_type": 0, "fill": 0.0})
    out = mf.enqueue([x.view(2, 6, 1, 40), lens])
    assert out.shape == (2, 6, 1, 40) and float(out[1, :, :, 13:].abs().max()) == 0.0 and torch.equal(out[0], x.view(2, 6, 1, 40)[0])
    glu = reg.get_plugin_creator("GluPluginDynamic", "1", "").create_plugin("g", {"data_type": 0, "axis_dim": 1})
    torch.testing.assert_close(glu.enqueue([x.view(2, 6, 1, 40)]), torch.nn.functional.glu(x.view(2, 6, 1, 40), 1),
                               rtol=2e-6, atol=1e-7)
    sm = reg.get_plugin_creator("AttMaskedSoftmaxPluginDynamic", "1", "").create_plugin("s", {"data_type": 0, "scale": 0.125})
    p = sm.enqueue([x, lens])          # [batch, seq_len, dim] as the reference's enqueue reads it
    assert p.shape == x.shape and float(p[1, :, 13:].abs().max()) == 0.0
    torch.testing.assert_close(p[1, :, :13].sum(-1), torch.ones(6, device="cuda"), rtol=1e-5, atol=1e-6)
    rp = reg.get_plugin_creator("RelPositionalEncodingPluginDynamic", "1", "").create_plugin(
        "r", {"data_type": 0, "scale": 2.0, "max_len": 64, "dim": 40, "streaming": 0})
    y, pos = rp.enqueue([x])
    torch.testing.assert_close(y, x * 2.0)
    assert pos.shape == (6, 40) and float(pos[0, 1]) == 1.0 and float(pos[0, 0]) == 0.0
    ln = reg.get_plugin_creator("LayerNormPluginDynamic", "1", "").create_plugin("l", {"data_type": 0, "eps": 1e-12, "dim": 40})
    gam, bet = torch.ones(40).cuda(), torch.zeros(40).cuda()
    torch.testing.assert_close(ln.enqueue([x, gam, bet]), torch.nn.functional.layer_norm(x, (40,), gam, bet, 1e-12),
                               rtol=1e-4, atol=1e-5)
    assert reg.get_plugin_creator("GluPluginDynamic", "2", "") is None
