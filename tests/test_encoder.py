"""the Conformer-MoE encoder wrapper around the hot path.
CPU: the module tree / state_dict keys are the reference's, lengths, padding invariance (fast_moe layer = CPU oracle).
GPU: the encoder with the fast_moe blocks running through the C ABI against the same graph on the CPU in fp32 with the
oracle in the MoE slot."""
import os
import re

import pytest
import torch

from conftest import pkg, rel_l2

SMALL = dict(input_dim=40, output_dim=96, attention_heads=4, attention_dim=128, num_blocks=2,
             embed_conf=dict(attention_heads=4, attention_dim=128, linear_units=256, num_blocks=1),
             moe_conf=dict(num_experts=8, hidden_units=256, rand_init_router=True))


def make_encoder(seed=0, **over):
    enc = pkg("encoder")
    torch.manual_seed(seed)
    cfg = dict(SMALL)
    cfg.update(over)
    m = enc.ConformerMoEEncoder(**cfg)
    with torch.no_grad():  # non-trivial norms, biases and batch-norm statistics; expert biases off zero
        for name, p in m.named_parameters():
            if name.endswith("norm.weight") or ".norm_" in name and name.endswith("weight"):
                p.add_(0.1 * torch.randn_like(p))
            if name.endswith("bias"):
                p.add_(0.05 * torch.randn_like(p))
        for b in m.modules():
            if isinstance(b, torch.nn.BatchNorm1d):
                b.running_mean.normal_(0, 0.1)
                b.running_var.uniform_(0.5, 1.5)
    return m.eval()


def oracle_block(oracle):
    """ffn_impl for the CPU reference: fmoe_transformer.py:144-166 through the oracle."""
    def impl(block, x, embed, x_len):
        ff = block.feed_forward
        B, T, D = x.shape
        r = oracle.moe_block_forward(
            x.reshape(B * T, D).float(), embed.reshape(B * T, -1).float(), ff.router_weights.detach().float(),
            None if ff.router_bias is None else ff.router_bias.detach().float(),
            ff.experts.w_1.weight.detach().float(), ff.experts.w_1.bias.detach().float(),
            ff.experts.w_2.weight.detach().float(), ff.experts.w_2.bias.detach().float(),
            norm_ff=(block.norm_ff.weight.detach().float(), block.norm_ff.bias.detach().float()),
            norm_final=(block.norm_final.weight.detach().float(), block.norm_final.bias.detach().float()),
            eps=block.norm_ff.eps, ff_scale=block.ff_scale, x_len=x_len, T=T)
        return r["out"].view(B, T, D).to(x.dtype)
    return impl


def test_block_state_dict_keys_are_the_reference_ones():
    enc = pkg("encoder")
    m = make_encoder()
    keys = set(m.state_dict())
    block0 = {k[len("blocks.0."):] for k in keys if k.startswith("blocks.0.")}
    assert block0 == set(enc.REFERENCE_BLOCK_KEYS)
    for top in ("subsampling.conv.0.weight", "subsampling.conv.2.bias", "subsampling.out.0.weight", "after_norm.weight",
                "after_norm_6.bias", "after_norm_12.weight", "out_linear.weight", "embed.subsampling.out.0.bias",
                "embed.after_norm.weight", "embed.out_linear.bias", "embed.blocks.0.feed_forward.w_1.weight",
                "embed.blocks.0.self_attn.linear_pos.weight", "embed.blocks.0.conv_module.depthwise_conv.weight"):
        assert top in keys, top
    assert not any("pe" == k.split(".")[-1] for k in keys)      # the positional table is not part of the checkpoint
    sd = m.state_dict()
    assert tuple(sd["blocks.0.feed_forward.experts.w_1.weight"].shape) == (8, 256, 128)
    assert tuple(sd["blocks.0.feed_forward.router_weights"].shape) == (256, 8)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_module_members_exist_in_reference_sources():
    """Every nn member the reference constructors create has a same-named member here (sources parsed, not imported:
    they need TensorRT)."""
    enc = pkg("encoder")
    m = make_encoder()
    root = "/root/reference/trainer_3m_fix/"
    pairs = [("layer/fmoe_transformer.py", "FmoeConformerLayer", m.blocks[0]),
             ("layer/attention.py", "RelPositionMultiHeadedAttention", m.blocks[0].self_attn),
             ("layer/attention.py", "MultiHeadedAttention", m.blocks[0].self_attn),
             ("layer/convolution.py", "ConvolutionModule", m.blocks[0].conv_module),
             ("layer/subsampling.py", "Conv2dSubsampling4", m.subsampling),
             ("model/conformer_fmoe_localComm_catEmbed_domain_acc_hier.py", "Net", m),
             ("model/conformer_embed_domain_acc.py", "Net", m.embed)]
    for path, cls, obj in pairs:
        src = open(root + path).read()
        body = src[src.index("class " + cls):]
        nxt = re.search(r"\nclass \w+", body[1:])
        body = body[: nxt.start() + 1] if nxt else body
        members = set(re.findall(r"^\s+self\.(\w+) = (?:torch\.)?nn\.(?!Dropout|Softmax)\w+", body, re.M))
        assert members, (path, cls)
        for name in members:
            assert hasattr(obj, name), (cls, name)


def test_subsampled_lengths_follow_the_mask_sample_plugin():
    """(in - left_padding - 1) / stride + 1 with left_padding = stride = 2, twice (mask_conv2d_sample_kernel.cu:27-35); for
    an unpadded utterance that is also what slicing the mask `[:, :, :-2:2]` twice gives and what the convolutions produce."""
    enc = pkg("encoder")
    lens = torch.arange(7, 1200)
    got = enc.subsampled_lengths(lens)
    for L, g in zip(lens.tolist(), got.tolist()):
        l1 = (L - 2 - 1) // 2 + 1
        assert g == (l1 - 2 - 1) // 2 + 1
        assert g == torch.ones(L)[:-2:2][:-2:2].numel()
    assert int(enc.subsampled_lengths(torch.tensor([206]))[0]) == 50
    sub = enc.Conv2dSubsampling4(40, 16)
    for L in (7, 63, 206):
        assert sub(torch.zeros(1, L, 40), None)[0].shape[1] == int(enc.subsampled_lengths(torch.tensor([L]))[0])


def test_padding_does_not_change_valid_frames(oracle):
    """An utterance alone and the same utterance padded inside a batch give (nearly) the same valid output frames."""
    m = make_encoder(3)
    impl = oracle_block(oracle)
    g = torch.Generator().manual_seed(1)
    a = torch.randn(1, 63, 40, generator=g)
    b = torch.randn(1, 103, 40, generator=g)
    batch = torch.zeros(2, 103, 40)
    batch[0, :63] = a
    batch[1] = b
    lens = torch.tensor([63, 103])
    with torch.no_grad():
        out = m(batch, lens, ffn_impl=impl)
        alone = m(a, torch.tensor([63]), ffn_impl=impl)
    n = int(pkg("encoder").subsampled_lengths(torch.tensor([63]))[0])
    assert alone.shape[1] == n and out.shape[1] == 25
    # Not exact, in the reference either: the convolution module zeroes padded frames BEFORE pointwise_conv1, whose bias
    # then re-fills them, so the depthwise convolution of the last 7 valid frames sees bias-driven values where the
    # utterance alone sees zero padding (convolution.py:104-131); attention spreads that over the utterance.
    assert rel_l2(out[0, :n], alone[0]) < 1e-2


def test_product_forward_has_no_cpu_path():
    m = make_encoder()
    with pytest.raises(RuntimeError):
        m(torch.randn(1, 63, 40), torch.tensor([63]))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-2), (torch.bfloat16, 6e-2)])
def test_encoder_gpu_matches_cpu_reference(ops, oracle, dtype, tol):
    """Two fast_moe blocks + the embed net on the GPU (C ABI in the MoE slot, library ops elsewhere) against the fp32 CPU
    graph with the oracle in the MoE slot.  fp32 activations isolate the MoE path's own bf16 arithmetic; bf16 activations
    are the deployed flavour (every library op rounds as well, hence the wider bar)."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    m = make_encoder(5)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(3, 120, 40, generator=g)
    lens = torch.tensor([120, 77, 101])
    with torch.no_grad():
        ref = m(x, lens, ffn_impl=oracle_block(oracle))
        import copy
        mg = copy.deepcopy(m).cuda()
        if dtype == torch.bfloat16:
            mg.to_inference(dtype)
        out = mg(x.cuda().to(dtype), lens.cuda()).float().cpu()
    sub = pkg("encoder").subsampled_lengths(lens)
    for b in range(3):
        n = int(sub[b])
        assert rel_l2(out[b, :n], ref[b, :n]) < tol, (b, rel_l2(out[b, :n], ref[b, :n]))
