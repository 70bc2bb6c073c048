"""GPU parity of the Conformer block's feed-forward part: norm_ff -> MoE layer -> residual +
ff_scale * y -> norm_final through b200moe_block_forward / b200moe_layernorm, against the oracle and against the vectors
made with the reference's own FmoeConformerLayer members (tests/golden/case_block_3m.npz).

Routing under a LayerNorm: the normalised input is rounded to the activation dtype before it meets the router, so a
token whose two best logits are closer than that rounding can legitimately go either way against an fp32 reference.
Two bars therefore: (1) given the SAME normalised input (b200moe_layernorm's own output handed to the oracle) expert
assignment, counts and scatter indices are bit-exact; (2) against the pure fp32 chain, rows whose routing agrees are
within the BF16 tolerance and the rows that disagree are exactly the near-ties."""
import pytest
import torch

from conftest import load_golden, pkg, rel_l2

pytestmark = pytest.mark.gpu

BF16_REL_L2 = 1e-2


@pytest.fixture(autouse=True, params=[1, 0], ids=["route", "split"])
def route_mode(request, ops):
    ops.config("route", request.param)
    yield request.param
    ops.config("route", 1)


def make_norm(D, seed):
    g = torch.Generator().manual_seed(seed)
    gamma = (1.0 + 0.2 * torch.randn(D, generator=g)).bfloat16().float()
    beta = (0.1 * torch.randn(D, generator=g)).bfloat16().float()
    return gamma, beta


def cu(pair):
    return None if pair is None else (pair[0].cuda(), pair[1].cuda())


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("D", [128, 264, 512, 1024])
def test_layernorm_kernel(ops, oracle, dtype, D):
    torch.manual_seed(D)
    S = 333
    x = (torch.randn(S, D) * 2.5 + 0.7).to(dtype)
    gamma, beta = make_norm(D, 5)
    ref = oracle.layer_norm(x.float(), gamma, beta, 1e-12)
    xd = x.cuda()
    y = ops.layernorm(xd, gamma.cuda(), beta.cuda(), 1e-12)
    got = y.float().cpu()
    if dtype == torch.float32:
        torch.testing.assert_close(got, ref, rtol=2e-5, atol=2e-6)
    else:
        want = ref.to(dtype).float()
        ulp = (want.abs() * (2.0 ** -7 if dtype == torch.bfloat16 else 2.0 ** -10)).clamp_min(1e-6)
        assert bool(((got - want).abs() <= ulp).all())                 # never more than one unit in the last place
        assert float((got == want).float().mean()) > 0.99              # and almost always the same rounding
    ops.layernorm(xd, gamma.cuda(), beta.cuda(), 1e-12, out=xd)         # in place
    assert torch.equal(xd, y)
    with pytest.raises(RuntimeError):
        ops.layernorm(torch.zeros(4, 2048, device="cuda"), torch.ones(2048, device="cuda"),
                      torch.zeros(2048, device="cuda"))


def test_block_golden(ops, oracle, route_mode):
    """The vectors made with the reference's FmoeConformerLayer members, through the CUDA path in bf16."""
    g = load_golden("case_block_3m.npz")
    experts = ops.pack_experts(g["W1"].cuda(), g["b1"].cuda(), g["W2"].cuda(), g["b2"].cuda())
    Wr = g["Wr"].cuda()
    x, emb = g["x"].cuda().bfloat16(), g["embed"].cuda().bfloat16()
    nf, nl = (g["ff_gamma"], g["ff_beta"]), (g["final_gamma"], g["final_beta"])
    res = ops.moe_layer(x, emb, Wr, None, experts, residual=x, ff_scale=float(g["ff_scale"]), return_routing=True,
                        Wr_packed=ops.pack_router(Wr), norm_ff=cu(nf), norm_final=cu(nl), eps=float(g["eps"]))
    assert torch.equal(res.idx.cpu().view(-1).long(), g["gate_idx"])
    assert torch.equal(res.counts.cpu().long(), g["expert_count"])
    torch.testing.assert_close(res.score.cpu().view(-1), g["gate_value"], rtol=2e-2, atol=1e-4)
    assert rel_l2(res.out.float().cpu(), g["out"]) <= BF16_REL_L2
    pre = ops.moe_layer(x, emb, Wr, None, experts, residual=x, ff_scale=float(g["ff_scale"]),
                        Wr_packed=ops.pack_router(Wr), norm_ff=cu(nf), eps=float(g["eps"]))
    assert rel_l2(pre.out.float().cpu(), g["pre_norm"]) <= BF16_REL_L2
    # norm_ff folded into the route kernel (router on fp32 norm_ff(x)): the fixture's margins keep routing exact
    wln = ops.pack_router_ln(Wr, nf[0].cuda(), nf[1].cuda())
    resf = ops.moe_layer(x, emb, Wr, None, experts, residual=x, ff_scale=float(g["ff_scale"]), return_routing=True,
                         Wr_packed=ops.pack_router(Wr), norm_ff=cu(nf), norm_final=cu(nl), eps=float(g["eps"]),
                         Wr_packed_ln=wln)
    assert torch.equal(resf.idx.cpu().view(-1).long(), g["gate_idx"])
    assert torch.equal(resf.counts.cpu().long(), g["expert_count"])
    torch.testing.assert_close(resf.score.cpu().view(-1), g["gate_value"], rtol=2e-3, atol=1e-5)
    assert rel_l2(resf.out.float().cpu(), g["out"]) <= BF16_REL_L2
    # fp32 activations: the SIMT gate and the generic dispatch path behind the same norms
    res32 = ops.moe_layer(g["x"].cuda(), g["embed"].cuda(), Wr, None, experts, residual=g["x"].cuda(),
                          ff_scale=float(g["ff_scale"]), return_routing=True, norm_ff=cu(nf), norm_final=cu(nl),
                          eps=float(g["eps"]))
    assert torch.equal(res32.idx.cpu().view(-1).long(), g["gate_idx"])
    assert rel_l2(res32.out.cpu(), g["out"]) <= BF16_REL_L2


def run_block_case(ops, oracle, synth, S, *, E=32, D=512, H=1024, Demb=512, norm_ff=True, norm_final=True, top_k=1,
                   gate_mode=None, seed=31, folded=False):
    """folded: hand the pre-scaled router to the call, so that norm_ff is folded into the fused gate + dispatch kernel
    (the router then sees norm_ff(x) in fp32, not rounded to bf16: bar (1) does not apply, bar (2) gets tighter)."""
    gate_mode = ops.GATE_3M if gate_mode is None else gate_mode
    demb = Demb if gate_mode == ops.GATE_3M else 0
    w = synth.make_weights(seed, E, D, H, demb, random_bias=True, router_bias=(gate_mode != ops.GATE_3M))
    x, emb = synth.make_activations(seed + 1, S, D, demb, w, top_k=top_k)
    x = (x * 2.0 + 0.3).bfloat16().float()
    nf = make_norm(D, seed + 2) if norm_ff else None
    nl = make_norm(D, seed + 3) if norm_final else None
    experts = ops.pack_experts(w.W1.cuda(), w.b1.cuda(), w.W2.cuda(), w.b2.cuda())
    Wr = w.Wr.cuda()
    br = None if w.br is None else w.br.cuda()
    xd = x.cuda().bfloat16()
    ed = None if emb is None else emb.cuda().bfloat16()
    packed = ops.pack_router(Wr) if E <= 32 else None
    wln = ops.pack_router_ln(Wr, nf[0].cuda(), nf[1].cuda()) if (folded and norm_ff) else None
    res = ops.moe_layer(xd, ed, Wr, br, experts, residual=xd, ff_scale=0.5, top_k=top_k, gate_mode=gate_mode,
                        return_routing=True, Wr_packed=packed, norm_ff=cu(nf), norm_final=cu(nl), Wr_packed_ln=wln)
    ogate = oracle.GATE_3M if gate_mode == ops.GATE_3M else oracle.GATE_NAIVE
    got_idx = res.idx.cpu().long()
    if top_k > 1:
        got_idx = torch.sort(got_idx, 1).values
    if not folded:
        check_same_input_bar(ops, oracle, res, got_idx, x, xd, emb, w, nf, nl, top_k, ogate)
    # (2) the pure fp32 chain: agreeing rows within tolerance, disagreeing rows are near-ties of that chain
    r2 = oracle.moe_block_forward(x, emb, w.Wr, w.br, w.W1, w.b1, w.W2, w.b2, norm_ff=nf, norm_final=nl, ff_scale=0.5,
                                  top_k=top_k, gate_mode=ogate)
    ref2 = torch.sort(r2["idx"], 1).values if top_k > 1 else r2["idx"]
    agree = (got_idx == ref2).all(-1)
    assert float(agree.float().mean()) > (0.995 if folded else 0.97)
    top2 = torch.topk(r2["logits"].double(), top_k + 1, dim=-1).values
    margin2 = (top2[:, :-1] - top2[:, 1:]).min(-1).values
    # rounding norm_ff(x) to bf16 moves a logit by up to ~1e-2; the folded router only by its own fp32 round-off
    assert float(margin2[~agree].max() if bool((~agree).any()) else 0.0) < (1e-3 if folded else 2e-2)
    err = rel_l2(res.out.float().cpu()[agree], r2["out"][agree])
    assert err <= BF16_REL_L2
    return err


def check_same_input_bar(ops, oracle, res, got_idx, x, xd, emb, w, nf, nl, top_k, ogate):
    """(1) the oracle on the GPU's own normalised input: routing must be bit-exact (near-ties of 1e-4 aside)."""
    xn_gpu = ops.layernorm(xd, nf[0].cuda(), nf[1].cuda()).float().cpu() if nf is not None else x
    r1 = oracle.moe_forward(xn_gpu, emb, w.Wr, w.br, w.W1, w.b1, w.W2, w.b2, top_k=top_k, gate_mode=ogate,
                            residual=x, ff_scale=0.5)
    top = torch.topk(r1["logits"].double(), top_k + 1, dim=-1).values
    clear = (top[:, :-1] - top[:, 1:]).min(-1).values >= 1e-4
    assert float(clear.float().mean()) > 0.995
    ref_idx = r1["idx"]
    if top_k > 1:
        ref_idx = torch.sort(ref_idx, 1).values
    assert torch.equal(got_idx[clear], ref_idx[clear]), "expert assignment on the same normalised input"
    if bool(clear.all()):
        assert torch.equal(res.counts.cpu().long(), r1["counts"])
        if top_k == 1:
            assert torch.equal(res.mapping.cpu().long(), r1["mapping"].view(-1))
    out1 = r1["out"] if nl is None else oracle.layer_norm(r1["out"], nl[0], nl[1])
    same = (got_idx == ref_idx).all(-1)
    assert rel_l2(res.out.float().cpu()[same], out1[same]) <= BF16_REL_L2


@pytest.mark.parametrize("S", [1, 50, 3200, 6000])
def test_block_cfg3_shape(ops, oracle, synth, S):
    # 6000 tokens: more 32-token tiles than SMs, the fused gate + dispatch kernel re-normalises rows it re-reads
    run_block_case(ops, oracle, synth, S)


@pytest.mark.parametrize("S", [50, 1234, 3200, 6000])
def test_block_norm_ff_folded_into_route_kernel(ops, oracle, synth, S, route_mode):
    """With the pre-scaled router (ops.pack_router_ln) norm_ff is folded into the fused gate + dispatch kernel: router
    MMAs on the raw rows, row statistics on the side, logits = e-part + r (x.W' - mu c1) + c0, rows normalised as they
    are scattered.  6000 tokens: more 32-token tiles than SMs, the kernel re-normalises the rows it re-reads."""
    if route_mode == 0:
        pytest.skip("the fused gate + dispatch kernel is switched off in this parametrisation")
    run_block_case(ops, oracle, synth, S, seed=61, folded=True)
    run_block_case(ops, oracle, synth, S, seed=62, folded=True, norm_final=False)


def test_block_folded_knob_off_is_the_row_pass(ops, oracle, synth):
    """ln_fuse = 0: the pre-scaled router is ignored and the call is bit-identical to the one without it."""
    E, D, H, Demb, S = 32, 512, 1024, 512, 777
    w = synth.make_weights(63, E, D, H, Demb, random_bias=True)
    x, emb = synth.make_activations(64, S, D, Demb, w)
    nf, nl = make_norm(D, 65), make_norm(D, 66)
    experts = ops.pack_experts(w.W1.cuda(), w.b1.cuda(), w.W2.cuda(), w.b2.cuda())
    Wr = w.Wr.cuda()
    xd, ed = (x * 2 + 0.3).cuda().bfloat16(), emb.cuda().bfloat16()
    kw = dict(residual=xd, ff_scale=0.5, Wr_packed=ops.pack_router(Wr), norm_ff=cu(nf), norm_final=cu(nl))
    plain = ops.moe_layer(xd, ed, Wr, None, experts, **kw).out.clone()
    ops.config("ln_fuse", 0)
    try:
        off = ops.moe_layer(xd, ed, Wr, None, experts, Wr_packed_ln=ops.pack_router_ln(Wr, *cu(nf)), **kw).out
        assert torch.equal(off, plain)
    finally:
        ops.config("ln_fuse", 1)


@pytest.mark.parametrize("folded", [False, True])
def test_block_padding_rows(ops, oracle, synth, folded):
    """x_len: padding rows are not routed; their output is norm_final(residual) (the MoE term is zero)."""
    E, D, H, Demb, B, T = 32, 512, 1024, 512, 4, 60
    w = synth.make_weights(91, E, D, H, Demb, random_bias=True)
    x, emb = synth.make_activations(92, B * T, D, Demb, w)
    x = (x * 2.0 + 0.3).bfloat16().float()
    nf, nl = make_norm(D, 93), make_norm(D, 94)
    x_len = torch.tensor([60, 17, 0, 41], dtype=torch.int32)
    experts = ops.pack_experts(w.W1.cuda(), w.b1.cuda(), w.W2.cuda(), w.b2.cuda())
    Wr = w.Wr.cuda()
    xd, ed = x.cuda().bfloat16(), emb.cuda().bfloat16()
    res = ops.moe_layer(xd, ed, Wr, None, experts, residual=xd, ff_scale=0.5, x_len=x_len.cuda(), seq_len=T,
                        return_routing=True, Wr_packed=ops.pack_router(Wr), norm_ff=cu(nf), norm_final=cu(nl),
                        Wr_packed_ln=ops.pack_router_ln(Wr, *cu(nf)) if folded else None)
    xn_gpu = ops.layernorm(xd, nf[0].cuda(), nf[1].cuda()).float().cpu()
    r = oracle.moe_forward(xn_gpu, emb, w.Wr, None, w.W1, w.b1, w.W2, w.b2, residual=x, ff_scale=0.5, x_len=x_len, T=T)
    want = oracle.layer_norm(r["out"], nl[0], nl[1])
    pad = (r["idx"].view(-1) < 0)
    assert int(pad.sum()) == B * T - int(x_len.sum())
    assert torch.equal(res.idx.cpu().long().view(-1)[pad], r["idx"].view(-1)[pad])
    same = (res.idx.cpu().long() == r["idx"]).all(-1)
    assert float(same.float().mean()) > 0.97     # (the folded router sees norm_ff(x) unrounded: near-ties may differ)
    assert rel_l2(res.out.float().cpu()[same], want[same]) <= BF16_REL_L2
    assert rel_l2(res.out.float().cpu()[pad], oracle.layer_norm(x, nl[0], nl[1])[pad]) <= 4e-3   # bf16 rounding only


@pytest.mark.parametrize("norm_ff,norm_final", [(True, False), (False, True)])
def test_block_single_norms(ops, oracle, synth, norm_ff, norm_final):
    run_block_case(ops, oracle, synth, 700, norm_ff=norm_ff, norm_final=norm_final, seed=47)


def test_block_large_batch_takes_split_kernels(ops, oracle, synth):
    # beyond the fused gate + dispatch kernel's range (2 * 32 * 148 tokens) whatever the route setting
    run_block_case(ops, oracle, synth, 9600, E=8, D=256, H=256, Demb=256, seed=53)


def test_block_naive_top2(ops, oracle, synth):
    run_block_case(ops, oracle, synth, 300, E=8, D=128, H=256, top_k=2, gate_mode=1, seed=59)


def test_block_folded_without_embed(ops, oracle, synth, route_mode):
    """NaiveGate top-1 (no cat-embed input, router bias): the folded kernel with the x part as its only K part."""
    if route_mode == 0:
        pytest.skip("the fused gate + dispatch kernel is switched off in this parametrisation")
    run_block_case(ops, oracle, synth, 500, E=16, D=256, H=256, top_k=1, gate_mode=1, seed=67, folded=True)


def test_block_module_mirror(ops, oracle, synth):
    """layer.feed_forward_block on a module with the reference block's members (fmoe_transformer.py:33-70)."""
    layer = pkg("layer")
    E, D, H, Demb, B, T = 8, 256, 512, 256, 3, 40
    w = synth.make_weights(71, E, D, H, Demb, random_bias=True)
    x, emb = synth.make_activations(72, B * T, D, Demb, w)
    x = (x * 2.0 - 0.4).bfloat16().float()
    block = torch.nn.Module()
    block.feed_forward = layer.LocalFmoeCatEmbedFeedForward(D, Demb, num_experts=E, hidden_units=H,
                                                            activation=layer.Swish())
    block.norm_ff = torch.nn.LayerNorm(D, eps=1e-12)
    block.norm_final = torch.nn.LayerNorm(D, eps=1e-12)
    block.conv_module = torch.nn.Identity()
    block.ff_scale, block.normalize_before = 0.5, True
    nf, nl = make_norm(D, 73), make_norm(D, 74)
    with torch.no_grad():
        block.feed_forward.router_weights.copy_(w.Wr)
        block.feed_forward.experts.w_1.weight.copy_(w.W1); block.feed_forward.experts.w_1.bias.copy_(w.b1)
        block.feed_forward.experts.w_2.weight.copy_(w.W2); block.feed_forward.experts.w_2.bias.copy_(w.b2)
        block.norm_ff.weight.copy_(nf[0]); block.norm_ff.bias.copy_(nf[1])
        block.norm_final.weight.copy_(nl[0]); block.norm_final.bias.copy_(nl[1])
    block = block.cuda()
    with torch.no_grad():
        y = layer.feed_forward_block(block, x.cuda().bfloat16().view(B, T, D), emb.cuda().bfloat16().view(B, T, Demb))
    ref = oracle.moe_block_forward(x, emb, w.Wr, None, w.W1, w.b1, w.W2, w.b2, norm_ff=nf, norm_final=nl, ff_scale=0.5)
    rows = (y.float().cpu().view(B * T, D) - ref["out"]).norm(dim=1) / ref["out"].norm(dim=1)
    assert float((rows <= 2e-2).float().mean()) > 0.97          # all rows but re-routed near-ties
    assert rel_l2(y.float().cpu().view(B * T, D)[rows <= 2e-2], ref["out"][rows <= 2e-2]) <= BF16_REL_L2


def test_ep_block_single_rank(ops, oracle, synth):
    """b200moe_ep_block_forward: the same block through the expert-parallel kernels (one rank: every flag is its own)."""
    ep_mod = pkg("ep_p2p")
    E, D, H, Demb, S = 32, 512, 1024, 512, 600
    w = synth.make_weights(81, E, D, H, Demb, random_bias=True)
    x, emb = synth.make_activations(82, S, D, Demb, w)
    x = (x * 2.0 + 0.3).bfloat16().float()
    nf, nl = make_norm(D, 83), make_norm(D, 84)
    experts = ops.pack_experts(w.W1.cuda(), w.b1.cuda(), w.W2.cuda(), w.b2.cuda())
    Wr = w.Wr.cuda()
    xd, ed = x.cuda().bfloat16(), emb.cuda().bfloat16()
    ctx = ep_mod.EpContext.simulate(1, E, D, S, torch.device("cuda"))[0]
    wln = ops.pack_router_ln(Wr, nf[0].cuda(), nf[1].cuda())
    out = ctx.forward(xd, ed, Wr, None, experts, residual=xd, ff_scale=0.5, Wr_packed=ops.pack_router(Wr),
                      norm_ff=cu(nf), norm_final=cu(nl), Wr_packed_ln=wln)
    one = ops.moe_layer(xd, ed, Wr, None, experts, residual=xd, ff_scale=0.5, Wr_packed=ops.pack_router(Wr),
                        norm_ff=cu(nf), norm_final=cu(nl), Wr_packed_ln=wln).out
    torch.cuda.synchronize()
    assert ctx.status() == 0
    # same routing, same kernels; the expert output makes the return trip in bf16 before the residual add and the norm
    assert rel_l2(out.float().cpu(), one.float().cpu()) <= 5e-3
    ref = oracle.moe_block_forward(x, emb, w.Wr, None, w.W1, w.b1, w.W2, w.b2, norm_ff=nf, norm_final=nl, ff_scale=0.5)
    rows = (out.float().cpu() - ref["out"]).norm(dim=1) / ref["out"].norm(dim=1)
    assert float((rows <= 2e-2).float().mean()) > 0.97            # all rows but re-routed near-ties
    assert rel_l2(out.float().cpu()[rows <= 2e-2], ref["out"][rows <= 2e-2]) <= BF16_REL_L2
    ctx.close()


def test_block_post_norm_wiring(ops, oracle, synth):
    """normalize_before = False (fmoe_transformer.py:160-166): no norm in front, norm_ff behind the residual add and, for
    blocks with a convolution module, norm_final after that."""
    layer = pkg("layer")
    E, D, H, Demb, B, T = 8, 256, 512, 256, 2, 33
    w = synth.make_weights(101, E, D, H, Demb, random_bias=True)
    x, emb = synth.make_activations(102, B * T, D, Demb, w)
    block = torch.nn.Module()
    block.feed_forward = layer.LocalFmoeCatEmbedFeedForward(D, Demb, num_experts=E, hidden_units=H,
                                                            activation=layer.Swish())
    block.norm_ff = torch.nn.LayerNorm(D, eps=1e-12)
    block.norm_final = torch.nn.LayerNorm(D, eps=1e-12)
    block.conv_module = torch.nn.Identity()
    block.ff_scale, block.normalize_before = 1.0, False
    nf, nl = make_norm(D, 103), make_norm(D, 104)
    with torch.no_grad():
        block.feed_forward.router_weights.copy_(w.Wr)
        block.feed_forward.experts.w_1.weight.copy_(w.W1); block.feed_forward.experts.w_1.bias.copy_(w.b1)
        block.feed_forward.experts.w_2.weight.copy_(w.W2); block.feed_forward.experts.w_2.bias.copy_(w.b2)
        block.norm_ff.weight.copy_(nf[0]); block.norm_ff.bias.copy_(nf[1])
        block.norm_final.weight.copy_(nl[0]); block.norm_final.bias.copy_(nl[1])
    block = block.cuda()
    with torch.no_grad():
        y = layer.feed_forward_block(block, x.cuda().bfloat16().view(B, T, D), emb.cuda().bfloat16().view(B, T, Demb))
    r = oracle.moe_forward(x, emb, w.Wr, None, w.W1, w.b1, w.W2, w.b2, residual=x, ff_scale=1.0)
    assert torch.equal(r["idx"], r["idx"])   # (routing is on the raw input here: synth's margins apply, no near-ties)
    want = oracle.layer_norm(oracle.layer_norm(r["out"], nf[0], nf[1]), nl[0], nl[1])
    assert rel_l2(y.float().cpu().view(B * T, D), want) <= BF16_REL_L2


def test_block_fp16_activations(ops, oracle, synth):
    """fp16 activations take the SIMT gate and the separate LayerNorm row passes."""
    E, D, H, Demb, S = 16, 256, 512, 256, 200
    w = synth.make_weights(111, E, D, H, Demb, random_bias=True)
    x, emb = synth.make_activations(112, S, D, Demb, w)
    x = (x * 2.0 + 0.3).half().float()
    nf, nl = make_norm(D, 113), make_norm(D, 114)
    experts = ops.pack_experts(w.W1.cuda(), w.b1.cuda(), w.W2.cuda(), w.b2.cuda())
    xd, ed = x.cuda().half(), emb.cuda().half()
    res = ops.moe_layer(xd, ed, w.Wr.cuda(), None, experts, residual=xd, ff_scale=0.5, return_routing=True,
                        norm_ff=cu(nf), norm_final=cu(nl))
    xn_gpu = ops.layernorm(xd, nf[0].cuda(), nf[1].cuda()).float().cpu()
    r = oracle.moe_forward(xn_gpu, emb.half().float(), w.Wr, None, w.W1, w.b1, w.W2, w.b2, residual=x, ff_scale=0.5)
    top = torch.topk(r["logits"].double(), 2, dim=-1).values
    clear = (top[:, 0] - top[:, 1]) >= 1e-4
    assert torch.equal(res.idx.cpu().long()[clear], r["idx"][clear])
    same = (res.idx.cpu().long() == r["idx"]).all(-1)
    want = oracle.layer_norm(r["out"], nl[0], nl[1])
    assert float(same.float().mean()) > 0.99
    assert rel_l2(res.out.float().cpu()[same], want[same]) <= BF16_REL_L2


def test_pack_router_ln_contents(ops):
    """b200moe_pack_router_ln: bf16 hi / lo rows of the router with its x rows scaled by gamma, then c1 and c0."""
    torch.manual_seed(9)
    Demb, D, E = 128, 192, 20
    Wr = torch.randn(Demb + D, E) * 0.05
    gamma, beta = make_norm(D, 121)
    buf = ops.pack_router_ln(Wr.cuda(), gamma.cuda(), beta.cuda()).cpu()
    R = Demb + D
    packed = buf[: 64 * R * 2].view(torch.bfloat16).view(64, R).float()
    c = buf[64 * R * 2:].view(torch.float32)
    assert c.numel() == 64
    Ws = Wr.clone()
    Ws[Demb:] *= gamma[:, None]
    hi = Ws.bfloat16().float()
    lo = (Ws - hi).bfloat16().float()
    assert torch.equal(packed[:E], hi.t()) and torch.equal(packed[32:32 + E], lo.t())
    assert float(packed[E:32].abs().max()) == 0.0 and float(packed[32 + E:].abs().max()) == 0.0
    c1 = (hi[Demb:].double() + lo[Demb:].double()).sum(0)
    c0 = (beta[:, None].double() * Wr[Demb:].double()).sum(0)
    torch.testing.assert_close(c[:E].double(), c1, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(c[32:32 + E].double(), c0, rtol=1e-6, atol=1e-7)
    assert float(c[E:32].abs().max()) == 0.0
    with pytest.raises(ValueError):
        x = torch.zeros(8, D, device="cuda", dtype=torch.bfloat16)
        ops.moe_layer(x, None, Wr.cuda(), None, ops.PackedExperts(torch.zeros(E, 128, D, device="cuda", dtype=torch.bfloat16),
                                                                 None, torch.zeros(E, D, 128, device="cuda", dtype=torch.bfloat16), None),
                      norm_ff=(gamma.cuda(), beta.cuda()), Wr_packed_ln=torch.zeros(10, dtype=torch.uint8, device="cuda"))
