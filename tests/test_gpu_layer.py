"""GPU parity of the whole layer through b200moe_forward / the plugin enqueue / the module mirrors."""
import pytest
import torch

from conftest import load_golden, load_golden_repo_dims, pkg, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=[1, 0], ids=["route", "split"])
def route_mode(request, ops):
    """Every case runs with the fused gate + dispatch kernel (small batches take it) and with the separate kernels."""
    ops.config("route", request.param)
    yield
    ops.config("route", 1)

BF16_REL_L2 = 1e-2   # north_star tolerance for BF16 outputs vs the fp32 reference


def dev(t, dtype=None):
    if t is None:
        return None
    t = t.cuda()
    return t.to(dtype) if dtype is not None else t


def run_layer(ops, w, x, embed, *, dtype=torch.bfloat16, residual=True, x_len=None, T=None, tc_gate=True, **kw):
    """tc_gate: hand the packed router to the layer so that bf16 / E <= 32 calls take the tensor-core gate (and the
    gate-fused histograms); fp32 / fp16 activations and E > 32 use the SIMT gate either way."""
    experts = ops.pack_experts(dev(w.W1), dev(w.b1), dev(w.W2), dev(w.b2))
    xd, ed = dev(x, dtype), dev(embed, dtype)
    Wr = dev(w.Wr)
    packed = ops.pack_router(Wr) if (tc_gate and Wr.shape[1] <= 32) else None
    return ops.moe_layer(xd, ed, Wr, dev(w.br), experts, residual=xd if residual else None,
                         x_len=dev(x_len), seq_len=T, return_routing=True, Wr_packed=packed, **kw)


def check_against_oracle(oracle, res, ref, tol=BF16_REL_L2):
    assert torch.equal(res.idx.cpu().long(), ref["idx"])                       # expert assignment: bit-exact
    assert torch.equal(res.counts.cpu().long(), ref["counts"])                 # per-expert counts: bit-exact
    assert torch.equal(res.mapping.cpu().long(), ref["mapping"].view(-1))      # scatter indices: bit-exact
    torch.testing.assert_close(res.score.cpu(), ref["score"], rtol=2e-5, atol=1e-7)
    out = res.out.float().cpu().view(ref["out"].shape)
    assert torch.isfinite(out).all()
    assert rel_l2(out, ref["out"]) <= tol
    return rel_l2(out, ref["out"])


def test_golden_3m_top1(ops, oracle):
    """The vectors produced by the reference's own Python (tests/golden/make_golden.py), through the CUDA path."""
    g = load_golden("case_3m_top1.npz")

    class W:
        pass
    w = W()
    w.Wr, w.br, w.W1, w.b1, w.W2, w.b2 = g["Wr"], None, g["W1"], g["b1"], g["W2"], g["b2"]
    res = run_layer(ops, w, g["x"], g["embed"], dtype=torch.float32, ff_scale=float(g["ff_scale"]))
    assert torch.equal(res.idx.cpu().view(-1).long(), g["gate_idx"])
    assert torch.equal(res.counts.cpu().long(), g["expert_count"])
    torch.testing.assert_close(res.score.cpu().view(-1), g["gate_value"], rtol=2e-5, atol=1e-7)
    assert rel_l2(res.out.cpu(), g["final"]) <= BF16_REL_L2
    # the MoE contribution alone (residual removed), so the residual cannot mask an error
    moe = (res.out.cpu() - g["x"]) / float(g["ff_scale"])
    assert rel_l2(moe, g["weighted"]) <= 2 * BF16_REL_L2
    # un-weighted, no residual == FMoEExpertPlugin's output
    res2 = run_layer(ops, w, g["x"], g["embed"], dtype=torch.float32, residual=False, keep_expert_output=True)
    assert rel_l2(res2.out.cpu(), g["expert_outputs"]) <= BF16_REL_L2


def test_golden_3m_repo_dims(ops, oracle, synth):
    """Reference-generated vectors at the repo's own dimensions (E=32, D=512, H=1024) through the CUDA path, bf16."""
    g, w = load_golden_repo_dims(synth)
    res = run_layer(ops, w, g["x"], g["embed"], ff_scale=float(g["ff_scale"]))
    assert torch.equal(res.idx.cpu().view(-1).long(), g["gate_idx"])
    assert torch.equal(res.counts.cpu().long(), g["expert_count"])
    torch.testing.assert_close(res.score.cpu().view(-1), g["gate_value"], rtol=2e-5, atol=1e-7)
    assert rel_l2(res.out.float().cpu(), g["final"]) <= BF16_REL_L2
    res2 = run_layer(ops, w, g["x"], g["embed"], residual=False, keep_expert_output=True)
    assert rel_l2(res2.out.float().cpu(), g["expert_outputs"]) <= BF16_REL_L2


def test_cfg4_shape_ragged_batch(ops, oracle, synth):
    """BASELINE.json configs[3], one GPU's share: 32 utterances of 100-1000 frames (24-249 tokens after subsampling),
    padded to the longest, padding masked by x_len (rows >= x_len[b] are not routed, output = residual:
    softmax_topk_kernel.cu:40, fmoe_expert_plugin.cpp:244-245)."""
    E, D, H, Demb, B = 32, 512, 1024, 512, 32
    g = torch.Generator().manual_seed(20260004)
    frames = torch.randint(100, 1001, (B,), generator=g)
    x_len = (((frames - 1) // 2 - 1) // 2).to(torch.int32)
    T = int(x_len.max())
    S = B * T
    w = synth.make_weights(20260401, E, D, H, Demb, random_bias=True)
    x, embed = synth.make_activations(20260402, S, D, Demb, w)
    ref = oracle.moe_forward(x, embed, w.Wr, None, w.W1, w.b1, w.W2, w.b2, residual=x, ff_scale=0.5, x_len=x_len, T=T)
    res = run_layer(ops, w, x, embed, ff_scale=0.5, x_len=x_len, T=T)
    check_against_oracle(oracle, res, ref)
    pad = (torch.arange(S) % T) >= x_len.long().repeat_interleave(T)
    assert int(pad.sum()) > 0
    assert torch.equal(res.idx.cpu().view(-1)[pad], torch.full((int(pad.sum()),), -1, dtype=torch.int32))
    assert torch.equal(res.out.float().cpu()[pad], x[pad])            # padded rows: the residual, bit for bit
    res2 = run_layer(ops, w, x, embed, residual=False, ff_scale=1.0, x_len=x_len, T=T)
    assert rel_l2(res2.out.float().cpu(), ref["moe"]) <= 2 * BF16_REL_L2
    assert float(res2.out.float().cpu()[pad].abs().max()) == 0.0


def test_layer_next_to_a_foreign_kernel(ops, oracle, synth):
    """The fused gate + dispatch kernel waits for the other CTAs of its own grid; with foreign kernels holding SMs on
    another stream its CTAs become resident late.  The layer must still be right (no trap, status word clear)."""
    E, D, H, Demb, S = 32, 512, 1024, 512, 3200
    w = synth.make_weights(20260501, E, D, H, Demb)
    x, embed = synth.make_activations(20260502, S, D, Demb, w)
    ref = oracle.moe_forward(x, embed, w.Wr, None, w.W1, w.b1, w.W2, w.b2, residual=x, ff_scale=0.5)
    side = torch.cuda.Stream()
    a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
    with torch.cuda.stream(side):
        for _ in range(40):          # ~1.3 ms each: every SM busy with foreign CTAs for the whole test
            a @ a
    outs = [run_layer(ops, w, x, embed, ff_scale=0.5) for _ in range(8)]
    torch.cuda.synchronize()
    assert ops.status(clear=True) == 0
    for res in outs:
        check_against_oracle(oracle, res, ref)


def test_golden_naive_top2(ops, oracle):
    g = load_golden("case_naive_top2.npz")

    class W:
        pass
    w = W()
    w.Wr, w.br, w.W1, w.b1, w.W2, w.b2 = g["Wr"], g["br"], g["W1"], g["b1"], g["W2"], g["b2"]
    res = run_layer(ops, w, g["x"], None, dtype=torch.float32, residual=False, top_k=2, gate_mode=ops.GATE_NAIVE,
                    act_type=ops.ACT_GELU)
    assert torch.equal(torch.sort(res.idx.cpu().long(), 1).values, torch.sort(g["gate_idx"], 1).values)
    assert torch.equal(res.counts.cpu().long(), g["expert_count"])
    assert rel_l2(res.out.cpu(), g["out"]) <= BF16_REL_L2


@pytest.mark.parametrize("tc_gate", [True, False])
@pytest.mark.parametrize("S,dtype,random_bias", [
    (50, torch.bfloat16, False),     # cfg1: batch 1 x 206 frames -> 50 tokens, reference init (zero biases)
    (206, torch.bfloat16, True),     # cfg1, frames as tokens, biases exercised
    (3200, torch.bfloat16, True),    # cfg3 per layer: batch 64 x 206 frames
    (3200, torch.float32, True),     # fp32 activations at the boundary (the reference plugin's data_type 0)
    (13184, torch.bfloat16, False),  # cfg3 with frames as tokens
    (20000, torch.bfloat16, False),  # > 512 histogram rows: falls back to the count kernel
    (1, torch.bfloat16, True),
])
def test_layer_3m_repo_dims(ops, oracle, synth, S, dtype, random_bias, tc_gate):
    E, D, H, Demb = 32, 512, 1024, 512
    w = synth.make_weights(20260001, E, D, H, Demb, random_bias=random_bias)
    x, embed = synth.make_activations(20260001 + S, S, D, Demb, w)
    ref = oracle.moe_forward(x, embed, w.Wr, None, w.W1, w.b1, w.W2, w.b2, residual=x, ff_scale=0.5)
    res = run_layer(ops, w, x, embed, dtype=dtype, ff_scale=0.5, tc_gate=tc_gate)
    check_against_oracle(oracle, res, ref)
    # The MoE term on its own (no residual), so the O(1) residual cannot mask an error in the O(1e-2) expert output.
    # (With a bf16 residual stream the sum is rounded to 8 bits of mantissa, which is coarser than the MoE term for
    # the reference's tiny xavier(gain=0.5) weights -- that is a property of bf16 storage, not of this kernel.)
    res2 = run_layer(ops, w, x, embed, dtype=dtype, residual=False, ff_scale=0.5, tc_gate=tc_gate)
    assert rel_l2(res2.out.float().cpu(), 0.5 * ref["moe"]) <= BF16_REL_L2
    if dtype == torch.float32:
        moe = (res.out.cpu() - x) / 0.5
        assert rel_l2(moe, ref["moe"]) <= 2 * BF16_REL_L2


@pytest.mark.parametrize("S,E,D,H,Demb", [
    (333, 16, 256, 512, 128),    # smaller model, short embed part (two K parts of different length in the route kernel)
    (77, 8, 128, 256, 0),        # no embed input: single K part
    (1500, 32, 384, 640, 64),    # odd number of 128-row blocks in the first GEMM (no CTA pairs even at large tiles)
    (5000, 4, 512, 1024, 512),   # few experts, > 1 000 tokens each: 256-token tiles -> CTA pairs; two route tiles per SM
])
def test_layer_other_dims(ops, oracle, synth, S, E, D, H, Demb):
    """The reference leaves idim / hidden_units / num_expert to the plugin fields (fmoe_expert_plugin.cpp:325-366)."""
    w = synth.make_weights(7000 + S, E, D, H, Demb, random_bias=True)
    x, embed = synth.make_activations(7100 + S, S, D, Demb, w)
    ref = oracle.moe_forward(x, embed, w.Wr, None, w.W1, w.b1, w.W2, w.b2, residual=x, ff_scale=0.5)
    res = run_layer(ops, w, x, embed, ff_scale=0.5)
    check_against_oracle(oracle, res, ref)
    res2 = run_layer(ops, w, x, embed, residual=False, ff_scale=1.0)
    assert rel_l2(res2.out.float().cpu(), ref["moe"]) <= BF16_REL_L2


TF32_REL_L2 = 1e-3   # north_star tolerance for TF32 outputs vs the fp32 reference


@pytest.mark.parametrize("S,top_k", [(50, 1), (3200, 1), (777, 2)])
def test_layer_tf32(ops, oracle, synth, S, top_k):
    """fp32 activations + the reference's own fp32 expert weights (no packing), tensor cores in TF32: <= 1e-3.
    The expert weights are taken OFF the bf16 grid here (the bf16-grid weights of the other tests are exactly
    representable in TF32, which would hide the TF32 rounding of a real checkpoint)."""
    E, D, H = 32, 512, 1024
    naive = top_k > 1
    Demb = 0 if naive else 512
    # zero expert biases (the reference init): the output is pure GEMM result, nothing exact dilutes the error
    w = synth.make_weights(9000 + S, E, D, H, Demb, random_bias=False, router_bias=naive)
    g = torch.Generator().manual_seed(9100 + S)
    w.W1 = w.W1 * (1.0 + 1e-3 * torch.randn(w.W1.shape, generator=g))
    w.W2 = w.W2 * (1.0 + 1e-3 * torch.randn(w.W2.shape, generator=g))
    x, embed = synth.make_activations(9200 + S, S, D, Demb, w, top_k=top_k)
    # (x and the router stay on the grid: the routing margin the generator guarantees must not be disturbed)
    gm_o = oracle.GATE_NAIVE if naive else oracle.GATE_3M
    gm = ops.GATE_NAIVE if naive else ops.GATE_3M
    ref = oracle.moe_forward(x, embed, w.Wr, w.br, w.W1, w.b1, w.W2, w.b2, top_k=top_k, gate_mode=gm_o, residual=x,
                             ff_scale=0.5)
    experts = ops.fp32_experts(dev(w.W1), dev(w.b1), dev(w.W2), dev(w.b2))
    xd, ed = dev(x), dev(embed)
    res = ops.moe_layer(xd, ed, dev(w.Wr), dev(w.br), experts, residual=xd, top_k=top_k, gate_mode=gm, ff_scale=0.5,
                        return_routing=True, compute=ops.COMPUTE_TF32)
    if top_k == 1:
        assert torch.equal(res.idx.cpu().long(), ref["idx"])
        assert torch.equal(res.mapping.cpu().long(), ref["mapping"].view(-1))
    assert torch.equal(res.counts.cpu().long(), ref["counts"])
    assert rel_l2(res.out.cpu(), ref["out"]) <= TF32_REL_L2
    res2 = ops.moe_layer(xd, ed, dev(w.Wr), dev(w.br), experts, residual=None, top_k=top_k, gate_mode=gm, ff_scale=1.0,
                         compute=ops.COMPUTE_TF32)
    err = rel_l2(res2.out.cpu(), ref["moe"])
    assert err <= TF32_REL_L2, f"MoE term rel-L2 {err:.2e}"


def test_layer_padding_and_keep_output(ops, oracle, synth):
    E, D, H, Demb, B, T = 32, 512, 1024, 512, 6, 60
    w = synth.make_weights(31, E, D, H, Demb, random_bias=True)
    x, embed = synth.make_activations(32, B * T, D, Demb, w)
    x_len = torch.tensor([60, 33, 0, 1, 59, 24], dtype=torch.int32)
    ref = oracle.moe_forward(x, embed, w.Wr, None, w.W1, w.b1, w.W2, w.b2, residual=x, ff_scale=0.5, x_len=x_len, T=T,
                             keep_expert_output=True)
    res = run_layer(ops, w, x, embed, x_len=x_len, T=T, ff_scale=0.5, keep_expert_output=True)
    check_against_oracle(oracle, res, ref)
    pad = (torch.arange(B * T) % T) >= x_len.long().repeat_interleave(T)
    assert torch.equal(res.out.float().cpu()[pad], x[pad])   # padded rows: output == residual, exactly


@pytest.mark.parametrize("S,E,k,act", [(500, 32, 2, 2), (64, 8, 4, 1), (4000, 32, 2, 0)])
def test_layer_naive_topk(ops, oracle, synth, S, E, k, act):
    D, H = 512, 1024
    w = synth.make_weights(41 + S, E, D, H, 0, router_bias=True, random_bias=True)
    x, _ = synth.make_activations(42 + S, S, D, 0, w, top_k=k)
    ref = oracle.moe_forward(x, None, w.Wr, w.br, w.W1, w.b1, w.W2, w.b2, top_k=k, gate_mode=oracle.GATE_NAIVE,
                             act_type=act)
    res = run_layer(ops, w, x, None, residual=False, top_k=k, gate_mode=ops.GATE_NAIVE, act_type=act)
    check_against_oracle(oracle, res, ref)


def test_layer_skewed_router(ops, oracle, synth):
    """Zipf-like load through router_bias: a few experts take most tokens (several token tiles), many stay empty."""
    E, D, H, Demb, S = 32, 512, 1024, 512, 6000
    w = synth.make_weights(51, E, D, H, Demb, router_bias=True, random_bias=True)
    w.br = (torch.arange(E, 0, -1).float() * 0.35).bfloat16().float()
    x, embed = synth.make_activations(52, S, D, Demb, w)
    ref = oracle.moe_forward(x, embed, w.Wr, w.br, w.W1, w.b1, w.W2, w.b2, residual=x, ff_scale=0.5)
    assert int(ref["counts"].max()) > 1000 and int((ref["counts"] == 0).sum()) >= 1
    res = run_layer(ops, w, x, embed, ff_scale=0.5)
    check_against_oracle(oracle, res, ref)


def test_layer_properties_at_large_size(ops, synth):
    """cfg5-sized call (65 536 tokens) checked through size-independent properties: permuting the tokens permutes the
    outputs bit-for-bit, counts sum to S, and doubling ff_scale doubles the MoE term."""
    E, D, H, Demb, S = 32, 512, 1024, 512, 65536
    w = synth.make_weights(61, E, D, H, Demb, random_bias=True)
    g = torch.Generator().manual_seed(62)
    x = torch.randn(S, D, generator=g).bfloat16()
    embed = torch.randn(S, Demb, generator=g).bfloat16()
    experts = ops.pack_experts(dev(w.W1), dev(w.b1), dev(w.W2), dev(w.b2))
    Wr = dev(w.Wr)
    xd, ed = x.cuda(), embed.cuda()
    a = ops.moe_layer(xd, ed, Wr, None, experts, ff_scale=1.0, return_routing=True)
    a_out, a_idx, a_counts = a.out.clone(), a.idx.clone(), a.counts.clone()
    assert int(a_counts.sum()) == S
    perm = torch.randperm(S, generator=g).cuda()
    b = ops.moe_layer(xd[perm].contiguous(), ed[perm].contiguous(), Wr, None, experts, ff_scale=1.0,
                      return_routing=True)
    assert torch.equal(b.idx, a_idx[perm])
    assert torch.equal(b.out, a_out[perm])
    c = ops.moe_layer(xd, ed, Wr, None, experts, ff_scale=2.0)
    torch.testing.assert_close(c.out.float(), 2.0 * a_out.float(), rtol=1e-2, atol=1e-3)
    assert torch.isfinite(a_out.float()).all()


def test_plugin_enqueue_matches_reference_contract(ops, oracle, synth):
    """FMoEExpertPluginDynamic: six inputs, un-weighted output in token order (fmoe_expert_plugin.cpp:241-269)."""
    plugin = pkg("plugin")
    E, D, H, S = 32, 512, 1024, 206
    w = synth.make_weights(71, E, D, H, 0, random_bias=True)
    g = torch.Generator().manual_seed(72)
    x = torch.randn(1, S, D, generator=g).bfloat16().float()
    gate_idx = torch.randint(0, E, (1, S, 1), generator=g, dtype=torch.int32)
    creator = plugin.PluginRegistry().get_plugin_creator("FMoEExpertPluginDynamic", "1", "")
    p = creator.create_plugin("plugin", {"data_type": 0, "num_expert": E, "idim": D, "hidden_units": H})
    out = p.enqueue([x.cuda(), gate_idx.cuda(), w.W1.cuda(), w.b1.cuda(), w.W2.cuda(), w.b2.cuda()])
    assert out.shape == x.shape and out.dtype == torch.float32
    prep = oracle.prepare(gate_idx.view(-1), E)
    ybuf = oracle.expert_ffn(x.view(S, D)[prep["pos"]], prep["counts"], w.W1, w.b1, w.W2, w.b2, 0)
    ref = ybuf[prep["mapping"]]
    assert rel_l2(out.cpu().view(S, D), ref) <= BF16_REL_L2
    # second enqueue re-uses the packed weights; a clone built from the serialised fields gives the same answer
    out2 = p.enqueue([x.cuda(), gate_idx.cuda(), w.W1.cuda(), w.b1.cuda(), w.W2.cuda(), w.b2.cuda()])
    q = creator.deserialize_plugin("plugin", p.serialize())
    out3 = q.enqueue([x.cuda(), gate_idx.cuda(), w.W1.cuda(), w.b1.cuda(), w.W2.cuda(), w.b2.cuda()])
    assert rel_l2(out2.cpu(), out.cpu()) < 1e-6 and torch.equal(out3, out2)


def test_plugin_shared_scratch_and_changing_shapes(ops, oracle, synth):
    """TensorRT hands ONE scratch workspace to every layer of an engine, does not preserve it between enqueues and
    passes a different S per utterance: the plugin may keep nothing in it (its weight copies are plugin-owned)."""
    plugin = pkg("plugin")
    E, D, H = 8, 256, 512
    wa = synth.make_weights(171, E, D, H, 0, random_bias=True)
    wb = synth.make_weights(172, E, D, H, 0, random_bias=True)
    creator = plugin.PluginRegistry().get_plugin_creator("FMoEExpertPluginDynamic", "1", "")
    fields = {"data_type": 0, "num_expert": E, "idim": D, "hidden_units": H}
    pa, pb = creator.create_plugin("a", fields), creator.create_plugin("b", fields)
    ws = torch.empty(max(pa.get_workspace_size(200), pb.get_workspace_size(200)), dtype=torch.uint8, device="cuda")
    dev_a = [t.cuda() for t in (wa.W1, wa.b1, wa.W2, wa.b2)]
    dev_b = [t.cuda() for t in (wb.W1, wb.b1, wb.W2, wb.b2)]

    def check(p, w, dev_w, S, seed):
        g = torch.Generator().manual_seed(seed)
        x = torch.randn(1, S, D, generator=g).bfloat16().float()
        gate_idx = torch.randint(0, E, (1, S, 1), generator=g, dtype=torch.int32)
        out = p.enqueue([x.cuda(), gate_idx.cuda(), *dev_w], workspace=ws)
        prep = oracle.prepare(gate_idx.view(-1), E)
        ybuf = oracle.expert_ffn(x.view(S, D)[prep["pos"]], prep["counts"], w.W1, w.b1, w.W2, w.b2, 0)
        assert rel_l2(out.cpu().view(S, D), ybuf[prep["mapping"]]) <= BF16_REL_L2

    check(pa, wa, dev_a, 200, 1)
    check(pa, wa, dev_a, 50, 2)      # another S on the same workspace: the layout of the scratch moves
    check(pb, wb, dev_b, 200, 3)     # another layer's plugin scribbles over the shared scratch
    ws.fill_(0xA5)                   # ... and so may anything else between two enqueues
    check(pa, wa, dev_a, 200, 4)
    check(pb, wb, dev_b, 37, 5)
    check(pa, wa, dev_a, 50, 6)
    # weights updated in place behind the same pointers need an explicit invalidate
    dev_a[0].copy_(wb.W1.cuda()); dev_a[1].copy_(wb.b1.cuda()); dev_a[2].copy_(wb.W2.cuda()); dev_a[3].copy_(wb.b2.cuda())
    pa.invalidate()
    check(pa, wb, dev_a, 64, 7)


def test_module_mirrors(ops, oracle, synth):
    layer = pkg("layer")
    fmoe = pkg("fmoe")
    E, D, H, Demb, B, T = 32, 512, 1024, 512, 4, 50
    w = synth.make_weights(81, E, D, H, Demb, random_bias=True)
    x, embed = synth.make_activations(82, B * T, D, Demb, w)
    m = layer.LocalFmoeCatEmbedFeedForward(D, Demb, num_experts=E, hidden_units=H, activation=layer.Swish())
    m.load_state_dict({"router_weights": w.Wr, "experts.w_1.weight": w.W1, "experts.w_1.bias": w.b1,
                       "experts.w_2.weight": w.W2, "experts.w_2.bias": w.b2})
    m = m.cuda()
    xin = x.view(B, T, D).cuda().bfloat16()
    out = m(xin, embed.view(B, T, Demb).cuda().bfloat16(), None)
    ref = oracle.moe_forward(x, embed, w.Wr, None, w.W1, w.b1, w.W2, w.b2)
    assert rel_l2(out.float().cpu().view(B * T, D), ref["out"]) <= BF16_REL_L2
    out_r = m(xin, embed.view(B, T, Demb).cuda().bfloat16(), None, residual=xin, ff_scale=0.5)
    assert rel_l2(out_r.float().cpu().view(B * T, D), x + 0.5 * ref["out"]) <= BF16_REL_L2

    w2 = synth.make_weights(83, 8, 256, 512, 0, router_bias=True, random_bias=True)
    x2, _ = synth.make_activations(84, 300, 256, 0, w2, top_k=2)
    mlp = fmoe.FMoETransformerMLP(num_expert=8, d_model=256, d_hidden=512, top_k=2)
    mlp.load_state_dict({"gate.gate.weight": w2.Wr.t().contiguous(), "gate.gate.bias": w2.br,
                         "experts.htoh4.weight": w2.W1, "experts.htoh4.bias": w2.b1,
                         "experts.h4toh.weight": w2.W2, "experts.h4toh.bias": w2.b2})
    mlp = mlp.cuda()
    y = mlp(x2.view(3, 100, 256).cuda().bfloat16())
    ref2 = oracle.moe_forward(x2, None, w2.Wr, w2.br, w2.W1, w2.b1, w2.W2, w2.b2, top_k=2,
                              gate_mode=oracle.GATE_NAIVE, act_type=oracle.ACT_GELU)
    assert tuple(y.shape) == (3, 100, 256)
    assert rel_l2(y.float().cpu().view(300, 256), ref2["out"]) <= BF16_REL_L2


def test_registered_torch_op(ops, oracle, synth):
    """torch.ops.b200moe.fmoe_forward (the registered custom op) == the oracle's FMoE forward."""
    E, D, H, Demb, S = 32, 512, 1024, 512, 300
    w = synth.make_weights(8800, E, D, H, Demb, random_bias=True)
    x, embed = synth.make_activations(8801, S, D, Demb, w)
    ref = oracle.moe_forward(x, embed, w.Wr, None, w.W1, w.b1, w.W2, w.b2, residual=x, ff_scale=0.5)
    ex = ops.pack_experts(dev(w.W1), dev(w.b1), dev(w.W2), dev(w.b2))
    xd = dev(x, torch.bfloat16)
    out = torch.ops.b200moe.fmoe_forward(xd, dev(embed, torch.bfloat16), dev(w.Wr), None, ex.W1, ex.b1, ex.W2, ex.b2, xd,
                                         1, ops.GATE_3M, ops.ACT_SILU, 0.5)
    assert rel_l2(out.float().cpu(), ref["out"]) <= BF16_REL_L2
