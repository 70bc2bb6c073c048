"""Pins the CPU oracle against vectors produced by the REFERENCE's own Python code (tests/golden/make_golden.py):
NaiveGate, moe_prepare_forward, MOEScatter, MOEbiasLinear, MOEGather, _fmoe_general_global_forward and the live 3M
router gate, run on CPU with only the absent fmoe_cuda primitives stubbed."""
import torch

from conftest import load_golden, load_golden_repo_dims, rel_l2


def test_3m_top1_matches_reference(oracle):
    g = load_golden("case_3m_top1.npz")
    r = oracle.moe_forward(g["x"], g["embed"], g["Wr"], None, g["W1"], g["b1"], g["W2"], g["b2"], top_k=1,
                           gate_mode=oracle.GATE_3M, act_type=oracle.ACT_SILU, residual=g["x"],
                           ff_scale=float(g["ff_scale"]))
    assert torch.equal(r["idx"].view(-1), g["gate_idx"])                      # routing: bit-exact
    assert torch.equal(r["counts"], g["expert_count"])                        # per-expert counts: bit-exact
    assert int(r["counts"].sum()) == int(g["fwd_batch_size"])
    torch.testing.assert_close(r["score"].view(-1), g["gate_value"], rtol=1e-5, atol=1e-7)
    # un-weighted expert output in token order == what FMoEExpertPlugin returns
    y_tok = r["ybuf"][r["mapping"].view(-1)]
    assert rel_l2(y_tok, g["expert_outputs"]) < 1e-6
    assert rel_l2(r["moe"], g["weighted"]) < 1e-6
    assert rel_l2(r["out"], g["final"]) < 1e-6


def test_3m_repo_dims_matches_reference(oracle, synth):
    """The same reference code at the repo's own dimensions (32 experts, 512 -> 1024 -> 512, cat-embed router)."""
    g, w = load_golden_repo_dims(synth)
    r = oracle.moe_forward(g["x"], g["embed"], w.Wr, None, w.W1, w.b1, w.W2, w.b2, top_k=1, gate_mode=oracle.GATE_3M,
                           act_type=oracle.ACT_SILU, residual=g["x"], ff_scale=float(g["ff_scale"]))
    assert torch.equal(r["idx"].view(-1), g["gate_idx"])
    assert torch.equal(r["counts"], g["expert_count"]) and int(r["counts"].sum()) == int(g["fwd_batch_size"])
    torch.testing.assert_close(r["score"].view(-1), g["gate_value"], rtol=1e-5, atol=1e-7)
    assert rel_l2(r["ybuf"][r["mapping"].view(-1)], g["expert_outputs"]) < 1e-6
    assert rel_l2(r["out"], g["final"]) < 1e-6


def test_naive_top2_matches_reference(oracle):
    g = load_golden("case_naive_top2.npz")
    r = oracle.moe_forward(g["x"], None, g["Wr"], g["br"], g["W1"], g["b1"], g["W2"], g["b2"], top_k=2,
                           gate_mode=oracle.GATE_NAIVE, act_type=oracle.ACT_GELU)
    # torch.topk(sorted=False) leaves the order of the k winners open: compare as sets, scores matched by expert
    ref_idx, ref_score = g["gate_idx"], g["gate_score"]
    assert torch.equal(torch.sort(r["idx"], dim=1).values, torch.sort(ref_idx, dim=1).values)
    for s in range(ref_idx.shape[0]):
        for j in range(2):
            jj = int((ref_idx[s] == r["idx"][s, j]).nonzero()[0])
            assert abs(float(r["score"][s, j]) - float(ref_score[s, jj])) < 1e-6
    assert torch.equal(r["counts"], g["expert_count"])
    assert rel_l2(r["out"], g["out"]) < 1e-6


def test_grouped_equals_dense_per_token(oracle, synth):
    """Grouping by expert must not change results: evaluate every token against its expert directly."""
    E, D, H, Demb, S = 8, 64, 96, 32, 53
    w = synth.make_weights(7, E, D, H, Demb, random_bias=True)
    x, embed = synth.make_activations(8, S, D, Demb, w)
    r = oracle.moe_forward(x, embed, w.Wr, None, w.W1, w.b1, w.W2, w.b2, residual=x, ff_scale=0.5)
    dense = torch.empty(S, D)
    for s in range(S):
        e = int(r["idx"][s, 0])
        h = oracle.activation(x[s] @ w.W1[e].t() + w.b1[e], oracle.ACT_SILU)
        dense[s] = x[s] + 0.5 * r["score"][s, 0] * (h @ w.W2[e].t() + w.b2[e])
    assert rel_l2(r["out"], dense) < 1e-6


def test_block_matches_reference(oracle):
    """norm_ff -> 3M MoE -> residual + ff_scale * y -> norm_final, with the LayerNorms, eps and ff_scale of the
    reference's own FmoeConformerLayer (layer/fmoe_transformer.py:54-65, 144-166)."""
    g = load_golden("case_block_3m.npz")
    r = oracle.moe_block_forward(g["x"], g["embed"], g["Wr"], None, g["W1"], g["b1"], g["W2"], g["b2"],
                                 norm_ff=(g["ff_gamma"], g["ff_beta"]), norm_final=(g["final_gamma"], g["final_beta"]),
                                 eps=float(g["eps"]), ff_scale=float(g["ff_scale"]))
    torch.testing.assert_close(r["xn"], g["xn"], rtol=1e-5, atol=1e-6)
    assert torch.equal(r["idx"].view(-1), g["gate_idx"])
    assert torch.equal(r["counts"], g["expert_count"])
    torch.testing.assert_close(r["score"].view(-1), g["gate_value"], rtol=1e-5, atol=1e-7)
    assert rel_l2(r["pre_norm_out"], g["pre_norm"]) < 1e-6
    assert rel_l2(r["out"], g["out"]) < 1e-6
    # the fixture is usable for bf16 parity with exact routing: every token's top-1 margin dwarfs a bf16 rounding of xn
    top = torch.topk(r["logits"].double(), 2, dim=-1).values
    assert float((top[:, 0] - top[:, 1]).min()) > 1e-2


def test_layer_norm_is_torch_layer_norm(oracle):
    torch.manual_seed(0)
    x = torch.randn(19, 96) * 3 + 1
    ln = torch.nn.LayerNorm(96, eps=1e-12)
    with torch.no_grad():
        ln.weight.normal_(1.0, 0.2)
        ln.bias.normal_(0.0, 0.1)
        torch.testing.assert_close(oracle.layer_norm(x, ln.weight, ln.bias, 1e-12), ln(x), rtol=1e-5, atol=1e-6)
