"""Routing invariants of the oracle (the definitions the CUDA dispatch is held to, bit for bit)."""
import torch


def test_prepare_is_stable_counting_sort(oracle):
    g = torch.Generator().manual_seed(3)
    E = 32
    idx = torch.randint(0, E, (1000,), generator=g)
    p = oracle.prepare(idx, E)
    assert int(p["counts"].sum()) == 1000
    assert torch.equal(p["offsets"][1:] - p["offsets"][:-1], p["counts"])
    # mapping is a permutation, pos its inverse
    assert torch.equal(torch.sort(p["mapping"]).values, torch.arange(1000))
    assert torch.equal(p["mapping"][p["pos"]], torch.arange(1000))
    # expert-contiguous and stable: within an expert, rows keep entry order
    sorted_e = idx[p["pos"]]
    assert torch.all(sorted_e[1:] >= sorted_e[:-1])
    same = sorted_e[1:] == sorted_e[:-1]
    assert torch.all(p["pos"][1:][same] > p["pos"][:-1][same])
    # equals the closed form  mapping[i] = offsets[e_i] + #{j < i : e_j == e_i}
    seen = [0] * E
    for i, e in enumerate(idx.tolist()):
        assert int(p["mapping"][i]) == int(p["offsets"][e]) + seen[e]
        seen[e] += 1


def test_prepare_drops_invalid_entries(oracle):
    idx = torch.tensor([3, -1, 0, 3, 7, -1, 0])
    p = oracle.prepare(idx, 4)  # 7 is out of range for E = 4, -1 is padding
    assert p["counts"].tolist() == [2, 0, 0, 2]
    assert p["mapping"].tolist() == [2, -1, 0, 3, -1, -1, 1]
    assert p["pos"].tolist() == [2, 6, 0, 3]


def test_empty_and_single(oracle):
    p = oracle.prepare(torch.zeros(0, dtype=torch.long), 4)
    assert p["counts"].tolist() == [0, 0, 0, 0] and p["pos"].numel() == 0
    p = oracle.prepare(torch.tensor([2]), 4)
    assert p["offsets"].tolist() == [0, 0, 0, 1, 1] and p["mapping"].tolist() == [0]


def test_gate_tie_goes_to_lowest_index(oracle):
    x = torch.zeros(3, 8)
    Wr = torch.zeros(8, 5)
    idx, value, _ = oracle.gate_3m(x, None, Wr)
    assert idx.tolist() == [0, 0, 0]
    torch.testing.assert_close(value, torch.full((3,), 0.2))
    order, score, _ = oracle.gate_naive(x, Wr, None, 2)
    assert order.tolist() == [[0, 1]] * 3
    torch.testing.assert_close(score, torch.full((3, 2), 0.5))


def test_padding_rows_return_residual(oracle, synth):
    E, D, H, Demb, B, T = 4, 32, 64, 32, 3, 5
    w = synth.make_weights(11, E, D, H, Demb)
    x, embed = synth.make_activations(12, B * T, D, Demb, w)
    x_len = torch.tensor([5, 2, 0])
    r = oracle.moe_forward(x, embed, w.Wr, None, w.W1, w.b1, w.W2, w.b2, residual=x, ff_scale=0.5, x_len=x_len, T=T)
    pad = torch.tensor([t >= int(x_len[b]) for b in range(B) for t in range(T)])
    assert torch.equal(r["out"][pad], x[pad])
    assert torch.all(r["idx"][pad] == -1)
    assert int(r["counts"].sum()) == int((~pad).sum())


def test_output_is_linear_in_score_and_scale(oracle, synth):
    E, D, H, Demb, S = 4, 32, 64, 32, 17
    w = synth.make_weights(21, E, D, H, Demb, random_bias=True)
    x, embed = synth.make_activations(22, S, D, Demb, w)
    a = oracle.moe_forward(x, embed, w.Wr, None, w.W1, w.b1, w.W2, w.b2, ff_scale=1.0)
    b = oracle.moe_forward(x, embed, w.Wr, None, w.W1, w.b1, w.W2, w.b2, ff_scale=0.5, residual=x)
    torch.testing.assert_close(b["out"], x + 0.5 * a["out"], rtol=1e-6, atol=1e-6)
    k = oracle.moe_forward(x, embed, w.Wr, None, w.W1, w.b1, w.W2, w.b2, keep_expert_output=True)
    torch.testing.assert_close(a["out"], k["out"] * a["score"], rtol=1e-6, atol=1e-6)
