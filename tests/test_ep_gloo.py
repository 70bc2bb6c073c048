"""Expert-parallel host logic on CPU: world_size 2 over gloo.  The arithmetic is supplied by the oracle through the
`backend` hook (tests may do that; the product backend is CUDA-only), so what is exercised here is exactly the code
that runs between the kernels on the GPU box: counts exchange, split sizes, all-to-all-v ordering, the second
(local-expert) dispatch, the inverse permutation and the final combine.  Bar: the W-rank result equals the 1-rank
oracle on the same tokens -- routing bit-exact, outputs to fp32 round-off."""
import importlib
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, pkg


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class OracleBackend:
    def __init__(self, oracle):
        self.o = oracle

    def gate(self, x, embed, Wr, br, x_len, top_k, gate_mode, seq_len, Wr_packed=None):
        if gate_mode == 0:
            idx, val, _ = self.o.gate_3m(x, embed, Wr, br)
            idx, score = idx.view(-1, 1), val.view(-1, 1)
        else:
            idx, score, _ = self.o.gate_naive(x, Wr, br, top_k)
        if x_len is not None:
            T = seq_len
            valid = (torch.arange(x.shape[0]) % T) < x_len.long().repeat_interleave(T)
            idx = torch.where(valid[:, None], idx, torch.full_like(idx, -1))
            score = torch.where(valid[:, None], score, torch.zeros_like(score))
        return idx.to(torch.int32), score.float()

    def dispatch(self, x, idx, num_expert, hidden=0):
        S = x.shape[0]
        k = idx.numel() // max(S, 1)
        p = self.o.prepare(idx.reshape(-1).long(), num_expert)
        xbuf = torch.zeros(S * k, x.shape[1], dtype=x.dtype)
        xbuf[: p["pos"].numel()] = x[p["pos"] // k]
        return p["counts"].to(torch.int32), p["offsets"].to(torch.int32), p["mapping"].to(torch.int32), xbuf

    def expert_ffn(self, xbuf, offsets, experts, act_type, out_dtype):
        counts = (offsets[1:] - offsets[:-1]).long()
        n = int(counts.sum())
        y = torch.zeros(xbuf.shape[0], experts["W2"].shape[1], dtype=out_dtype)
        y[:n] = self.o.expert_ffn(xbuf[:n].float(), counts, experts["W1"], experts["b1"], experts["W2"],
                                  experts["b2"], act_type).to(out_dtype)
        return y

    def combine(self, ybuf, mapping, score, residual, ff_scale, top_k):
        S = mapping.numel() // top_k
        m = mapping.view(S, top_k).long()
        rows = ybuf.float()[m.clamp_min(0)] * (m >= 0)[..., None]
        w = torch.ones(S, top_k) if score is None else score
        out = ff_scale * (w[..., None] * rows).sum(1)
        if residual is not None:
            out = residual.float() + out
        return out.to(ybuf.dtype)


def _worker(rank, world, port, case, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        oracle = importlib.import_module("oracle.moe_oracle")
        synth = pkg("synth")
        ep = pkg("ep")
        torch.set_num_threads(1)
        E_total, D, H, Demb, top_k, gate_mode, act = case
        E_local = E_total // world
        w = synth.make_weights(4242, E_total, D, H, Demb, router_bias=(gate_mode == 1), random_bias=True)
        S = 37 + 11 * rank                      # ragged: every rank has a different number of tokens
        x, embed = synth.make_activations(5000 + rank, S, D, Demb, w, top_k=top_k)
        x_len = None
        T = None
        if gate_mode == 0 and rank == 1:        # one rank also has padding rows
            T = S
            x_len = torch.tensor([S - 5], dtype=torch.int32)
        sl = slice(rank * E_local, (rank + 1) * E_local)
        local = {"W1": w.W1[sl], "b1": w.b1[sl], "W2": w.W2[sl], "b2": w.b2[sl]}
        out, idx, score, counts, mapping = ep.ep_moe_layer(
            x, embed, w.Wr, w.br, local, num_local_expert=E_local, top_k=top_k, gate_mode=gate_mode, act_type=act,
            ff_scale=0.5, residual=x, x_len=x_len, seq_len=T, backend=OracleBackend(oracle), return_routing=True)
        ref = oracle.moe_forward(x, embed, w.Wr, w.br, w.W1, w.b1, w.W2, w.b2, top_k=top_k, gate_mode=gate_mode,
                                 act_type=act, residual=x, ff_scale=0.5, x_len=x_len, T=T)
        ok_idx = torch.equal(idx.long(), ref["idx"])
        ok_counts = torch.equal(counts.long(), ref["counts"])
        ok_map = torch.equal(mapping.long(), ref["mapping"].view(-1))
        err = float((out.double() - ref["out"].double()).norm() / ref["out"].double().norm())
        ret[rank] = (ok_idx, ok_counts, ok_map, err)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", [
    (8, 64, 96, 32, 1, 0, 0),     # 3M router top-1, cat-embed, SiLU; 4 experts per rank
    (4, 32, 64, 0, 2, 1, 2),      # NaiveGate top-2, GELU; 2 experts per rank
])
def test_two_ranks_match_single_rank_oracle(case):
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, case, ret), nprocs=world, join=True)
        assert len(ret) == world
        for rank in range(world):
            ok_idx, ok_counts, ok_map, err = ret[rank]
            assert ok_idx and ok_counts and ok_map, f"rank {rank}: routing differs"
            assert err < 1e-5, f"rank {rank}: output rel-L2 {err}"


def test_received_expert_ids_and_splits():
    ep = pkg("ep")
    recv = torch.tensor([[2, 0, 1], [0, 3, 1]], dtype=torch.int32)   # [source rank, local expert]
    ids = ep.received_expert_ids(recv, 7)
    assert ids.tolist() == [0, 0, 2, 1, 1, 1, 2]
    send = torch.tensor([[1, 1, 0], [4, 0, 2]], dtype=torch.int32)
    s, r = ep.split_sizes(send, recv)
    assert s == [2, 6] and r == [3, 4]
