"""Checkpoint handling either side of the MoE path: the reference's whole-model state dict,
sliced per expert-parallel rank and gathered back (model/conformer_fmoe_localComm_catEmbed_domain_acc_hier.py:236-273),
and the one-time device packing of the MoE layers."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, pkg, rel_l2

E_LOCAL, WORLD, D, H, DEMB = 4, 2, 64, 128, 64


class TinyEncoder(torch.nn.Module):
    """Two blocks with the reference's attribute names: `encoders.<i>.feed_forward` is the MoE module, `norm_ff` a
    plain parameter that must pass through the expert slicing untouched."""

    def __init__(self, layer, world_size, rank, seed=0):
        super().__init__()
        torch.manual_seed(seed)
        blocks = []
        for _ in range(2):
            b = torch.nn.Module()
            b.feed_forward = layer.LocalFmoeCatEmbedFeedForward(D, DEMB, num_experts=E_LOCAL, rank=rank,
                                                                world_size=world_size, hidden_units=H,
                                                                activation=layer.Swish(), rand_init_router=True)
            b.norm_ff = torch.nn.LayerNorm(D, eps=1e-12)
            blocks.append(b)
        self.encoders = torch.nn.ModuleList(blocks)


def whole_state(seed=7):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for i in range(2):
        p = f"encoders.{i}."
        sd[p + "feed_forward.router_weights"] = torch.randn(D + DEMB, E_LOCAL * WORLD, generator=g)
        sd[p + "feed_forward.experts.w_1.weight"] = torch.randn(E_LOCAL * WORLD, H, D, generator=g)
        sd[p + "feed_forward.experts.w_1.bias"] = torch.randn(E_LOCAL * WORLD, H, generator=g)
        sd[p + "feed_forward.experts.w_2.weight"] = torch.randn(E_LOCAL * WORLD, D, H, generator=g)
        sd[p + "feed_forward.experts.w_2.bias"] = torch.randn(E_LOCAL * WORLD, D, generator=g)
        sd[p + "norm_ff.weight"] = torch.randn(D, generator=g)
        sd[p + "norm_ff.bias"] = torch.randn(D, generator=g)
    return sd


def test_slice_experts_matches_reference_rule_and_loads():
    ck, layer = pkg("checkpoint"), pkg("layer")
    sd = whole_state()
    for rank in range(WORLD):
        local = ck.slice_experts(sd, rank, WORLD, E_LOCAL)
        assert list(local) == list(sd)                                    # same keys, same order
        for k, v in sd.items():
            if "experts" in k:
                assert torch.equal(local[k], v[rank * E_LOCAL:(rank + 1) * E_LOCAL]), k
            else:
                assert local[k] is v, k                                   # router and norms: whole, untouched
        m = TinyEncoder(layer, WORLD, rank)
        res = ck.load_state_dict_comm(m, sd, rank, WORLD, E_LOCAL)
        assert not res.missing_keys and not res.unexpected_keys
        assert torch.equal(m.encoders[1].feed_forward.experts.w_2.weight,
                           sd["encoders.1.feed_forward.experts.w_2.weight"][rank * E_LOCAL:(rank + 1) * E_LOCAL])
        assert tuple(m.encoders[0].feed_forward.router_weights.shape) == (D + DEMB, E_LOCAL * WORLD)
    assert ck.slice_experts(sd, 0, 1, E_LOCAL * WORLD) == sd              # world 1: pass-through (:263-264)
    with pytest.raises(ValueError):
        ck.slice_experts(sd, 0, 4, E_LOCAL)                               # 8 experts in the file != 4 * 4
    assert ck.moe_layer_prefixes(sd) == ["encoders.0.feed_forward.", "encoders.1.feed_forward."]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gather_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        ck = pkg("checkpoint")
        sd = whole_state()
        local = ck.slice_experts(sd, rank, world, E_LOCAL)
        back = ck.gather_experts(local, rank, world, E_LOCAL)
        ret[rank] = list(back) == list(sd) and all(torch.equal(back[k], sd[k]) for k in sd)
    finally:
        dist.destroy_process_group()


def test_gather_experts_inverts_slicing_over_two_ranks():
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_gather_worker, args=(WORLD, port, ret), nprocs=WORLD, join=True)
        assert dict(ret) == {0: True, 1: True}


@pytest.mark.gpu
def test_packed_checkpoint_layers_run_and_match_oracle(ops, oracle, synth):
    """A state dict in the reference's names -> pack_moe_layers -> the layer call, against the oracle on the same
    checkpoint tensors (values on the bf16 grid, so the cast in the packing is exact)."""
    ck = pkg("checkpoint")
    E, D_, H_, Demb = 32, 512, 1024, 512
    sd = {}
    ws = []
    for i in range(2):
        w = synth.make_weights(900 + i, E, D_, H_, Demb, random_bias=True)
        ws.append(w)
        p = f"encoders.{i}.feed_forward."
        sd[p + "router_weights"] = w.Wr
        sd[p + "experts.w_1.weight"], sd[p + "experts.w_1.bias"] = w.W1, w.b1
        sd[p + "experts.w_2.weight"], sd[p + "experts.w_2.bias"] = w.W2, w.b2
        sd[f"encoders.{i}.norm_ff.weight"] = torch.ones(D_)
    packed = ck.pack_moe_layers(sd, "cuda")
    assert list(packed) == ["encoders.0.feed_forward.", "encoders.1.feed_forward."]
    for i, (p, L) in enumerate(packed.items()):
        x, emb = synth.make_activations(77 + i, 300, D_, Demb, ws[i])
        res = ops.moe_layer(x.cuda().bfloat16(), emb.cuda().bfloat16(), L["Wr"], L["br"], L["experts"],
                            residual=x.cuda().bfloat16(), ff_scale=0.5, Wr_packed=L["Wr_packed"], return_routing=True)
        ref = oracle.moe_forward(x, emb, ws[i].Wr, None, ws[i].W1, ws[i].b1, ws[i].W2, ws[i].b2, residual=x,
                                 ff_scale=0.5)
        assert torch.equal(res.idx.cpu().long().view(-1), ref["idx"].view(-1))
        assert rel_l2(res.out.float().cpu(), ref["out"]) <= 1e-2
    # expert-parallel slice of the same file: rank 1 of 2 holds experts 16..31
    half = ck.pack_moe_layers(sd, "cuda", rank=1, world_size=2)
    assert tuple(half["encoders.0.feed_forward."]["experts"].W1.shape) == (16, H_, D_)
    assert torch.equal(half["encoders.0.feed_forward."]["experts"].W1.float().cpu(), ws[0].W1[16:])
    assert tuple(half["encoders.0.feed_forward."]["Wr"].shape) == (D_ + Demb, E)
