"""Expert parallelism over peer-mapped memory, exercised on ONE GPU: `world` ranks live in one process, their symmetric
buffers are ordinary allocations of the same device, and the ranks are driven stage by stage (push, compute + push back,
combine) so that no kernel is queued before the kernels it waits for.  The kernels, flags, counts and buffer layout are
exactly those of the multi-process path (tools/ep_check.py runs that one under torchrun on a multi-GPU box).

Bar: every rank's routing (expert assignment, per-expert counts, scatter indices) bit-exact against the oracle, outputs
within the BF16 tolerance -- i.e. the W-rank result equals the 1-GPU result on the same tokens.
"""
import pytest
import torch

from conftest import pkg, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=[1, 0], ids=["route", "split"])
def route_mode(request, ops):
    """Every case runs with the fused gate + dispatch kernel (small batches take it) and with the separate kernels."""
    ops.config("route", request.param)
    yield
    ops.config("route", 1)

BF16_REL_L2 = 1e-2
E, D, H, DEMB = 32, 512, 1024, 512


def run_world(ops, ep_mod, synth, oracle, world, sizes, *, top_k=1, gate_mode=None, random_bias=True, seed=4242,
              layers=1, keep_expert_output=False, stage_seq=(1, 8, 2, 4), x_lens=None, fold=False):
    """x_lens: per rank an int32 tensor of valid lengths (sizes[r] = len(x_lens[r]) * T_r padded rows) or None.
    stage_seq: 1 gate + counts, 8 scatter, 2 expert FFN, 4 results (15 = everything in one call: one rank only).
    fold: outputs live in the symmetric buffers, so that top-1 calls with residual = x (or none) take the folded combine."""
    gate_mode = ops.GATE_3M if gate_mode is None else gate_mode
    dev = torch.device("cuda")
    E_local = E // world
    cap = max(max(sizes) * top_k, 1)
    ctxs = ep_mod.EpContext.simulate(world, E_local, D, cap, dev)
    demb = DEMB if gate_mode == ops.GATE_3M else 0   # NaiveGate has no cat-embed input (fmoe/gates.py:51-66)
    ws = [synth.make_weights(seed + li, E, D, H, demb, random_bias=random_bias, router_bias=(gate_mode != ops.GATE_3M))
          for li in range(layers)]
    acts = [synth.make_activations(seed * 7 + r, sizes[r], D, demb, ws[0], top_k=top_k) for r in range(world)]
    cur = [a[0].cuda().bfloat16() for a in acts]
    emb = [None if a[1] is None else a[1].cuda().bfloat16() for a in acts]
    ref_cur = [a[0] for a in acts]
    xl = [None] * world if x_lens is None else x_lens
    Ts = [None if xl[r] is None else sizes[r] // len(xl[r]) for r in range(world)]
    xl_dev = [None if t is None else t.cuda() for t in xl]
    last = None
    for li, w in enumerate(ws):
        Wr = w.Wr.cuda()
        br = None if w.br is None else w.br.cuda()
        packed = ops.pack_router(Wr)
        full = ops.pack_experts(w.W1.cuda(), w.b1.cuda(), w.W2.cuda(), w.b2.cuda())
        mine = [ops.PackedExperts(full.W1[r * E_local:(r + 1) * E_local].contiguous(),
                                  full.b1[r * E_local:(r + 1) * E_local].contiguous(),
                                  full.W2[r * E_local:(r + 1) * E_local].contiguous(),
                                  full.b2[r * E_local:(r + 1) * E_local].contiguous()) for r in range(world)]
        outs = ([ctxs[r].out_buffer(li & 1, sizes[r]) for r in range(world)] if fold
                else [torch.empty_like(c) for c in cur])
        rbufs = [(torch.empty(sizes[r], top_k, dtype=torch.int32, device=dev),
                  torch.empty(sizes[r], top_k, dtype=torch.float32, device=dev),
                  torch.empty(E, dtype=torch.int32, device=dev),
                  torch.empty(sizes[r] * top_k, dtype=torch.int32, device=dev)) for r in range(world)]
        # the MoE term alone first (no residual, ff_scale 1): the O(1) residual must not be able to mask an error in it
        moes = ([ctxs[r].out_buffer((li + 1) & 1, sizes[r]) for r in range(world)] if fold
                else [torch.empty_like(c) for c in cur])
        moe_keep = [None] * world
        for stage in stage_seq:
            for r in range(world):
                ctxs[r].forward(cur[r], emb[r], Wr, br, mine[r], residual=None, top_k=top_k, gate_mode=gate_mode,
                                ff_scale=1.0, out=moes[r], out_slot=((li + 1) & 1) if fold else None, Wr_packed=packed,
                                return_routing=True, stages=stage,
                                routing_bufs=rbufs[r], keep_expert_output=keep_expert_output, x_len=xl_dev[r],
                                seq_len=Ts[r])
        torch.cuda.synchronize()
        moe_keep = [m.clone() for m in moes]   # (fold: the slot is reused two layers on)
        for stage in stage_seq:
            for r in range(world):
                ctxs[r].forward(cur[r], emb[r], Wr, br, mine[r], residual=cur[r], top_k=top_k, gate_mode=gate_mode,
                                ff_scale=0.5, out=outs[r], out_slot=(li & 1) if fold else None, Wr_packed=packed,
                                return_routing=True, stages=stage,
                                routing_bufs=rbufs[r], keep_expert_output=keep_expert_output, x_len=xl_dev[r],
                                seq_len=Ts[r])
        torch.cuda.synchronize()
        for r in range(world):
            assert ctxs[r].status() == 0, f"rank {r}: a wait on a peer flag timed out (status {ctxs[r].status()})"
        last = []
        for r in range(world):
            # the oracle sees what the GPU saw: this layer's input is the previous layer's (bf16) GPU output
            xin = cur[r].float().cpu()
            ref = oracle.moe_forward(xin, None if emb[r] is None else emb[r].float().cpu(), w.Wr, w.br, w.W1, w.b1,
                                     w.W2, w.b2, top_k=top_k,
                                     gate_mode=(oracle.GATE_3M if gate_mode == ops.GATE_3M else oracle.GATE_NAIVE),
                                     residual=xin, ff_scale=0.5, keep_expert_output=keep_expert_output,
                                     x_len=xl[r], T=Ts[r])
            idx, score, counts, mapping = rbufs[r]
            if sizes[r] > 0:
                if top_k == 1:
                    assert torch.equal(idx.cpu().long().view(-1), ref["idx"].view(-1)), f"rank {r}: expert assignment"
                else:
                    assert torch.equal(torch.sort(idx.cpu().long(), 1).values, torch.sort(ref["idx"], 1).values)
                assert torch.equal(counts.cpu().long(), ref["counts"]), f"rank {r}: per-expert counts"
                if top_k == 1:
                    assert torch.equal(mapping.cpu().long(), ref["mapping"].view(-1)), f"rank {r}: scatter indices"
                err = rel_l2(outs[r].float().cpu(), ref["out"])
                moe_err = rel_l2(moe_keep[r].float().cpu(), ref["moe"])
                assert err <= BF16_REL_L2 and moe_err <= 2 * BF16_REL_L2, f"rank {r}: rel-L2 {err:.2e} / {moe_err:.2e}"
                last.append((err, moe_err))
        cur = [o.clone() for o in outs] if fold else outs
    for c in ctxs:
        c.close()
    return last


@pytest.fixture(scope="module")
def ep_mod(ops):
    return pkg("ep_p2p")


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_ep_matches_single_gpu(ops, ep_mod, synth, oracle, world):
    sizes = [50 + 13 * r for r in range(world)]
    run_world(ops, ep_mod, synth, oracle, world, sizes)


def test_ep_ragged_and_empty_ranks(ops, ep_mod, synth, oracle):
    # a rank without tokens still has to raise its flags; one token; a batch larger than a tile per (expert, source)
    run_world(ops, ep_mod, synth, oracle, 4, [0, 1, 700, 37])


def test_ep_cfg3_shape_two_layers(ops, ep_mod, synth, oracle):
    # cfg3's per-layer shape on every rank, two layers back to back through the same buffers (flag sequence numbers)
    run_world(ops, ep_mod, synth, oracle, 2, [3200, 3200], layers=2, random_bias=False)


def test_ep_cfg4_shape_ragged_batches(ops, ep_mod, synth, oracle):
    """BASELINE.json configs[3]: 256 utterances of 100-1000 frames over 8 ranks (32 each, padded per rank, x_len-masked),
    4 experts per rank."""
    g = torch.Generator().manual_seed(20260004)
    frames = torch.randint(100, 1001, (256,), generator=g)
    lens = (((frames - 1) // 2 - 1) // 2).to(torch.int32)
    x_lens = [lens[32 * r:32 * (r + 1)].contiguous() for r in range(8)]
    sizes = [32 * int(t.max()) for t in x_lens]
    run_world(ops, ep_mod, synth, oracle, 8, sizes, x_lens=x_lens, random_bias=False)


def test_ep_naive_top2(ops, ep_mod, synth, oracle):
    run_world(ops, ep_mod, synth, oracle, 4, [64, 100, 3, 129], top_k=2, gate_mode=ops.GATE_NAIVE)


def test_ep_keep_expert_output(ops, ep_mod, synth, oracle):
    run_world(ops, ep_mod, synth, oracle, 2, [77, 50], keep_expert_output=True)


@pytest.mark.parametrize("S", [50, 3200])
def test_ep_whole_layer_in_one_call(ops, ep_mod, synth, oracle, S):
    """stages = 7, the call a multi-process rank makes: the dispatch kernel's last CTA itself waits for the arrival
    flags and builds the group table (no separate wait kernel).  With one rank every flag it waits for is its own."""
    run_world(ops, ep_mod, synth, oracle, 1, [S], layers=2, stage_seq=(15,))


@pytest.mark.parametrize("world,sizes", [(1, [3200]), (2, [50, 63]), (8, [400 + 7 * r for r in range(8)]),
                                         (4, [0, 1, 700, 37])])
def test_ep_folded_combine(ops, ep_mod, synth, oracle, world, sizes):
    """Outputs inside the symmetric buffers + residual = x: the owners' second-GEMM epilogue writes the finished rows
    residual + ff_scale * score * y into the source rank's `out`; no combine kernel runs."""
    n0 = ops.launch_count()
    run_world(ops, ep_mod, synth, oracle, world, sizes, layers=2, fold=True,
              stage_seq=(15,) if world == 1 else (1, 8, 2, 4))
    assert ops.launch_count() > n0


def test_ep_folded_cfg4_shape(ops, ep_mod, synth, oracle):
    g = torch.Generator().manual_seed(20260004)
    frames = torch.randint(100, 1001, (256,), generator=g)
    lens = (((frames - 1) // 2 - 1) // 2).to(torch.int32)
    x_lens = [lens[32 * r:32 * (r + 1)].contiguous() for r in range(8)]
    sizes = [32 * int(t.max()) for t in x_lens]
    run_world(ops, ep_mod, synth, oracle, 8, sizes, x_lens=x_lens, random_bias=False, fold=True)


def test_ep_stalled_peer_poisons_the_output(ops, ep_mod, synth):
    """A peer that never shows up must not pass for a result: the waits time out (bounded spin), the status word says
    which one, and the layer's output is NaN instead of stale rows."""
    dev = torch.device("cuda")
    world, S = 2, 40
    ctxs = ep_mod.EpContext.simulate(world, E // world, D, S, dev, timeout_ms=100)
    w = synth.make_weights(991, E, D, H, DEMB)
    x, embed = synth.make_activations(992, S, D, DEMB, w)
    xd, ed, Wr = x.cuda().bfloat16(), embed.cuda().bfloat16(), w.Wr.cuda()
    full = ops.pack_experts(w.W1.cuda(), w.b1.cuda(), w.W2.cuda(), w.b2.cuda())
    El = E // world
    mine = ops.PackedExperts(full.W1[:El].contiguous(), full.b1[:El].contiguous(), full.W2[:El].contiguous(),
                             full.b2[:El].contiguous())
    out = torch.zeros_like(xd)
    for stage in (1, 8, 2, 4):   # rank 0 alone: rank 1 sends neither counts nor rows and returns nothing
        ctxs[0].forward(xd, ed, Wr, None, mine, residual=xd, ff_scale=0.5, out=out, Wr_packed=ops.pack_router(Wr),
                        stages=stage)
    torch.cuda.synchronize()
    assert ctxs[0].status() != 0
    assert torch.isnan(out.float()).all()
    for c in ctxs:
        c.close()
