#!/usr/bin/env python
"""Multi-GPU parity of the expert-parallel paths (run under torchrun on a multi-GPU box):
the W-rank result on each rank's tokens must equal the single-GPU fused layer on the same tokens with all experts --
routing bit-exact, outputs within bf16 round-off.  Checks the peer-memory path (ep_p2p.EpContext, the product) and the
NCCL formulation (ep.ep_moe_layer).
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ep_check.py
"""
import importlib
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "3m-asr-inference_b200"


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ops = importlib.import_module(PKG + ".ops")
    ep = importlib.import_module(PKG + ".ep")
    ep_p2p = importlib.import_module(PKG + ".ep_p2p")
    synth = importlib.import_module(PKG + ".synth")
    E, D, H, Demb = 32, 512, 1024, 512
    E_local = E // world
    w = synth.make_weights(777, E, D, H, Demb, random_bias=True)
    sizes = (50 + 13 * rank, 3200, 1 if rank else 0)
    ctx = ep_p2p.EpContext.from_process_group(E_local, D, cap=8192, timeout_ms=5000)
    ok = True
    Wr = w.Wr.to(dev)
    Wrp = ops.pack_router(Wr)
    full = ops.pack_experts(w.W1.to(dev), w.b1.to(dev), w.W2.to(dev), w.b2.to(dev))
    sl = slice(rank * E_local, (rank + 1) * E_local)
    mine = ops.PackedExperts(full.W1[sl].contiguous(), full.b1[sl].contiguous(), full.W2[sl].contiguous(),
                             full.b2[sl].contiguous())
    for S in sizes:
        x, embed = synth.make_activations(1000 * S + rank, S, D, Demb, w)
        xd, ed = x.to(dev).bfloat16(), embed.to(dev).bfloat16()
        if S > 0:
            ref = ops.moe_layer(xd, ed, Wr, None, full, residual=xd, ff_scale=0.5, return_routing=True, Wr_packed=Wrp)
            ref2 = ops.moe_layer(xd, ed, Wr, None, full, residual=None, ff_scale=1.0, Wr_packed=Wrp)
        for name in ("p2p", "nccl"):
            if name == "p2p":
                out, idx, score, counts, mapping = ctx.forward(xd, ed, Wr, None, mine, residual=xd, ff_scale=0.5,
                                                               Wr_packed=Wrp, return_routing=True)
                out2 = ctx.forward(xd, ed, Wr, None, mine, residual=None, ff_scale=1.0, Wr_packed=Wrp)
            else:
                out, idx, score, counts, mapping = ep.ep_moe_layer(
                    xd, ed, Wr, None, mine, num_local_expert=E_local, top_k=1, gate_mode=ops.GATE_3M,
                    act_type=ops.ACT_SILU, ff_scale=0.5, residual=xd, Wr_packed=Wrp, return_routing=True)
                out2 = ep.ep_moe_layer(xd, ed, Wr, None, mine, num_local_expert=E_local, ff_scale=1.0, residual=None,
                                       Wr_packed=Wrp)
            torch.cuda.synchronize()
            if S == 0:
                print(f"rank {rank} {name} S=0: participated", flush=True)
                continue
            same_idx = torch.equal(idx, ref.idx)
            same_map = torch.equal(mapping, ref.mapping)
            err = float((out.float() - ref.out.float()).norm() / ref.out.float().norm())
            err2 = float((out2.float() - ref2.out.float()).norm() / ref2.out.float().norm().clamp_min(1e-20))
            good = same_idx and same_map and err < 5e-3 and err2 < 1e-2
            ok = ok and good
            print(f"rank {rank} {name} S={S}: routing {'bit-exact' if same_idx and same_map else 'DIFFERS'}, "
                  f"out rel-L2 {err:.2e}, moe-term rel-L2 {err2:.2e} -> {'ok' if good else 'FAIL'}", flush=True)
    st = ctx.status()
    if st != 0:
        print(f"rank {rank}: peer-flag wait timed out, status {st}", flush=True)
        ok = False
    # many layers back to back without any host synchronisation (flag sequence numbers, buffer reuse)
    S = 3200
    x, embed = synth.make_activations(5 + rank, S, D, Demb, w)
    xd, ed = x.to(dev).bfloat16(), embed.to(dev).bfloat16()
    bufs = [torch.empty_like(xd), torch.empty_like(xd)]
    cur = xd
    ref_cur = xd
    for li in range(12):
        cur = ctx.forward(cur, ed, Wr, None, mine, residual=cur, ff_scale=0.5, Wr_packed=Wrp, out=bufs[li & 1])
    for li in range(12):
        ref_cur = ops.moe_layer(ref_cur, ed, Wr, None, full, residual=ref_cur, ff_scale=0.5, Wr_packed=Wrp).out
    torch.cuda.synchronize()
    err = float((cur.float() - ref_cur.float()).norm() / ref_cur.float().norm())
    good = err < 2e-2 and ctx.status() == 0
    ok = ok and good
    print(f"rank {rank} p2p 12 layers back to back: rel-L2 vs single-GPU chain {err:.2e}, status {ctx.status()} -> "
          f"{'ok' if good else 'FAIL'}", flush=True)
    # the same chain with the combine folded into the owners' epilogue: outputs in the symmetric buffers, the wait for the
    # owners' rows deferred to the next layer's call (the benchmarked configuration)
    cur = xd
    for li in range(12):
        cur = ctx.forward(cur, ed, Wr, None, mine, residual=cur, ff_scale=0.5, Wr_packed=Wrp,
                          out_slot=li & 1, wait=li == 11)
    torch.cuda.synchronize()
    err = float((cur.float() - ref_cur.float()).norm() / ref_cur.float().norm())
    good = err < 2e-2 and ctx.status() == 0
    ok = ok and good
    print(f"rank {rank} p2p folded combine, 12 layers back to back: rel-L2 vs single-GPU chain {err:.2e}, status "
          f"{ctx.status()} -> {'ok' if good else 'FAIL'}", flush=True)
    for S2 in (50 + 13 * rank, 3200, 7000 + 100 * rank, 1 if rank else 0):
        x2, e2 = synth.make_activations(31 * S2 + rank, S2, D, Demb, w)
        x2d, e2d = x2.to(dev).bfloat16(), e2.to(dev).bfloat16()
        o2, idx2, _sc2, _c2, map2 = ctx.forward(x2d, e2d, Wr, None, mine, residual=x2d, ff_scale=0.5, Wr_packed=Wrp,
                                                out_slot=0, return_routing=True)
        torch.cuda.synchronize()
        if S2 == 0:
            continue
        r2 = ops.moe_layer(x2d, e2d, Wr, None, full, residual=x2d, ff_scale=0.5, return_routing=True, Wr_packed=Wrp)
        err2 = float((o2.float() - r2.out.float()).norm() / r2.out.float().norm())
        same = torch.equal(idx2, r2.idx) and torch.equal(map2, r2.mapping)
        good = same and err2 < 5e-3
        ok = ok and good
        print(f"rank {rank} p2p folded S={S2}: routing {'bit-exact' if same else 'DIFFERS'}, out rel-L2 {err2:.2e} -> "
              f"{'ok' if good else 'FAIL'}", flush=True)
    # the Conformer block's feed-forward part (norm_ff in the route kernel, norm_final in the combine kernel), 6 in a row
    gen = torch.Generator().manual_seed(4321)
    nf = tuple((t.to(dev)) for t in (1.0 + 0.2 * torch.randn(D, generator=gen), 0.1 * torch.randn(D, generator=gen)))
    nl = tuple((t.to(dev)) for t in (1.0 + 0.2 * torch.randn(D, generator=gen), 0.1 * torch.randn(D, generator=gen)))
    wln = ops.pack_router_ln(Wr, *nf)   # norm_ff folded into the route kernel on both sides
    cur = ref_cur = xd
    first = None
    for li in range(6):
        cur = ctx.forward(cur, ed, Wr, None, mine, residual=cur, ff_scale=0.5, Wr_packed=Wrp, norm_ff=nf, norm_final=nl,
                          Wr_packed_ln=wln)
        ref_cur = ops.moe_layer(ref_cur, ed, Wr, None, full, residual=ref_cur, ff_scale=0.5, Wr_packed=Wrp, norm_ff=nf,
                                norm_final=nl, Wr_packed_ln=wln).out
        if li == 0:
            first = float((cur.float() - ref_cur.float()).norm() / ref_cur.float().norm())
    torch.cuda.synchronize()
    err = float((cur.float() - ref_cur.float()).norm() / ref_cur.float().norm())
    good = first < 5e-3 and err < 5e-2 and ctx.status() == 0   # (a re-routed near-tie row differs wholesale later on)
    ok = ok and good
    print(f"rank {rank} p2p block (norm_ff + layer + norm_final): rel-L2 vs single GPU {first:.2e} after 1, {err:.2e} "
          f"after 6 -> {'ok' if good else 'FAIL'}", flush=True)
    # the module mirror (LocalFmoeCatEmbedFeedForward with world_size > 1) takes the same path
    layer_mod = importlib.import_module(PKG + ".layer")
    mod = layer_mod.LocalFmoeCatEmbedFeedForward(D, Demb, num_experts=E_local, rank=rank, world_size=world,
                                                 hidden_units=H, activation=layer_mod.Swish(), rand_init_router=True)
    mod = mod.to(dev)
    with torch.no_grad():
        mod.router_weights.copy_(w.Wr.to(dev))
        mod.experts.w_1.weight.copy_(w.W1[sl].to(dev))
        mod.experts.w_1.bias.copy_(w.b1[sl].to(dev))
        mod.experts.w_2.weight.copy_(w.W2[sl].to(dev))
        mod.experts.w_2.bias.copy_(w.b2[sl].to(dev))
    Bm, Tm = 4, 60 + rank
    xm, em = synth.make_activations(77 + rank, Bm * Tm, D, Demb, w)
    xm, em = xm.to(dev).bfloat16().view(Bm, Tm, D), em.to(dev).bfloat16().view(Bm, Tm, Demb)
    with torch.no_grad():
        om = mod(xm, em, residual=xm, ff_scale=0.5)
    refm = ops.moe_layer(xm.view(-1, D), em.view(-1, Demb), Wr, None, full, residual=xm.view(-1, D), ff_scale=0.5,
                         Wr_packed=Wrp).out
    torch.cuda.synchronize()
    errm = float((om.view(-1, D).float() - refm.float()).norm() / refm.float().norm())
    goodm = errm < 5e-3 and getattr(mod, "_ep_ctx", None) is not None and mod._ep_ctx.status() == 0
    ok = ok and goodm
    print(f"rank {rank} module LocalFmoeCatEmbedFeedForward(world_size={world}) over peer memory: rel-L2 {errm:.2e} -> "
          f"{'ok' if goodm else 'FAIL'}", flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    mod._ep_ctx.close()
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
