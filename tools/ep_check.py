#!/usr/bin/env python
"""Multi-GPU parity of the expert-parallel path (run under torchrun on a multi-GPU box):
the W-rank result on each rank's tokens must equal the single-GPU fused layer on the same tokens with all experts --
routing bit-exact, outputs within bf16 round-off.
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
    synth = importlib.import_module(PKG + ".synth")
    E, D, H, Demb = 32, 512, 1024, 512
    E_local = E // world
    w = synth.make_weights(777, E, D, H, Demb, random_bias=True)
    ok = True
    for S in (50 + 13 * rank, 3200, 1):
        x, embed = synth.make_activations(1000 * S + rank, S, D, Demb, w)
        xd, ed = x.to(dev).bfloat16(), embed.to(dev).bfloat16()
        Wr = w.Wr.to(dev)
        full = ops.pack_experts(w.W1.to(dev), w.b1.to(dev), w.W2.to(dev), w.b2.to(dev))
        sl = slice(rank * E_local, (rank + 1) * E_local)
        mine = ops.PackedExperts(full.W1[sl].contiguous(), full.b1[sl].contiguous(), full.W2[sl].contiguous(),
                                 full.b2[sl].contiguous())
        ref = ops.moe_layer(xd, ed, Wr, None, full, residual=xd, ff_scale=0.5, return_routing=True,
                            Wr_packed=ops.pack_router(Wr))
        out, idx, score, counts, mapping = ep.ep_moe_layer(
            xd, ed, Wr, None, mine, num_local_expert=E_local, top_k=1, gate_mode=ops.GATE_3M, act_type=ops.ACT_SILU,
            ff_scale=0.5, residual=xd, Wr_packed=ops.pack_router(Wr), return_routing=True)
        torch.cuda.synchronize()
        same_idx = torch.equal(idx, ref.idx)
        same_map = torch.equal(mapping, ref.mapping)
        err = float((out.float() - ref.out.float()).norm() / ref.out.float().norm())
        # MoE term alone
        out2 = ep.ep_moe_layer(xd, ed, Wr, None, mine, num_local_expert=E_local, ff_scale=1.0, residual=None,
                               Wr_packed=ops.pack_router(Wr))
        ref2 = ops.moe_layer(xd, ed, Wr, None, full, residual=None, ff_scale=1.0)
        err2 = float((out2.float() - ref2.out.float()).norm() / ref2.out.float().norm().clamp_min(1e-20))
        good = same_idx and same_map and err < 5e-3 and err2 < 1e-2
        ok = ok and good
        print(f"rank {rank} S={S}: routing {'bit-exact' if same_idx and same_map else 'DIFFERS'}, out rel-L2 {err:.2e}, "
              f"moe-term rel-L2 {err2:.2e} -> {'ok' if good else 'FAIL'}", flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
