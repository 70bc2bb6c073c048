#!/usr/bin/env python
"""FFN-kernel timeline of one expert-parallel layer on rank 0 (run under torchrun on a multi-GPU box):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 tools/ep_trace.py [S]
"""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
PKG = "3m-asr-inference_b200"


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 3200
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ops = importlib.import_module(PKG + ".ops")
    ep_p2p = importlib.import_module(PKG + ".ep_p2p")
    lib = importlib.import_module(PKG + "._lib").load()
    from ffn_trace import analyze
    E, D, H, Demb = 32, 512, 1024, 512
    El = E // world
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    gs = torch.Generator(device=dev).manual_seed(99)
    layers = []
    for _ in range(4):
        W1 = ((torch.rand(El, H, D, generator=g, device=dev) * 2 - 1) * 0.05).bfloat16()
        W2 = ((torch.rand(El, D, H, generator=g, device=dev) * 2 - 1) * 0.05).bfloat16()
        Wr = ((torch.rand(Demb + D, E, generator=gs, device=dev) * 2 - 1) * 0.04)
        layers.append((Wr, ops.PackedExperts(W1, torch.zeros(El, H, device=dev), W2, torch.zeros(El, D, device=dev)),
                       ops.pack_router(Wr)))
    x = torch.randn(S, D, generator=g, device=dev).bfloat16()
    emb = torch.randn(S, Demb, generator=g, device=dev).bfloat16()
    out = torch.empty_like(x)
    ctx = ep_p2p.EpContext.from_process_group(El, D, cap=S, timeout_ms=10000)
    for _ in range(3):
        for li, (Wr, ex, wp) in enumerate(layers):
            ctx.forward(x, emb, Wr, None, ex, residual=x, ff_scale=0.5, out_slot=li & 1, Wr_packed=wp, wait=li == 3)
    torch.cuda.synchronize()
    dist.barrier()
    cap, n_cta = 256, 148
    buf = torch.zeros(n_cta * cap, 4, dtype=torch.int32, device=dev)
    rbuf = torch.zeros(148 * 16, 4, dtype=torch.int32, device=dev)
    if rank == 0:
        lib.b200moe_debug_ffn_trace(buf.data_ptr(), cap)
        lib.b200moe_debug_route_trace(rbuf.data_ptr())
    Wr, ex, wp = layers[0]
    ctx.forward(x, emb, Wr, None, ex, residual=x, ff_scale=0.5, out_slot=0, Wr_packed=wp)
    torch.cuda.synchronize()
    if rank == 0:
        lib.b200moe_debug_ffn_trace(None, 0)
        lib.b200moe_debug_route_trace(None)
        from route_trace import analyze as route_analyze
        route_analyze(rbuf.cpu().numpy().astype(np.int64).reshape(148, 16, 4), S)
        rec = buf.cpu().numpy().astype(np.int64).reshape(n_cta, 4, cap // 4, 4)
        analyze(rec, cap, n_cta, S, 2)
    print(f"rank {rank} status {ctx.status()}", flush=True)
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
