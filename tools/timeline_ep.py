#!/usr/bin/env python
"""tools/timeline.py for the expert-parallel layer: cross-kernel timeline of a graph replay of L layers on rank 0.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 tools/timeline_ep.py [S] [L]"""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "3m-asr-inference_b200"
MARKS = ["start", "wait_ok", "first_data", "last_req", "last_mma", "end"]


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 3200
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ops = importlib.import_module(PKG + ".ops")
    ep_p2p = importlib.import_module(PKG + ".ep_p2p")
    lib = importlib.import_module(PKG + "._lib").load()
    E, D, H, Demb = 32, 512, 1024, 512
    El = E // world
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    gs = torch.Generator(device=dev).manual_seed(99)
    layers = []
    for _ in range(L):
        W1 = ((torch.rand(El, H, D, generator=g, device=dev) * 2 - 1) * 0.05).bfloat16()
        W2 = ((torch.rand(El, D, H, generator=g, device=dev) * 2 - 1) * 0.05).bfloat16()
        Wr = ((torch.rand(Demb + D, E, generator=gs, device=dev) * 2 - 1) * 0.04)
        layers.append((Wr, ops.PackedExperts(W1, torch.zeros(El, H, device=dev), W2, torch.zeros(El, D, device=dev)),
                       ops.pack_router(Wr)))
    x0 = torch.randn(S, D, generator=g, device=dev).bfloat16()
    emb = torch.randn(S, Demb, generator=g, device=dev).bfloat16()
    ctx = ep_p2p.EpContext.from_process_group(El, D, cap=S, timeout_ms=10000)

    def step():
        cur = x0
        for li, (Wr, ex, wp) in enumerate(layers):
            cur = ctx.forward(cur, emb, Wr, None, ex, residual=cur, ff_scale=0.5, out_slot=li & 1, Wr_packed=wp,
                              wait=li == L - 1)
        return cur
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.synchronize()
    dist.barrier()
    n_slots = 3 * L + 4
    tl = torch.zeros(n_slots, 148, 8, dtype=torch.int64, device=dev)
    lib.b200moe_debug_timeline(tl.data_ptr(), n_slots)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            step()
    kinds = [lib.b200moe_debug_timeline_kind(i) for i in range(n_slots)]
    lib.b200moe_debug_timeline(None, 0)
    for _ in range(5):
        dist.barrier()
        graph.replay()
    torch.cuda.synchronize()
    tl.zero_()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    if rank == 0:
        t = tl.cpu().numpy().astype(np.float64)
        t[:, :, 6:] = 0
        base = t[t > 0].min()
        print(f"S={S}/rank layers={L} world={world}: one replay {e0.elapsed_time(e1) * 1e3:.1f} us ({e0.elapsed_time(e1) * 1e3 / L:.2f} us / layer)")
        print(f"{'launch':>6} {'kind':>6} {'ctas':>5} | " + " | ".join(f"{m:>22}" for m in MARKS))
        for i, k in enumerate(kinds):
            if k == 0:
                continue
            a = t[i]
            used = a[:, 0] > 0
            cols = []
            for m in range(6):
                v = a[used, m]
                v = v[v > 0]
                cols.append(f"{'-':>7}{'-':>7}{'-':>8}" if v.size == 0 else
                            f"{(v.min() - base) / 1e3:7.2f}{(np.median(v) - base) / 1e3:7.2f}{(v.max() - base) / 1e3:8.2f}")
            print(f"{i:>6} {('route' if k == 1 else 'ffn'):>6} {int(used.sum()):>5} | " + " | ".join(cols))
    print(f"rank {rank} status {ctx.status()}", flush=True)
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
