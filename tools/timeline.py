#!/usr/bin/env python
"""Cross-kernel timeline of a CUDA-graph replay of N fast_moe layers (GPU box only): when the CTAs of every gate+dispatch
and expert-kernel launch start, pass their dependency wait, see their first data, stop requesting weights, issue their
last MMA and exit, all on the %globaltimer base.   usage: python tools/timeline.py [S] [layers] [block]"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "3m-asr-inference_b200"
MARKS = ["start", "wait_ok", "first_data", "last_req", "last_mma", "end"]


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 3200
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    ops = importlib.import_module(PKG + ".ops")
    lib = importlib.import_module(PKG + "._lib").load()
    E, D, H, Demb = 32, 512, 1024, 512
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(1)
    layers = []
    for _ in range(L):
        W1 = ((torch.rand(E, H, D, generator=g, device=dev) * 2 - 1) * 0.05).bfloat16()
        W2 = ((torch.rand(E, D, H, generator=g, device=dev) * 2 - 1) * 0.05).bfloat16()
        Wr = ((torch.rand(Demb + D, E, generator=g, device=dev) * 2 - 1) * 0.04)
        layers.append((Wr, ops.PackedExperts(W1, torch.zeros(E, H, device=dev), W2, torch.zeros(E, D, device=dev)),
                       ops.pack_router(Wr)))
    x0 = torch.randn(S, D, generator=g, device=dev).bfloat16()
    emb = torch.randn(S, Demb, generator=g, device=dev).bfloat16()
    bufs = [torch.empty_like(x0), torch.empty_like(x0)]

    def step():
        cur = x0
        for i, (Wr, ex, wp) in enumerate(layers):
            out = bufs[i & 1]
            ops.moe_layer(cur, emb, Wr, None, ex, residual=cur, ff_scale=0.5, out=out, Wr_packed=wp)
            cur = out
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.synchronize()
    n_slots = 2 * L + 4
    tl = torch.zeros(n_slots, 148, 8, dtype=torch.int64, device=dev)
    lib.b200moe_debug_timeline(tl.data_ptr(), n_slots)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            step()
    kinds = [lib.b200moe_debug_timeline_kind(i) for i in range(n_slots)]
    lib.b200moe_debug_timeline(None, 0)
    for _ in range(5):
        graph.replay()
    torch.cuda.synchronize()
    tl.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    tl_raw = tl.cpu().numpy()
    t = tl_raw.astype(np.float64)
    t[:, :, 6:] = 0
    base = t[t > 0].min()
    print(f"S={S} layers={L}: one replay {e0.elapsed_time(e1) * 1e3:.1f} us ({e0.elapsed_time(e1) * 1e3 / L:.2f} us / layer, marks on)")
    print(f"{'launch':>8} {'kind':>6} {'ctas':>5} | " + " | ".join(f"{m:>22}" for m in MARKS))
    print(" " * 23 + " | ".join(f"{'min':>7}{'med':>7}{'max':>8}" for _ in MARKS))
    prev_end = None
    for i, k in enumerate(kinds):
        if k == 0:
            continue
        a = t[i]
        used = a[:, 0] > 0
        cols = []
        for m in range(len(MARKS)):
            v = a[used, m]
            v = v[v > 0]
            if v.size == 0:
                cols.append(f"{'-':>7}{'-':>7}{'-':>8}")
            else:
                v = (v - base) / 1e3
                cols.append(f"{v.min():7.2f}{np.median(v):7.2f}{v.max():8.2f}")
        print(f"{i:>8} {('route' if k == 1 else 'ffn'):>6} {int(used.sum()):>5} | " + " | ".join(cols))
    # compact: expert-kernel launches relative to the end of the kernel in front of them (max over its CTAs)
    print("\nexpert kernel, us after the LAST CTA of the gate+dispatch kernel in front has ended (median / max over CTAs):")
    names = ["start", "wait_ok", "first_data", "last_req", "last_mma", "end"]
    idx = [0, 1, 2, 3, 4, 5]
    print(f"{'launch':>8} " + " ".join(f"{n:>15}" for n in names) + f" {'route: wait_ok->end':>22}")
    for i, k in enumerate(kinds):
        if k != 2 or i == 0 or kinds[i - 1] != 1:
            continue
        r = t[i - 1]
        ru = r[:, 0] > 0
        ref = r[ru, 5].max()
        a = t[i]
        used = a[:, 0] > 0
        cols = []
        for m in idx:
            v = a[used, m]
            v = (v[v > 0] - ref) / 1e3
            cols.append(f"{np.median(v):7.2f}/{v.max():7.2f}" if v.size else f"{'-':>15}")
        rw = r[ru, 2]
        rw = rw[rw > 0]
        route = f"{(np.median(r[ru, 5]) - np.median(rw)) / 1e3:7.2f} (max end {(ref - np.median(rw)) / 1e3:6.2f})" if rw.size else "-"
        print(f"{i:>8} " + " ".join(cols) + f" {route:>22}")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.save(os.path.join(ROOT, "gpurun_out", f"timeline_{S}.npy"), t)


if __name__ == "__main__":
    main()
