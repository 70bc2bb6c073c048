#!/usr/bin/env python
"""Timeline of the fused gate + dispatch kernel (GPU box only): python tools/route_trace.py [S] [ln]
(`ln`: with the block's norm_ff fused into the kernel, b200moe_config("ln_fuse", 1))"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "3m-asr-inference_b200"
EV = ["start", "setup_done", "tma_issued", "acc_ready", "idx_written", "barrier_arrive", "barrier_passed",
      "offsets_done", "ranks_done", "copies_done", "end"]


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 3200
    ln = len(sys.argv) > 2 and sys.argv[2] == "ln"
    ops = importlib.import_module(PKG + ".ops")
    lib = importlib.import_module(PKG + "._lib").load()
    E, D, H, Demb = 32, 512, 1024, 512
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(1)
    layers = []
    for _ in range(4):
        W1 = ((torch.rand(E, H, D, generator=g, device=dev) * 2 - 1) * 0.05).bfloat16()
        W2 = ((torch.rand(E, D, H, generator=g, device=dev) * 2 - 1) * 0.05).bfloat16()
        Wr = ((torch.rand(Demb + D, E, generator=g, device=dev) * 2 - 1) * 0.04)
        layers.append((Wr, ops.PackedExperts(W1, torch.zeros(E, H, device=dev), W2, torch.zeros(E, D, device=dev)),
                       ops.pack_router(Wr)))
    x = torch.randn(S, D, generator=g, device=dev).bfloat16()
    emb = torch.randn(S, Demb, generator=g, device=dev).bfloat16()
    out = torch.empty_like(x)
    kw = {}
    if ln:
        kw = {"norm_ff": (torch.ones(D, device=dev) * 1.1, torch.zeros(D, device=dev) + 0.05)}
        layers = [(Wr, ex, wp, ops.pack_router_ln(Wr, *kw["norm_ff"])) for Wr, ex, wp in layers]
    else:
        layers = [(Wr, ex, wp, None) for Wr, ex, wp in layers]
    for _ in range(3):
        for Wr, ex, wp, wln in layers:
            ops.moe_layer(x, emb, Wr, None, ex, residual=x, ff_scale=0.5, out=out, Wr_packed=wp, Wr_packed_ln=wln, **kw)
    torch.cuda.synchronize()
    buf = torch.zeros(148 * 16, 4, dtype=torch.int32, device=dev)
    lib.b200moe_debug_route_trace(buf.data_ptr())
    Wr, ex, wp, wln = layers[1]
    ops.moe_layer(x, emb, Wr, None, ex, residual=x, ff_scale=0.5, out=out, Wr_packed=wp, Wr_packed_ln=wln, **kw)
    torch.cuda.synchronize()
    lib.b200moe_debug_route_trace(None)
    analyze(buf.cpu().numpy().astype(np.int64).reshape(148, 16, 4), S)


def analyze(rec, S):
    val = (rec[..., 1] & 0xFFFFFFFF) | ((rec[..., 2] & 0xFFFFFFFF) << 32)
    ok = rec[..., 3] == 1
    ctas = [c for c in range(148) if ok[c, 14] and ok[c, 15] and ok[c, 0] and ok[c, 10]]
    g0 = min(val[c, 14] for c in ctas)
    t = np.full((148, 11), np.nan)
    for c in ctas:
        rate = (val[c, 10] - val[c, 0]) / max(val[c, 15] - val[c, 14], 1)   # cycles per ns
        for e in range(11):
            if ok[c, e]:
                t[c, e] = (val[c, 14] - g0) + (val[c, e] - val[c, 0]) / rate
    print(f"S={S}: {len(ctas)} CTAs, kernel span {np.nanmax(t) / 1e3:.2f} us")
    for e in range(11):
        col = t[:, e][~np.isnan(t[:, e])] / 1e3
        if col.size:
            print(f"  {EV[e]:16s} n={col.size:4d} first={col.min():7.2f} median={np.median(col):7.2f} last={col.max():7.2f} us")


if __name__ == "__main__":
    main()
