#!/usr/bin/env python
"""Runs one fast_moe layer with the FFN kernel's debug tracer on and prints per-CTA timelines (GPU box only).

usage: python tools/ffn_trace.py [S] [n_cta_to_print]     -> text on stdout, raw records in gpurun_out/ffn_trace_S.npy
"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "3m-asr-inference_b200"
EV = ["kernel_start", "prod_tile_start", "prod_dep_ok", "prod_issued", "mma_acc_free", "mma_first_data", "mma_issued",
      "epi_acc_ready", "epi_acc_released", "epi_stored", "epi_published", "kernel_end", "epi_chunk_ld", "epi_chunk_staged",
      "epi_chunk_done", "clock", "cycles", "prod_slot_free(kb)", "mma_stage_data(kb)", "mma_stage_issued(kb)", "mma_instr(i)", "warm_issued(b)", "warm_done(b)"]


def analyze(rec, cap, n_cta, S, n_print):
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.save(os.path.join(ROOT, "gpurun_out", f"ffn_trace_{S}.npy"), rec)
    raw = (rec[..., 2] & 0xFFFFFFFF) | (rec[..., 3] << 32)     # clock64 for events, ns for sync records
    evc = rec[..., 1]
    t = np.zeros(raw.shape, dtype=np.float64)                  # ns on the common time base; 0 = no record
    rates = []
    gbase = int(min(raw[c, 0, i] for c in range(n_cta) for i in range(cap // 4) if evc[c, 0, i] == 15 and raw[c, 0, i] > 0))
    for c in range(n_cta):
        # producer role: (globaltimer, clock64) pairs at kernel start and end
        g = [raw[c, 0, i] for i in range(cap // 4) if evc[c, 0, i] == 15 and raw[c, 0, i] > 0]
        k = [raw[c, 0, i] for i in range(cap // 4) if evc[c, 0, i] == 16 and raw[c, 0, i] > 0]
        if len(g) < 2 or len(k) < 2 or g[-1] <= g[0]:
            continue
        rate = (k[-1] - k[0]) / (g[-1] - g[0])                 # cycles per ns
        rates.append(rate * 1e3)
        m = (raw[c] > 0) & ((evc[c] < 15) | (evc[c] > 16))
        t[c][m] = 1.0 + float(g[0] - gbase) + (raw[c][m] - k[0]).astype(np.float64) / rate
    valid = t > 0
    t0 = t[valid].min()
    print(f"S={S}: kernel span {(t[valid].max() - t0) / 1e3:.2f} us over {int(valid.sum())} records")
    if rates:
        print(f"  SM clock during the kernel: median {np.median(rates):.0f} MHz (min {min(rates):.0f}, max {max(rates):.0f})")
    # per-event statistics relative to kernel start
    for ev in list(range(15)) + [17, 18, 19, 20, 21, 22]:
        m = valid & (rec[..., 1] == ev)
        if m.any():
            tt = (t[m] - t0) / 1e3
            print(f"  {EV[ev]:18s} n={int(m.sum()):5d} first={tt.min():8.2f} median={np.median(tt):8.2f} last={tt.max():8.2f} us")
    ctas = list(range(0, n_cta, max(1, n_cta // n_print)))[:n_print]
    for c in ctas:
        rows = []
        for role in range(4):
            for i in range(cap // 4):
                if t[c, role, i] > 0:
                    rows.append(((t[c, role, i] - t0) / 1e3, role, int(rec[c, role, i, 0]), int(rec[c, role, i, 1])))
        rows.sort()
        print(f"--- CTA {c}")
        for ts, role, tile, ev in rows:
            print(f"   {ts:8.2f} us  {'PMEU'[role]}  tile {tile:4d}  {EV[ev]}")



def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 3200
    n_print = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    ops = importlib.import_module(PKG + ".ops")
    lib = importlib.import_module(PKG + "._lib").load()
    E, D, H, Demb = 32, 512, 1024, 512
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(1)
    layers = []
    for _ in range(4):  # cycle 4 weight sets so that weights are not L2 resident
        W1 = ((torch.rand(E, H, D, generator=g, device=dev) * 2 - 1) * 0.05).bfloat16()
        W2 = ((torch.rand(E, D, H, generator=g, device=dev) * 2 - 1) * 0.05).bfloat16()
        Wr = ((torch.rand(Demb + D, E, generator=g, device=dev) * 2 - 1) * 0.04)
        layers.append((Wr, ops.PackedExperts(W1, torch.zeros(E, H, device=dev), W2, torch.zeros(E, D, device=dev)), ops.pack_router(Wr)))
    x = torch.randn(S, D, generator=g, device=dev).bfloat16()
    emb = torch.randn(S, Demb, generator=g, device=dev).bfloat16()
    out = torch.empty_like(x)
    for _ in range(3):
        for Wr, ex, wp in layers:
            ops.moe_layer(x, emb, Wr, None, ex, residual=x, ff_scale=0.5, out=out, Wr_packed=wp)
    torch.cuda.synchronize()
    cap = 256  # records per CTA (4 roles x 64)
    n_cta = 148
    buf = torch.zeros(n_cta * cap, 4, dtype=torch.int32, device=dev)
    lib.b200moe_debug_ffn_trace(buf.data_ptr(), cap)
    Wr, ex, wp = layers[0]
    ops.moe_layer(x, emb, Wr, None, ex, residual=x, ff_scale=0.5, out=out, Wr_packed=wp)
    torch.cuda.synchronize()
    lib.b200moe_debug_ffn_trace(None, 0)
    rec = buf.cpu().numpy().astype(np.int64).reshape(n_cta, 4, cap // 4, 4)
    analyze(rec, cap, n_cta, S, n_print)


if __name__ == "__main__":
    main()
