#!/usr/bin/env python
"""Per-phase statistics of an FFN timeline produced by tools/ffn_trace.py (gpurun_out/ffn_trace_<S>.npy)."""
import sys
import numpy as np

path = sys.argv[1]
m1, m2, nsm = 8, 4, 148
rec = np.load(path)
t = (rec[..., 2] & 0xFFFFFFFF) | (rec[..., 3] << 32)
valid = t > 0
t0 = t[valid].min()
ev = {}
for c in range(rec.shape[0]):
    for r in range(3):
        for i in range(rec.shape[2]):
            if t[c, r, i] > 0:
                ev.setdefault(int(rec[c, r, i, 0]), {})[int(rec[c, r, i, 1])] = (t[c, r, i] - t0) / 1e3


def stat(name, tiles, a, b):
    d = [ev[x][b] - ev[x][a] for x in tiles if a in ev[x] and b in ev[x]]
    if d:
        print(f"  {name:34s} n={len(d):4d} min={min(d):6.2f} med={np.median(d):6.2f} max={max(d):6.2f}")


tiles = sorted(x for x in ev if x >= 0)
n_tiles = max(tiles) + 1
ng = n_tiles // (m1 + m2)
n_g1 = ng * m1
print(f"span {(t[valid].max() - t0) / 1e3:.2f} us, {n_tiles} tiles, {ng} groups")
sets = [("G1 first wave", [x for x in tiles if x < min(nsm, n_g1)]),
        ("G1 later", [x for x in tiles if nsm <= x < n_g1]),
        ("G2", [x for x in tiles if x >= n_g1])]
for nm, tl in sets:
    if not tl:
        continue
    print(nm)
    stat("dep wait (start->dep_ok)", tl, 1, 2)
    stat("TMA latency (dep_ok->first data)", tl, 2, 5)
    stat("first data->mma issued", tl, 5, 6)
    stat("epi acc_ready->released", tl, 7, 8)
    stat("epi released->stored", tl, 8, 9)
    stat("publish", tl, 9, 10)
    rd = [ev[x][7] for x in tl if 7 in ev[x]]
    st = [ev[x][9] for x in tl if 9 in ev[x]]
    if rd and st:
        print(f"  acc_ready median {np.median(rd):.2f}, last stored {max(st):.2f}")
