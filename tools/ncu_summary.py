#!/usr/bin/env python
"""Summarises ncu outputs brought back in gpurun_out/: launch list (share per kernel) and top stall sites."""
import collections
import csv
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr = rows[h]
    kn, mv, mu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    d = collections.defaultdict(list)
    for r in rows[h + 1:]:
        if len(r) <= mv:
            continue
        name = r[kn].split('(')[0][-58:]
        v = float(r[mv].replace(',', ''))
        v = v / 1000 if r[mu] == 'ns' else (v * 1000 if r[mu] == 'ms' else v)
        d[name].append(v)
    tot = sum(sum(v) for v in d.values())
    print(f"{'kernel':58s} {'n':>5s} {'avg us':>9s} {'share':>6s}")
    for k, v in sorted(d.items(), key=lambda t: -sum(t[1])):
        print(f"{k:58s} {len(v):5d} {sum(v)/len(v):9.2f} {100*sum(v)/tot:5.1f}%")


def raw(rep, keys):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(r[hdr.index('Kernel Name')][:80])
        for k in keys:
            if k in hdr:
                i = hdr.index(k)
                print(f"   {k:70s} {r[i]:>14s} {units[i]}")


def stalls(rep, kernel_regex, top=25):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-id', f'::regex:{kernel_regex}:1'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    si, src, ex = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
    cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    data = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            break
        try:
            data.append((int(r[si]), r))
        except ValueError:
            pass
    tot = sum(n for n, _ in data) or 1
    print("total samples", tot)
    for n, r in sorted(data, key=lambda t: -t[0])[:top]:
        st = sorted(((hdr[i][6:], int(r[i])) for i in cols if r[i] not in ('', '0')), key=lambda t: -t[1])[:2]
        print(f"{n:6d} {100*n/tot:5.1f}% ex={r[ex]:>8} {r[src].strip()[:64]:64s} {st}")


KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'lts__t_bytes.sum',
        'lts__t_sector_hit_rate.pct', 'sm__cycles_active.avg']

if __name__ == '__main__':
    cmd = sys.argv[1]
    if cmd == 'launches':
        launches(sys.argv[2])
    elif cmd == 'raw':
        raw(sys.argv[2], KEYS)
    elif cmd == 'stalls':
        stalls(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 25)
