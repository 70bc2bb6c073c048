#!/usr/bin/env python
"""One line per bench JSON line found in the given log files."""
import json
import sys

for f in sys.argv[1:]:
    try:
        lines = [l for l in open(f) if l.startswith("{")]
    except OSError:
        continue
    if not lines:
        print(f, "NO JSON LINE")
    for l in lines:
        d = json.loads(l)
        if "error" in d or "layer_roofline" not in d:
            print(f.split("/")[-1], {k: d[k] for k in ("impl", "metric", "value", "unit", "ms_per_step", "error") if k in d},
                  (d.get("cpu_baseline") or {}).get("sample", "")[:120])
            continue
        rf = d.get("roofline") or {}
        st = {k: (round(v, 1) if v else v) for k, v in (d.get("stage_us_per_layer") or {}).items()}
        rs = {k[:8]: round(v["frac_of_hbm_peak"], 2) for k, v in (d.get("roofline_stages") or {}).items()}
        sp = d.get("sweep_point")
        tag = d["config"]["workload"][:5] if not sp else f"S={sp['global_tokens']} k={sp['top_k']}{' zipf' if sp['zipf'] else ''}"
        print(f.split("/")[-1], tag, "n", d["n_gpus"], "us/layer", round(d.get("us_per_layer", 0), 2), "stages", st,
              "ffn", rf.get("bound"), round(rf.get("frac") or 0, 3), "in-graph", round(rf.get("frac_in_graph") or 0, 3),
              "span", round(rf.get("in_graph_span_us") or 0, 2), "stage_frac", rs,
              "layer_hbm", round(d["layer_roofline"]["hbm_frac"], 3), "layer_tc", round(d["layer_roofline"]["tensor_frac"], 3),
              "tok/s", f"{d['value']:.3e}", "e2e", f"{d['e2e']['value']:.3e}" if d.get("e2e") else None,
              "parity", d.get("parity") and (d["parity"]["ok"], round(d["parity"]["rel_l2"], 5)),
              "sus", d.get("sustained") and f"{d['sustained']['value']:.3e}", "clk", d.get("clocks", {}).get("sm_mhz"),
              d.get("clocks", {}).get("reasons"))
