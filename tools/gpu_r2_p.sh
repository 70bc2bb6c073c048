#!/bin/bash
mkdir -p gpurun_out
echo skip-pytest

timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --encoder > gpurun_out/bench_cfg3_encoder.log 2>&1; echo "bench exit=$?"
tail -c 1500 gpurun_out/bench_cfg3_encoder.log | grep -v "^{" | tail -5
python - <<'PY'
import json
for f in ["gpurun_out/bench_cfg3_encoder.log"]:
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l); print(f, "us/layer", round(d["us_per_layer"], 2), json.dumps(d.get("encoder"))[:900])
PY
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --encoder --workload cfg2 > gpurun_out/bench_cfg2_encoder.log 2>&1; echo "bench exit=$?"
python - <<'PY'
import json
for f in ["gpurun_out/bench_cfg2_encoder.log"]:
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l); print(f, "us/layer", round(d["us_per_layer"], 2), json.dumps(d.get("encoder"))[:900])
PY
