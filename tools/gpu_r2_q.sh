#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoder.py tests/test_encoder_ops.py -m gpu -q -x --timeout 300 > gpurun_out/pytest_encoder.log 2>&1; echo "pytest exit=$?"
tail -n 12 gpurun_out/pytest_encoder.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --encoder > gpurun_out/bench_cfg3_encoder.log 2>&1; echo "bench exit=$?"
python - <<'PY'
import json
for f in ["gpurun_out/bench_cfg3_encoder.log"]:
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l); print(f, "us/layer", round(d["us_per_layer"], 2), json.dumps(d.get("encoder"))[:400])
PY
