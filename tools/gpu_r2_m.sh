#!/bin/bash
for d in 0 256 32 288; do
  echo "=== B200MOE_DBG=$d (single layer)"
  B200MOE_DBG=$d python tools/timeline.py 3200 1 2>&1 | sed -n "/expert kernel, us/,\$p" | tail -n 2
done
