#!/bin/bash
for pr in 1 2; do
  echo "=== B200MOE_PAIR=$pr"
  B200MOE_PAIR=$pr python tools/timeline.py 3200 5 2>&1 | sed -n "/expert kernel, us/,\$p" | tail -n 4
done
echo "=== KPS=1"
B200MOE_KPS=1 python tools/timeline.py 3200 5 2>&1 | sed -n "/expert kernel, us/,\$p" | tail -n 2
