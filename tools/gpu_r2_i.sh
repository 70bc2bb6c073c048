#!/bin/bash
mkdir -p gpurun_out
B200MOE_FFN_PREFETCH=0 timeout 200 python tools/timeline.py 3200 6 > gpurun_out/timeline_pf0.txt 2>&1; cat gpurun_out/timeline_pf0.txt
B200MOE_FFN_PREFETCH=1 B200MOE_PDL_TRIG=5 timeout 200 python tools/timeline.py 3200 6 > gpurun_out/timeline_pf1_trig5.txt 2>&1; cat gpurun_out/timeline_pf1_trig5.txt
