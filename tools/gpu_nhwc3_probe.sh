#!/bin/bash
# streaming NHWC 3x3: stage-skipping timings (KDCC_TC_DEBUG bits: 1 no FMAs, 2 no TMA ring, 4 no stores) + one ncu --set full capture
mkdir -p gpurun_out
for dbg in 0 1 2 4 6; do echo "== KDCC_TC_DEBUG=$dbg"; KDCC_TC_DEBUG=$dbg timeout -s KILL 120 python tools/time_dw_nhwc.py 2>&1 | tail -3; done
timeout -s KILL 120 python tools/time_dw_nhwc.py > /dev/null 2>&1 &&
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:dw_nhwc3 -s 4 -c 1 -o gpurun_out/prof_nhwc3 -f python tools/time_dw_nhwc.py > gpurun_out/ncu_n3.out 2>&1
echo "rc=$?"; ls -la gpurun_out/prof_nhwc3.ncu-rep
