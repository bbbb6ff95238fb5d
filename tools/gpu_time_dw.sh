#!/bin/bash
# per-kernel depthwise timings, plain and with stages skipped (KDCC_TC_DEBUG bits)
for dbg in ${DBGS:-0 1 2 3}; do echo "== KDCC_TC_DEBUG=$dbg"; KDCC_TC_DEBUG=$dbg python tools/time_dw.py 2>&1 | tail -${TAILN:-3}; done
