#!/bin/bash
# streaming source-major backward: ring depth x CTAs per SM (FSW_RANKT_STREAM: 1 = ring of 6 pairs x 5 CTAs (default), 2 = 8 x 4,
# 0 = the register-staged predecessor) and source rows per warp (FSW_RANKT_RPW).  bwd_sweep.log holds the sweep of the first
# version, which also had 5 x 6 CTAs (18.28 ms) and 12 x 3 CTAs (20.66 ms).
for m in "1 16" "2 16" "0 16" "1 8" "1 32"; do set -- $m; FSW_RANKT_STREAM=$1 FSW_RANKT_RPW=$2 python profiles/r2/layer_times.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('mode $1 rpw $2', d['kernels']['bwd_rankT_u32768_f32'])"; done
