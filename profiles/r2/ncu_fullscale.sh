#!/bin/bash
# ncu --set full capture of ONE layer (fwd+bwd) of BASELINE.json configs[3] at FULL size with this binary:
#   every fsw_* kernel of the step once (forward size classes, source-major backward, K1 tensor-core kernels).
# Run on the GPU box:  bash profiles/r2/ncu_fullscale.sh
#   -> gpurun_out/r2_full_raw.csv (all metrics per launch), r2_full_summary.txt, r2_src_<kernel>.csv (per-instruction stalls)
# then here:           python profiles/r2/ncu_traffic.py gpurun_out/r2_full_raw.csv   -> profiles/r2/ncu_traffic.json
# The .ncu-rep itself is deleted on the box (gpurun returns at most 64 MiB).
set -e
mkdir -p gpurun_out
python profiles/prof_conv.py 1.0 3 > gpurun_out/r2_full_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:'fsw_(rank_bwdS|rank_bwdT|coop_fwd|small_fwd|medium|umma|scale_grad)' \
    -s 30 -c 40 -o /tmp/r2_full -f python profiles/prof_conv.py 1.0 3 > gpurun_out/r2_full_ncu.log 2>&1
ncu -i /tmp/r2_full.ncu-rep --page raw --csv > gpurun_out/r2_full_raw.csv
python profiles/ncu_summary.py gpurun_out/r2_full_raw.csv > gpurun_out/r2_full_summary.txt
ncu -i /tmp/r2_full.ncu-rep --page source --csv -k regex:fsw_rank_bwd > gpurun_out/r2_src_rank_bwdS.csv 2>/dev/null || true
ncu -i /tmp/r2_full.ncu-rep --page source --csv -k regex:fsw_umma --launch-count 1 > gpurun_out/r2_src_umma.csv 2>/dev/null || true
ls -la /tmp/r2_full.ncu-rep gpurun_out | tail -12
tail -3 gpurun_out/r2_full_ncu.log
