#!/bin/bash
# source-level (per-instruction) ncu capture of two forward classes on a 0.3-scale configs[3] graph:
#   gpurun_out/r2_src_small16.csv, r2_src_coop96.csv
set -e
mkdir -p gpurun_out
python profiles/prof_conv.py 0.3 2 > gpurun_out/r2_fwd_plain.log 2>&1 &&
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:'fsw_small_fwd_kernel<float, .int.16|fsw_coop_fwd_kernel<.int.12, .int.8,' \
    -s 2 -c 2 -o /tmp/r2_fwd -f python profiles/prof_conv.py 0.3 2 > gpurun_out/r2_fwd_ncu.log 2>&1
ncu -i /tmp/r2_fwd.ncu-rep --page raw --csv > gpurun_out/r2_fwd_raw.csv
ncu -i /tmp/r2_fwd.ncu-rep --page source --csv -k regex:fsw_small_fwd > gpurun_out/r2_src_small16.csv 2>/dev/null || true
ncu -i /tmp/r2_fwd.ncu-rep --page source --csv -k regex:fsw_coop_fwd > gpurun_out/r2_src_coop96.csv 2>/dev/null || true
python profiles/ncu_summary.py gpurun_out/r2_fwd_raw.csv > gpurun_out/r2_fwd_summary.txt
ls -la gpurun_out | tail -8; tail -2 gpurun_out/r2_fwd_ncu.log
