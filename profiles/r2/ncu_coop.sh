#!/bin/bash
# ncu --set full capture (with per-instruction source page) of three coop forward classes on a 0.3-scale configs[3] graph
set -e
mkdir -p gpurun_out
TAG=${1:-coop}
python profiles/prof_conv.py 0.3 2 > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:'fsw_coop_fwd_kernel<.int.12, .int.8,|fsw_coop_fwd_kernel<.int.16, .int.4,|fsw_coop_fwd_kernel<.int.12, .int.16,' \
    -s 3 -c 3 -o /tmp/${TAG} -f python profiles/prof_conv.py 0.3 2 > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i /tmp/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv
ncu -i /tmp/${TAG}.ncu-rep --page source --csv -k regex:'int.12, .int.8,' > gpurun_out/${TAG}_src_12_8.csv 2>/dev/null || true
python profiles/ncu_summary.py gpurun_out/${TAG}_raw.csv > gpurun_out/${TAG}_summary.txt
tail -2 gpurun_out/${TAG}_ncu.log
