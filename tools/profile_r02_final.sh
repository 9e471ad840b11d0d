#!/bin/bash
# Round-2 profile set (run under gpurun, 1 GPU): launch lists of c2 / c4 and `--set full` captures of the
# three dominant kernels.  Summaries are extracted into profiles/ by tools/ncu_summary.py.
cd "$(dirname "$0")/.."
tools/profile_r02.sh launches c2 c4
B="python bench.py --sub none --no-e2e --no-cpu --no-check --steps 2 --warmup 1"
ncu --set full --clock-control none --import-source on -k regex:qgram_seq_kernel -s 1 -c 1 -o gpurun_out/r02_qgram_c4 $B --workload c4 > gpurun_out/r02_qgram_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:filter_kernel -s 1 -c 1 -o gpurun_out/r02_filter_c2 $B --workload c2 > gpurun_out/r02_filter_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan2_kernel -c 1 -o gpurun_out/r02_scan2_c3p64 $B --workload c3 --patterns 64 --steps 1 > gpurun_out/r02_scan2_c3p64.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:scan2|scan_kernel|trace|post_small|minima|DeviceRadixSort|DeviceSelect|DeviceCompact" -c 60 --csv --log-file gpurun_out/r02_launches_c3p64.csv $B --workload c3 --patterns 64 --steps 1 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
