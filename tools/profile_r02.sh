#!/bin/bash
# Launch lists (device time per launch of OUR kernels + the cub sort/select) of two steps of a
# workload, and optional `--set full` captures.  Run under gpurun; outputs in gpurun_out/.
#   tools/profile_r02.sh launches c2 c4 ...      tools/profile_r02.sh full <kernel regex> <workload> <out name>
cd "$(dirname "$0")/.."
B="python bench.py --sub none --no-e2e --no-cpu --no-check --steps 2 --warmup 1"
OURS='regex:qgram|filter_kernel|verify|refine|scan|trace|post_small|minima|push_kernel|collect_kernel|unpack|texts_kernel|overhang|best|DeviceRadixSort|DeviceSelect|DeviceCompact|DeviceScan|suffix'
mode=$1; shift
if [ "$mode" = launches ]; then
  for w in "$@"; do
    ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -c 400 --csv --log-file gpurun_out/r02_launches_$w.csv $B --workload $w > gpurun_out/r02_launches_$w.log 2>&1
  done
else
  ncu --set full --clock-control none --import-source on -k "regex:$1" -s 1 -c 2 -o gpurun_out/$3 $B --workload $2 > gpurun_out/$3.log 2>&1
fi
