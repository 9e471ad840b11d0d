#!/bin/bash
# Ring depth / stage size experiments of the piece-automaton prefilter (c2) on one GPU: variant libraries
# built with -DSB_STAGES / -DSB_STAGE_BYTES (sassy_b200/build.py build_variant).  usage: tools/ring_experiments.sh > profiles/<name>.txt
cd "$(dirname "$0")/.."
run() {  # label, env..., -- bench args
  label=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  out=$(env "${envs[@]}" python bench.py --no-cpu --no-e2e --no-check --sub none --steps 40 --warmup 5 "$@" 2>/dev/null | tail -1)
  python - "$label" "$out" <<'PY'
import json, sys
try:
    j = json.loads(sys.argv[2]); r = j["roofline"]
    print(f"{sys.argv[1]:40s} ms/step {j['ms_per_step']:.4f}  kernel {r['kernel']} {r['kernel_ms']:.4f} ms  frac {r['frac']:.3f}  verify {r['verify_kernel_ms']:.4f}  bps {j['config']['blocks_per_sm']} rows {j['config']['rows']} row_bytes {j['config']['row_bytes']}", flush=True)
except Exception as e:
    print(sys.argv[1], "FAILED", e, sys.argv[2][:200], flush=True)
PY
}
run "c2 default" X=1 -- --workload c2
for v in st3 st4 sb128 sb128st3; do
  lib=$PWD/sassy_b200/lib/libsassy_b200_$v.so
  [ -f $lib ] || continue
  for rb in 0 2048 8192; do
    run "c2 lib=$v row_bytes=$rb" SASSY_B200_LIB=$lib SASSY_B200_FILTER_ROW_BYTES=$rb -- --workload c2
  done
done
run "c2 default again" X=1 -- --workload c2
