#!/bin/bash
# Tiling / residency / ring-depth experiments of the prefilter kernels on one GPU (bench --no-cpu --no-e2e, 40 steps).
# usage: tools/tile_experiments.sh > profiles/<name>.txt
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
for w in c4 c2; do
  run "$w default" X=1 -- --workload $w
  for rb in 512 1024 2048 4096 8192; do run "$w row_bytes=$rb" SASSY_B200_FILTER_ROW_BYTES=$rb -- --workload $w; done
done
for bps in 4 5 7 8; do run "c4 qgram_bps=$bps" SASSY_B200_QGRAM_BPS=$bps -- --workload c4; done
for v in st3 st4 sb128; do
  lib=$PWD/sassy_b200/lib/libsassy_b200_$v.so
  [ -f $lib ] || continue
  for w in c4 c2; do
    run "$w lib=$v" SASSY_B200_LIB=$lib -- --workload $w
    run "$w lib=$v row_bytes=2048" SASSY_B200_LIB=$lib SASSY_B200_FILTER_ROW_BYTES=2048 -- --workload $w
  done
  for bps in 4 8; do run "c4 lib=$v qgram_bps=$bps" SASSY_B200_LIB=$lib SASSY_B200_QGRAM_BPS=$bps -- --workload c4; done
done
