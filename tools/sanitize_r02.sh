#!/bin/bash
# compute-sanitizer over the round-2 kernels (run under gpurun, 1 GPU).  Output: gpurun_out/r02_sanitizer_*.log
cd "$(dirname "$0")/.."
CS="compute-sanitizer --error-exitcode 99"
$CS --tool memcheck python __graft_entry__.py smoke > gpurun_out/r02_sanitizer_smoke_memcheck.log 2>&1; echo "smoke memcheck rc=$?"
$CS --tool racecheck python __graft_entry__.py smoke > gpurun_out/r02_sanitizer_smoke_racecheck.log 2>&1; echo "smoke racecheck rc=$?"
# q-gram prefilter (both tilings), refinement, warp-per-match traceback from 3 words, regional fallback, scan2, sharded search
$CS --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q -k "qgram_prefilter_fuzz or capacity or kat_props" > gpurun_out/r02_sanitizer_qgram_memcheck.log 2>&1; echo "qgram memcheck rc=$?"
$CS --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_gather.py -x -q -k "v2_fuzz or prefilter_routes or overflow or world1" > gpurun_out/r02_sanitizer_scan2_memcheck.log 2>&1; echo "scan2/regional/gather memcheck rc=$?"
$CS --tool racecheck python -m pytest tests/test_gpu_parity.py -x -q -k "qgram_large_text and contiguous" > gpurun_out/r02_sanitizer_qgram_racecheck.log 2>&1; echo "qgram racecheck rc=$?"
for f in gpurun_out/r02_sanitizer_*.log; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $f | tail -3; done
