#!/bin/bash
# compute-sanitizer over the kernels added late in round 2 (run under gpurun, 1 GPU): lane-group re-scan,
# shared-memory column store of the warp-per-match traceback, patterns beyond 32 words (cover entries),
# concatenated many-text route, ring transport.  Output: gpurun_out/r02b_sanitizer_*.log
cd "$(dirname "$0")/.."
CS="compute-sanitizer --error-exitcode 99"
$CS --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q -k "long_patterns or beyond_32 or qgram_prefilter_fuzz or kat_props" > gpurun_out/r02b_sanitizer_wide_memcheck.log 2>&1; echo "wide memcheck rc=$?"
$CS --tool racecheck python -m pytest tests/test_gpu_parity.py -x -q -k "test_gpu_long_patterns_many_pieces or qgram_prefilter_fuzz" > gpurun_out/r02b_sanitizer_wide_racecheck.log 2>&1; echo "wide racecheck rc=$?"
$CS --tool memcheck python -m pytest tests/test_gpu_options.py -x -q -k "search_many or search_texts" > gpurun_out/r02b_sanitizer_texts_memcheck.log 2>&1; echo "texts memcheck rc=$?"
$CS --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q -k "test_dna_packed_transport" > gpurun_out/r02b_sanitizer_transport_memcheck.log 2>&1; echo "transport memcheck rc=$?"
$CS --tool memcheck python __graft_entry__.py smoke > gpurun_out/r02b_sanitizer_smoke_memcheck.log 2>&1; echo "smoke memcheck rc=$?"
for f in gpurun_out/r02b_sanitizer_*.log; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $f | tail -3; done
