#!/bin/bash
# compute-sanitizer passes over the round-2 kernels (single GPU; every step under a timeout)
set -u
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $S --tool memcheck --error-exitcode 0 python -m pytest tests/test_round2_paths.py tests/test_resample.py tests/test_grad_handover.py -q -m gpu -k "touched or kernel_matches or handover or direct or frozen or retuning or fused_adam" > gpurun_out/r2_san_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds" gpurun_out/r2_san_memcheck.log | head -10
timeout 900 $S --tool memcheck --error-exitcode 0 python -m pytest tests/test_random_sweep.py -q -m gpu -k "0 or 4 or 10 or 21 or 33 or 35 or 45" > gpurun_out/r2_san_memcheck_sweep.log 2>&1
echo "memcheck sweep rc=$?"; grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds" gpurun_out/r2_san_memcheck_sweep.log | head -10
timeout 600 $S --tool racecheck --error-exitcode 0 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_san_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|smoke" gpurun_out/r2_san_racecheck.log | head
timeout 600 $S --tool initcheck --error-exitcode 0 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_san_initcheck.log 2>&1
echo "initcheck rc=$?"; grep -E "ERROR SUMMARY|smoke|Uninitialized" gpurun_out/r2_san_initcheck.log | head
timeout 900 $S --tool memcheck --error-exitcode 0 python -m pytest tests/test_point_query.py tests/test_regularizers.py tests/test_fused_step.py tests/test_round2_paths.py -q -m gpu -k "point or query or edge or tv or trail or touched or kernels_match" > gpurun_out/r2_san_memcheck_new.log 2>&1
echo "memcheck (point query, TV strips, brick trail, tiled hand-over) rc=$?"; grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds" gpurun_out/r2_san_memcheck_new.log | head -10
