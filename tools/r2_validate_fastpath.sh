#!/bin/bash
# First GPU call of round 2 for this branch (every step under its own timeout -- a helper that waited on stdin once ate a
# whole call):
#   gpurun --timeout 900 -- 'bash tools/r2_validate_fastpath.sh'
# 1. parity of the specialised kernel variants (A/B against the generic kernels + the parity suites under the switch)
# 2. bench line with the generic kernels, then with VOXE_SPECIALISED_KERNELS=1 (same box, back to back)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_specialised_kernels.py -q -m gpu -x > gpurun_out/r2_fast_tests.log 2>&1
echo "specialised tests rc=$?"; tail -3 gpurun_out/r2_fast_tests.log
timeout 300 python -m pytest tests -q -m gpu -x > gpurun_out/r2_all_tests.log 2>&1
echo "whole suite rc=$?"; tail -2 gpurun_out/r2_all_tests.log
timeout 200 python bench.py --no-cpu > gpurun_out/r2_bench_generic.json 2> gpurun_out/r2_bench_generic.err
timeout 60 python tools/show_bench.py gpurun_out/r2_bench_generic.json generic < /dev/null
VOXE_SPECIALISED_KERNELS=1 timeout 200 python bench.py --no-cpu > gpurun_out/r2_bench_fast.json 2> gpurun_out/r2_bench_fast.err
timeout 60 python tools/show_bench.py gpurun_out/r2_bench_fast.json specialised < /dev/null
