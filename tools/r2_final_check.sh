#!/bin/bash
# What the driver runs at round end, in one single-GPU session: GPU tests, smoke(), the reference arm, the bench line.
set -u
mkdir -p gpurun_out
tag=${1:-final}
t0=$(date +%s)
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$? ($(( $(date +%s) - t0 )) s)"; tail -2 gpurun_out/${tag}_tests.log
t0=$(date +%s)
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$? ($(( $(date +%s) - t0 )) s)"; tail -1 gpurun_out/${tag}_smoke.log
t0=$(date +%s)
timeout 600 python bench.py --impl reference --gpus 1 --steps 200 --warmup 10 > gpurun_out/${tag}_reference.json 2> gpurun_out/${tag}_reference.err; echo "reference rc=$? ($(( $(date +%s) - t0 )) s)"; cut -c1-300 gpurun_out/${tag}_reference.json
t0=$(date +%s)
timeout 900 python bench.py --gpus 1 --steps 200 --warmup 10 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$? ($(( $(date +%s) - t0 )) s)"; tail -3 gpurun_out/${tag}_bench.err
timeout 60 python tools/show_bench.py gpurun_out/${tag}_bench.json $tag < /dev/null
