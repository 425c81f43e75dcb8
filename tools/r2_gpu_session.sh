#!/bin/bash
# One single-GPU session of round 2 (every step under its own timeout):
#   gpurun --timeout 1500 -- 'bash tools/r2_gpu_session.sh <tag> [tests] [bench] [ncu]'
set -u
tag=${1:-r2}; shift
mkdir -p gpurun_out
for what in "$@"; do
  case $what in
    tests)
      timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/${tag}_tests.log 2>&1
      echo "tests rc=$?"; tail -4 gpurun_out/${tag}_tests.log ;;
    newtests)
      timeout 600 python -m pytest tests/test_round2_paths.py tests/test_grad_handover.py tests/test_kernel_jitter.py tests/test_fused_step.py -q -m gpu > gpurun_out/${tag}_newtests.log 2>&1
      echo "newtests rc=$?"; tail -15 gpurun_out/${tag}_newtests.log ;;
    each)
      for t in test_render_and_backward_capture_into_a_cuda_graph test_retuning_between_forward_and_backward_is_refused test_packed_volume_of_a_live_graph_is_not_overwritten test_fused_adam_skips_without_gradients_and_leaves_frozen_tensors_alone; do
        timeout 300 python -m pytest tests/test_round2_paths.py -q -m gpu -k $t > gpurun_out/${tag}_$t.log 2>&1
        echo "$t rc=$?"; grep -n "^E \|Error\|passed\|failed" gpurun_out/${tag}_$t.log | head -12
      done ;;
    stages)
      timeout 200 python tools/e2e_stages.py > gpurun_out/${tag}_stages.txt 2>&1; timeout 200 python tools/e2e_stages.py nothreads >> gpurun_out/${tag}_stages.txt 2>&1
      cat gpurun_out/${tag}_stages.txt ;;
    cfg5sweep2)
      timeout 600 python bench.py --workload cfg5 --lanes 1 --sweep "64,16,128;64,32,128;32,16,128;32,32,128;16,16,128;48,16,128;64,24,128" 2>&1 | grep sweep
      echo cfg3; timeout 600 python bench.py --workload cfg3 --lanes 1 --sweep "16,8,128;16,16,128;16,32,128;32,16,128;32,32,128;64,32,128" 2>&1 | grep sweep
      echo cfg4; timeout 600 python bench.py --workload cfg4 --lanes 1 --sweep "16,8,80;16,8,96;32,16,64;16,16,64;16,32,64;32,32,64;16,8,128;16,16,128;32,16,128" 2>&1 | grep sweep ;;
    cfg5sweep)
      timeout 600 python bench.py --workload cfg5 --lanes 1 --sweep "16,4,128;32,8,128;64,16,128;32,4,128;64,8,128;16,8,128" 2>&1 | grep sweep ;;
    rpc)
      timeout 300 python bench.py --lanes 1 --sweep "16,8,96;16,7,96;16,6,96;16,5,96;16,4,96;15,7,96;14,7,96;16,7,80;16,7,128" 2>&1 | grep sweep
      echo "3 lanes"; timeout 300 python bench.py --lanes 3 --sweep "16,8,96;16,7,96;16,4,96" 2>&1 | grep sweep ;;
    rot)
      for r in 0 25 13 37 74; do
        echo "rotation $r"
        VOXE_GROUP_ROTATION=$r timeout 200 python bench.py --lanes 1 --sweep "16,8,96;16,4,96;16,4,128;8,4,96" 2>&1 | grep sweep
      done
      echo "rotation 25, 3 lanes"; VOXE_GROUP_ROTATION=25 timeout 200 python bench.py --lanes 3 --sweep "16,8,96" 2>&1 | grep sweep
      echo "rotation 0, 3 lanes"; VOXE_GROUP_ROTATION=0 timeout 200 python bench.py --lanes 3 --sweep "16,8,96" 2>&1 | grep sweep ;;
    bench)
      timeout 500 python bench.py --steps 60 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
      echo "bench rc=$?"; tail -5 gpurun_out/${tag}_bench.err
      timeout 60 python tools/show_bench.py gpurun_out/${tag}_bench.json $tag < /dev/null ;;
    sweep1)
      timeout 400 python bench.py --lanes 1 --sweep "16,8,96;16,8,128;16,8,64;16,16,64;8,4,96;8,4,80;8,4,64;8,8,64;8,16,64;8,8,128;4,2,96;4,4,64;4,8,64;4,8,128;32,8,96;32,16,64" > gpurun_out/${tag}_sweep1.txt 2>&1
      echo "sweep1 rc=$?"; cat gpurun_out/${tag}_sweep1.txt ;;
    ncu)
      timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches_${tag}.csv python bench.py --ncu --steps 2 > /dev/null 2>&1
      echo "ncu launches rc=$?"
      for k in bwd fwd; do
        timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:render_$k -s 30 -c 2 \
          -f -o gpurun_out/prof_${k}_${tag} python bench.py --ncu --steps 1 > gpurun_out/${tag}_ncu_$k.log 2>&1
        echo "ncu $k rc=$?"
      done ;;
  esac
done
