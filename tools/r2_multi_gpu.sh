#!/bin/bash
# Multi-GPU session (every step under its own timeout):  gpurun --gpus N --timeout 1500 -- 'bash tools/r2_multi_gpu.sh N tag [check] [bench] [nccl] [cfg4] [cfg5]'
set -u
n=$1; tag=$2; shift 2
mkdir -p gpurun_out
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $n "${@:3}"; }
port=29500
for what in "$@"; do
  port=$((port + 7))
  case $what in
    check)
      run 300 $port --check > gpurun_out/${tag}_check_n$n.json 2> gpurun_out/${tag}_check_n$n.err
      echo "check rc=$?"; cat gpurun_out/${tag}_check_n$n.json; tail -5 gpurun_out/${tag}_check_n$n.err ;;
    bench)
      run 400 $port --steps 40 --warmup 5 --e2e-steps 6 > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/${tag}_bench_n$n.err
      echo "bench rc=$?"; tail -3 gpurun_out/${tag}_bench_n$n.err; timeout 60 python tools/show_bench.py gpurun_out/${tag}_bench_n$n.json ${tag}_n$n < /dev/null ;;
    nccl)
      run 400 $port --steps 40 --warmup 5 --device-only --collective nccl > gpurun_out/${tag}_bench_nccl_n$n.json 2> gpurun_out/${tag}_bench_nccl_n$n.err
      echo "bench nccl rc=$?"; tail -3 gpurun_out/${tag}_bench_nccl_n$n.err; timeout 60 python tools/show_bench.py gpurun_out/${tag}_bench_nccl_n$n.json ${tag}_nccl_n$n < /dev/null ;;
    arblocks)
      for b in 8 16 32 64 128; do
        echo "VOXE_ALLREDUCE_BLOCKS=$b"
        VOXE_ALLREDUCE_BLOCKS=$b run 200 $((port + b)) --check 2>/dev/null | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln)['collectives']; print({k[:40]: v.get('us') for k,v in d.items()})"
      done ;;
    onecpu)
      run 400 $port --steps 10 --warmup 3 --e2e-steps 6 --cpus-per-rank 1 > gpurun_out/${tag}_bench_onecpu_n$n.json 2> gpurun_out/${tag}_bench_onecpu_n$n.err
      echo "bench onecpu rc=$?"; tail -3 gpurun_out/${tag}_bench_onecpu_n$n.err; timeout 60 python tools/show_bench.py gpurun_out/${tag}_bench_onecpu_n$n.json ${tag}_onecpu_n$n < /dev/null ;;
    p2p)
      run 400 $port --steps 40 --warmup 5 --device-only --collective peer-p2p > gpurun_out/${tag}_bench_p2p_n$n.json 2> gpurun_out/${tag}_bench_p2p_n$n.err
      echo "bench p2p rc=$?"; tail -3 gpurun_out/${tag}_bench_p2p_n$n.err; timeout 60 python tools/show_bench.py gpurun_out/${tag}_bench_p2p_n$n.json ${tag}_p2p_n$n < /dev/null ;;
    cfg4)
      run 600 $port --workload cfg4 --steps 3 --warmup 3 > gpurun_out/${tag}_cfg4_n$n.json 2> gpurun_out/${tag}_cfg4_n$n.err
      echo "cfg4 rc=$?"; tail -3 gpurun_out/${tag}_cfg4_n$n.err; timeout 60 python tools/show_bench.py gpurun_out/${tag}_cfg4_n$n.json ${tag}_cfg4_n$n < /dev/null ;;
    cfg5)
      run 900 $port --workload cfg5 --steps 5 --warmup 3 > gpurun_out/${tag}_cfg5_n$n.json 2> gpurun_out/${tag}_cfg5_n$n.err
      echo "cfg5 rc=$?"; tail -3 gpurun_out/${tag}_cfg5_n$n.err; timeout 60 python tools/show_bench.py gpurun_out/${tag}_cfg5_n$n.json ${tag}_cfg5_n$n < /dev/null ;;
  esac
done
