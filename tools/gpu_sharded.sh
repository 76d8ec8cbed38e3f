#!/bin/bash
# Multi-GPU round trip (gpurun --gpus N): landmark-sharded BA vs single-GPU BA, every run bounded by `timeout`.
N=${1:-2}
mkdir -p gpurun_out
export NRSLAM_B200_XTIMEOUT_MS=15000
run() {
  name=$1; shift
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      tools/sharded_ba_check.py "$@" > gpurun_out/sharded_${name}_n$N.log 2>&1
  echo "== $name rc=$?"; grep -h '^{' gpurun_out/sharded_${name}_n$N.log | tail -1; grep -v '^{' gpurun_out/sharded_${name}_n$N.log | grep -i "error\|Traceback\|assert" | head -5
}
nvidia-smi -L | head -8
[ -z "$SKIP_SMALL" ] && run c1 --config c1
[ -z "$SKIP_SMALL" ] && run c3small --config c3 --landmarks 1500 --keyframes 10 --visible 6
[ -n "$FULL" ] && run c3 --config c3 --reps 3
[ -n "$FULL4" ] && run c4 --config c4 --reps 2 --pose-tol 2e-5 --pt-tol 4e-4
true
