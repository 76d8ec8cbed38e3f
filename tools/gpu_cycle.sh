#!/bin/bash
# One GPU round trip: parity tests, then bench variants (env overrides given as arguments, "-" = none).
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
for v in "$@"; do
  [ "$v" = "-" ] && v="NRSLAM_B200_DUMMY=1"
  echo "== $v"
  env $v NRSLAM_B200_PROF=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu 2> "gpurun_out/bench_$v.err" | tee "gpurun_out/bench_$v.json" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','pcg_iterations_per_step')}, 'e2e ms', d['e2e']['ms_per_step'], 'grid', d['config']['grid_ctas'], d['config']['block_threads'], 'BA it/s', d['ba']['value'], 'ms', d['ba']['launch_ms'], 'pcg', d['ba']['pcg_iterations'])"
  grep -h "nrs prof" "gpurun_out/bench_$v.err" | awk '{k=$4" "$6; if(!(k in seen)){seen[k]=1; print}}' | head -6
done
