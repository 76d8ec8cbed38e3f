#!/bin/bash
# compute-sanitizer over a 2-GPU landmark-sharded BA (fused grid reduction + NVLink exchange), bounded by timeouts.
mkdir -p gpurun_out
export NRSLAM_B200_XTIMEOUT_MS=60000
for tool in memcheck racecheck; do
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 \
      --no-python compute-sanitizer --tool $tool --print-limit 10 python tools/sharded_ba_check.py --config c3 --landmarks 1500 --keyframes 10 --visible 6 --reps 1 \
      > gpurun_out/r02_sanitizer_${tool}_ba_sharded_n2.log 2>&1
  echo "sanitizer $tool sharded rc=$?"; grep -h "ERROR SUMMARY\|RACECHECK SUMMARY\|\"ok\"" gpurun_out/r02_sanitizer_${tool}_ba_sharded_n2.log | cut -c1-160
done
