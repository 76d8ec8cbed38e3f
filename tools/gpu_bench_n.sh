#!/bin/bash
# bench.py under torchrun on N GPUs of one box, exactly as the driver launches it; the JSON line lands in gpurun_out/.
N=${1:-2}
mkdir -p gpurun_out
export NRSLAM_B200_XTIMEOUT_MS=15000
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"
tail -c 600 gpurun_out/bench_n$N.err | tail -3
python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_n$N.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"])
for k in ("ba_c4", "ba", "ba_sharded", "ba_sharded_c3"):
    x = d.get(k)
    if x: print(k, {a: x[a] for a in ("value", "e2e_value", "launch_ms", "e2e_ms", "single_gpu_launch_ms", "per_rank_launch_ms", "poses_identical", "max_pose_diff", "max_point_diff", "chi2_trace_equal_1e-6") if a in x})
PY
