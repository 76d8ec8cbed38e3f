#!/bin/bash
# Round-end style GPU round trip on ONE GPU: parity tests, the driver's two bench commands, small logs only.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
NRSLAM_B200_PROF=1 timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
if [ "$1" = "ref" ]; then
  timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
fi
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_n1.json"))
print("value", d["value"], "e2e", d["e2e"], "cpu", d.get("cpu_baseline", {}).get("value"))
print("device_ms", d["device_ms"])
for k in ("ba_window5", "ba", "ba_c4"):
    if k in d: print(k, {x: d[k][x] for x in ("value", "e2e_value", "launch_ms", "e2e_ms", "pcg_iterations", "grid_ctas")})
PY
