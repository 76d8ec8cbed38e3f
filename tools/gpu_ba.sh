#!/bin/bash
# GPU round trip for the BA engine: parity tests, then the bench's BA blocks with the phase counters.
mkdir -p gpurun_out
if [ "$1" = "test" ]; then
  timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
fi
NRSLAM_B200_PROF=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_ba.json 2> gpurun_out/bench_ba.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_ba.json"))
print("value", d["value"], "e2e", d["e2e"]["value"])
for k in ("ba_window5", "ba", "ba_c4"):
    if k in d: print(k, {x: d[k][x] for x in ("value", "e2e_value", "launch_ms", "e2e_ms", "pcg_iterations", "grid_ctas")})
PY
grep -h "nrs prof" gpurun_out/bench_ba.err | awk '{k=$4" "$6; if(!(k in seen)){seen[k]=1; print}}' | head -8
