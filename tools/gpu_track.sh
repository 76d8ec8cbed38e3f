#!/bin/bash
# One GPU round trip for the tracking engine: parity tests (optional, arg 1 = "test"), then bench with the phase counters.
mkdir -p gpurun_out
if [ "$1" = "test" ]; then
  timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
fi
NRSLAM_B200_PROF=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu 2> gpurun_out/bench.err > gpurun_out/bench.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "grid", d["config"].get("grid_ctas"))
names = {0: "stage_ab", 1: "stage_c", 2: "backward", 3: "ab.zero+orig", 4: "ab.pull", 5: "linearise", 6: "solve", 7: "update+chi2",
         8: "ab.factor", 9: "ab.store", 10: "bw.load+gemv", 11: "bw.subst", 12: "barriers", 13: "ab.factor.R", 14: "ab.factor.T-lookahead(warp0)", 15: "total"}
for line in open("gpurun_out/bench.err"):
    if "nrs prof] grid 128" in line or "nrs prof] grid 64" in line:
        v = [int(x) for x in line.split("cycles:")[1].split()]
        print(line.split("cycles:")[0].strip())
        for k in sorted(names):
            print("   %-14s %10d  %5.1f %%  %8.1f us" % (names[k], v[k], 100.0 * v[k] / max(v[15], 1), v[k] / 1965.0))
        break
PY
grep -h "nrs plev" gpurun_out/bench.err | tail -24
