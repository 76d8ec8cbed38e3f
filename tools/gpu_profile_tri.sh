#!/bin/bash
# Round-1 late profile run: triangulation / graph-update tests, bench line, reference arm, launch list, full capture of
# the triangulation kernel.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_tri.py -q > gpurun_out/pytest_tri.log 2>&1; tail -4 gpurun_out/pytest_tri.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_full.json').read()); print({k:d[k] for k in ('value','ms_per_step','pcg_iterations_per_step')}, 'e2e', d['e2e']['value'], 'cpu', d.get('cpu_baseline',{}).get('value'), 'ba', d['ba']['value'], 'tri', d['triangulation'], 'graph', d['graph_update'])"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cut -c1-300 gpurun_out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-ba > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"; wc -l gpurun_out/r01_bench_launches.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:nrs_tri_kernel -s 1 -c 1 -o gpurun_out/r01_tri python tools/prof_tri.py > gpurun_out/ncu_tri.log 2>&1; echo "ncu tri rc=$?"; tail -2 gpurun_out/ncu_tri.log
ls -la gpurun_out/*.ncu-rep
