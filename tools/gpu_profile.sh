#!/bin/bash
# Round profile run: bench line, launch list of the same bench command, full captures of the two dominant kernels.
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"
cat gpurun_out/bench_full.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','pcg_iterations_per_step')}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'no-klt', d['without_klt'], 'klt', d['klt'], 'cpu', d.get('cpu_baseline',{}).get('value'), 'ba', d['ba']['value'], 'frac', d['roofline']['frac'])"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cat gpurun_out/bench_reference.json | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-ba > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"; wc -l gpurun_out/r01_bench_launches.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:nrs_lm -s 4 -c 1 -o gpurun_out/r01_lm_track_final python tools/prof_track.py track 2 > gpurun_out/ncu_lm.log 2>&1; echo "ncu lm rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:klt_track -s 1 -c 1 -o gpurun_out/r01_klt_track python tools/prof_klt.py > gpurun_out/ncu_klt.log 2>&1; echo "ncu klt rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:nrs_lm_kernel_wide -s 1 -c 1 -o gpurun_out/r01_lm_ba_wide python tools/prof_ba.py > gpurun_out/ncu_ba.log 2>&1; echo "ncu ba rc=$?"
ls -la gpurun_out/*.ncu-rep
