#!/bin/bash
# Round-end evidence on ONE GPU: parity tests, ncu capture + launch list of the wide BA kernel and the bench command,
# compute-sanitizer over the wide BA variant, the configs[4] sweep. Bench lines: tools/gpu_round.sh.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nrs_lm_kernel_wide -s 1 -c 1 -f -o gpurun_out/r02_lm_ba_wide python tools/prof_ba.py > gpurun_out/ncu_ba.log 2>&1; echo "ncu ba rc=$?"; grep "BA c3" gpurun_out/ncu_ba.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"; wc -l gpurun_out/r02_bench_launches.csv
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/prof_ba_wide_small.py > gpurun_out/r02_sanitizer_${tool}_ba_wide.log 2>&1; echo "sanitizer $tool wide rc=$?"; tail -3 gpurun_out/r02_sanitizer_${tool}_ba_wide.log
done
timeout 900 python bench.py --sweep > gpurun_out/r02_sweep_n1.json 2> gpurun_out/sweep.err; echo "sweep rc=$?"; tail -9 gpurun_out/sweep.err
ls -la gpurun_out/*.ncu-rep
