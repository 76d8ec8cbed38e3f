#!/bin/bash
# Round-2 evidence run (one GPU): ncu captures of the exact-solve tracking kernel and of the kernels either side of it,
# the launch list of the bench command, and compute-sanitizer runs over the hand-rolled synchronisation (grid / team
# barriers, look-ahead warp, TMA + mbarrier, cluster-native CG loop). Small files only land in gpurun_out/ (64 MiB cap).
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:nrs_track_direct -c 1 -f -o gpurun_out/r02_track_direct python tools/prof_track.py track 1 > gpurun_out/ncu_direct.log 2>&1; echo "ncu direct rc=$?"
ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats --clock-control none -c 80 -f -o /tmp/r02_misc python tools/prof_misc.py > gpurun_out/ncu_misc.log 2>&1; echo "ncu misc rc=$?"
python tools/ncu_summary.py /tmp/r02_misc.ncu-rep gpurun_out/r02_misc_kernels_ncu_summary.json; echo "misc summary rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-ba > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"; wc -l gpurun_out/r02_bench_launches.csv
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/prof_track.py track 1 > gpurun_out/r02_sanitizer_${tool}_track.log 2>&1; echo "sanitizer $tool track rc=$?"; tail -2 gpurun_out/r02_sanitizer_${tool}_track.log
done
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/prof_ba_small.py > gpurun_out/r02_sanitizer_${tool}_ba_cluster.log 2>&1; echo "sanitizer $tool ba rc=$?"; tail -2 gpurun_out/r02_sanitizer_${tool}_ba_cluster.log
done
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_*
