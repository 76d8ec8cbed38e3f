#!/bin/bash
# ncu --set full capture of the wide BA kernel (configs[2] window); report lands in gpurun_out/.
mkdir -p gpurun_out
N=${1:-r02_lm_ba_wide}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nrs_lm_kernel_wide -s 1 -c 1 -f -o gpurun_out/$N python tools/prof_ba.py > gpurun_out/ncu_ba.log 2>&1; echo "ncu ba rc=$?"
tail -3 gpurun_out/ncu_ba.log; ls -la gpurun_out/$N.ncu-rep
