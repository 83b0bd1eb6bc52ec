#!/bin/bash
# Profiling recipe of /opt/skills/guides/B200_PROFILING.md for this repo (run under gpurun, one GPU).
# 1) plain run must exit 0, 2) launch list with per-launch device time, 3) one full capture of the dominant
# kernels (Cholesky update/panel GEMM, tcgen05 Gram).  Outputs land in gpurun_out/.
set -u
TAG=${1:-r01}
CMD="python bench.py --steps 1 --warmup 1 --pop 64 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
tail -1 gpurun_out/plain_$TAG.log | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:chol_gemm_kernel -s 76 -c 4 -o gpurun_out/prof_chol_$TAG -f $CMD > gpurun_out/ncu_chol_$TAG.log 2>&1
echo "chol capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gram_tc_kernel -s 1 -c 1 -o gpurun_out/prof_gram_$TAG -f $CMD > gpurun_out/ncu_gram_$TAG.log 2>&1
echo "gram capture rc=$?"
ls -la gpurun_out/
