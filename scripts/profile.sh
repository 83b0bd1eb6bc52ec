#!/bin/bash
# Profiling recipe of /opt/skills/guides/B200_PROFILING.md for this repo (run under gpurun, one GPU).
# 1) plain run must exit 0, 2) launch list with per-launch device time, 3) one full capture of the dominant
# kernels.  Outputs land in gpurun_out/.   usage: scripts/profile.sh <tag> [mixed|fp64]
set -u
TAG=${1:-r01}
PREC=${2:-mixed}
POP=${3:-64}
CMD="python bench.py --steps 1 --warmup 1 --pop $POP --no-cpu-baseline --precision $PREC"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
tail -1 gpurun_out/plain_$TAG.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
if [ "$PREC" = "mixed" ]; then
  ncu --set full --clock-control none --import-source on -k regex:solve_mixed_kernel -s 0 -c 1 -o gpurun_out/prof_solve_$TAG -f $CMD > gpurun_out/ncu_solve_$TAG.log 2>&1
  echo "solve capture rc=$?"
  # 64th tf32_gemm launch of the first evaluation = outer update of block column J = 8 (K = 2048, N = 256)
  ncu --set full --clock-control none --import-source on -k regex:tf32_gemm_kernel -s 63 -c 2 -o gpurun_out/prof_tf32gemm_$TAG -f $CMD > gpurun_out/ncu_tf32_$TAG.log 2>&1
  echo "tf32 gemm capture rc=$?"
else
  ncu --set full --clock-control none --import-source on -k regex:chol_gemm_kernel -s 76 -c 4 -o gpurun_out/prof_chol_$TAG -f $CMD > gpurun_out/ncu_chol_$TAG.log 2>&1
  echo "chol capture rc=$?"
fi
ncu --set full --clock-control none --import-source on -k regex:gram_tc_kernel -s 1 -c 1 -o gpurun_out/prof_gram_$TAG -f $CMD > gpurun_out/ncu_gram_$TAG.log 2>&1
echo "gram capture rc=$?"
ls -la gpurun_out/ | tail -12
