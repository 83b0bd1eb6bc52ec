#!/bin/bash
# Profiling recipe of /opt/skills/guides/B200_PROFILING.md for this repo (run under gpurun, one GPU).
# 1) plain run must exit 0, 2) launch list with per-launch device time, 3) one full capture of each dominant kernel.
# Outputs land in gpurun_out/.   usage: scripts/profile.sh <tag> [mixed|fp64] [pop]
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
FULL="ncu --set full --clock-control none --import-source on"
if [ "$PREC" = "mixed" ]; then
  $FULL -k regex:solve_mixed_kernel -s 0 -c 1 -o gpurun_out/prof_solve_$TAG -f $CMD > gpurun_out/ncu_solve_$TAG.log 2>&1
  echo "solve capture rc=$?"
  # tf32_gemm_kernel launches of one evaluation, in order: block column 0 has 7 (narrow solves / updates of the
  # diagonal block + the wide solve), every later block column J has 8, starting with its outer update.
  # J = 6: launch 7 + 5 * 8 = 47 is the outer update (fp16 operands, K = 1536, N = 256), launch 54 the wide
  # triangular solve (tf32, K = 256, N = 256)
  $FULL -k regex:tf32_gemm_kernel -s 47 -c 1 -o gpurun_out/prof_update_$TAG -f $CMD > gpurun_out/ncu_update_$TAG.log 2>&1
  echo "update capture rc=$?"
  $FULL -k regex:tf32_gemm_kernel -s 54 -c 1 -o gpurun_out/prof_trsm_$TAG -f $CMD > gpurun_out/ncu_trsm_$TAG.log 2>&1
  echo "wide trsm capture rc=$?"
else
  $FULL -k regex:chol_gemm_kernel -s 76 -c 4 -o gpurun_out/prof_chol_$TAG -f $CMD > gpurun_out/ncu_chol_$TAG.log 2>&1
  echo "chol capture rc=$?"
fi
$FULL -k regex:gram_tc_kernel -s 1 -c 1 -o gpurun_out/prof_gram_$TAG -f $CMD > gpurun_out/ncu_gram_$TAG.log 2>&1
echo "gram capture rc=$?"
ls -la gpurun_out/ | grep $TAG
