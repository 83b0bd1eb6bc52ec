#!/bin/bash
# Profiling recipe of /opt/skills/guides/B200_PROFILING.md for this repo (run under gpurun, one GPU).
# 1) plain run must exit 0, 2) launch list with per-launch device time, 3) one full capture of each dominant kernel.
# Outputs land in gpurun_out/.   usage: scripts/profile.sh <tag> [mixed|fp64] [pop]
set -u
TAG=${1:-r01}
PREC=${2:-mixed}
POP=${3:-64}
CMD="python bench.py --steps 1 --warmup 1 --pop $POP --no-cpu-baseline --no-parity --no-sustained-peaks --precision $PREC"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
tail -1 gpurun_out/plain_$TAG.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
FULL="ncu --set full --clock-control none --import-source on"
if [ "$PREC" = "mixed" ]; then
  $FULL -k regex:solve_mixed_kernel -s 0 -c 1 -o gpurun_out/prof_solve_$TAG -f $CMD > gpurun_out/ncu_solve_$TAG.log 2>&1
  echo "solve capture rc=$?"
  # tf32_gemm_kernel launches of one evaluation, in order: every full block column J has 2 -- its outer update (J = 0: the
  # K = 0 launch that only forms the block column from the cross-products), then its wide panel GEMM.  J = 6: launch 12 is
  # the outer update (fp16 operands, K = 1536, N = 256, block column formed from the int16 cross-products, rows below the
  # diagonal block written as halves), launch 13 the wide panel GEMM (fp16 operands, K = N = 256, in place in L16)
  $FULL -k regex:tf32_gemm_kernel -s 12 -c 1 -o gpurun_out/prof_update_$TAG -f $CMD > gpurun_out/ncu_update_$TAG.log 2>&1
  echo "update capture rc=$?"
  $FULL -k regex:tf32_gemm_kernel -s 13 -c 1 -o gpurun_out/prof_trsm_$TAG -f $CMD > gpurun_out/ncu_trsm_$TAG.log 2>&1
  echo "wide trsm capture rc=$?"
else
  $FULL -k regex:chol_gemm_kernel -s 76 -c 4 -o gpurun_out/prof_chol_$TAG -f $CMD > gpurun_out/ncu_chol_$TAG.log 2>&1
  echo "chol capture rc=$?"
fi
$FULL -k regex:gram_tc_kernel -s 1 -c 1 -o gpurun_out/prof_gram_$TAG -f $CMD > gpurun_out/ncu_gram_$TAG.log 2>&1
echo "gram capture rc=$?"
# raw metric pages as CSV (what profiles/ keeps; the .ncu-rep files stay in gpurun_out/)
for k in solve update trsm gram chol; do
  if [ -f gpurun_out/prof_${k}_$TAG.ncu-rep ]; then
    ncu -i gpurun_out/prof_${k}_$TAG.ncu-rep --page raw --csv > gpurun_out/${TAG}_${k}_raw.csv 2>/dev/null
  fi
done
ls -la gpurun_out/ | grep $TAG
