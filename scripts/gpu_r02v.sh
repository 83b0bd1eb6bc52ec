#!/bin/bash
# round 2: fused diagonal-block chain for small waves (A/B at 13 / 125 / 296 genomes per GPU)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r02v_tests.log 2>&1; echo tests rc=$?; tail -5 gpurun_out/r02v_tests.log
for v in 1000000 0; do
# (chain_fused: jobs per wave up to which the fused chain kernel runs)
timeout 300 python scripts/sweep.py --ks 5001 --pops 13,125,148,296 --steps 10 --opt chain_fused=$v --out gpurun_out/r02v_sweep_fused$v.json 2>&1 | grep -v "^crossover" | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()[:300]); continue
    print('chain_fused=$v pop', d['pop'], 'ms', round(d['ms_per_step'], 3), 'panel', d['stage_ms']['chol_panel'], 'update', d['stage_ms']['chol_update'], 'solve', d['stage_ms']['solve'], 'fallbacks', d['fp64_fallbacks'])
"
done
