#!/bin/bash
# first GPU pass of round 2: bench (default), microbench peaks, reference arm, full GPU test suite
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02a_gpus.txt
free -g >> gpurun_out/r02a_gpus.txt; nproc >> gpurun_out/r02a_gpus.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02a_bench.log 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02a_bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r02a_ref.log 2> gpurun_out/r02a_ref.err; echo "ref rc=$?"
timeout 2400 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r02a_tests.log 2>&1; echo "tests rc=$?"
tail -40 gpurun_out/r02a_tests.log
