#!/bin/bash
# 2-GPU pass: knockout / clone / 2-GPU evaluator tests, then the bench under torchrun at N=2 (weak primary, strong inside)
mkdir -p gpurun_out
nvidia-smi -L
timeout 1200 python -m pytest tests/test_gpu_knockout.py tests/test_gpu_packed.py tests/test_gpu_large.py tests/test_gpu_de.py -m gpu -x -q --durations=8 > gpurun_out/r02d_tests.log 2>&1; echo "tests rc=$?"
tail -22 gpurun_out/r02d_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02d_bench_n2.log 2> gpurun_out/r02d_bench_n2.err; echo "bench n2 rc=$?"
tail -c 800 gpurun_out/r02d_bench_n2.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r02d_bench_n2.log") if l.startswith("{")][-1])
    print("N=2 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "parity", d["parity_ok"], d["details"]["cross_product_storage"], d["details"]["scaled_matrix_formed_inside_cholesky_updates"])
    print("other", json.dumps(d["details"]["other_scaling"]))
except Exception as e:
    print("no line", e)
PY
