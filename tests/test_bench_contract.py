"""bench.py's JSON contract on what can run without a GPU: the reference arm (the reference algorithm on the host cores)
on the tiny workload, and the bookkeeping helpers behind the roofline numbers."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


import pytest


@pytest.mark.parametrize("kind,workload", [("port", "tiny"), ("auto", "tiny"), ("auto", "tiny_3fold")])
def test_reference_arm_prints_one_contract_line(kind, workload):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload,
                          "--steps", "1", "--warmup", "1", "--cpu-sample", "2", "--cpu-kind", kind],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "gblup_fitness_evals_per_sec" and d["unit"] == "evals/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1
    assert d["config"]["workload"] == workload and d["config"]["pop_is"] == "per_gpu"
    have_ref = os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "tblup")) or os.path.isdir("/root/reference/tblup")
    assert d["cpu_baseline"]["kind"] == ("reference" if kind == "auto" and have_ref else "port")
    assert d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["gpu_launches"] == 0


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny",
                          "--gpus", "2", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120,
                         cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_algorithmic_work_formulas():
    sys.path.insert(0, ROOT)
    import bench
    # Cholesky update flops: sum over block columns equals n^3/3 to leading order
    n = 3200
    assert abs(bench.chol_update_flops(n, 64) / (n ** 3 / 3) - 1) < 0.05
    assert abs(bench.chol_update_flops(n, 256) / (n ** 3 / 3) - 1) < 0.15
    # update traffic: 8 B read-modify-write per entry below the first block column + the fp16 row operand
    b = bench.chol_update_bytes(n, 256)
    rmw = sum((n - c0) * min(256, n - c0) * 8 for c0 in range(256, n, 256))
    assert b > rmw and abs(b / 80.2e6 - 1) < 0.01
    # block column formed from int16 cross-products (2 B read + 4 B written per entry), and with t16 the rows below the
    # diagonal block of every full-width block column written as halves (2 + 2 B): the figures quoted in DESIGN.md §5
    assert abs(bench.chol_update_bytes(n, 256, 2) / 70.8e6 - 1) < 0.01
    b16 = bench.chol_update_bytes(n, 256, 2, True)
    below = sum((n - c0 - 256) * 256 for c0 in range(256, n, 256) if c0 + 256 < n)
    assert b16 == bench.chol_update_bytes(n, 256, 2) - 2 * below and abs(b16 / 62.9e6 - 1) < 0.01
