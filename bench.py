#!/usr/bin/env python
"""Benchmark of the GBLUP fitness path: metric = fitness evals/s (individuals x folds), BASELINE.json.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host cores

A "step" is one generation's evaluation: P genomes (uniform random k-subsets) scored on one train/validation
split.  Workload (config 2 of BASELINE.json): 5 000 animals x 50 000 markers, k = 5 001 markers per genome
(~10 %; k > n so the reference's blup() takes its GBLUP branch, tblup/evaluator.py:257), pop = 1 000 per GPU.
Weak scaling: every rank holds a replica of the genotypes and evaluates its own P genomes; the only
collective is the all-gather of the fitness vector.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: animals, markers, k, population per GPU, folds
    "c2_5000x50000_k5001_pop1000": dict(n=5000, m=50000, k=5001, pop=1000, folds=1),
    "c1_1000x10000_k1500_pop50": dict(n=1000, m=10000, k=1500, pop=50, folds=1),
    "c3_5000x50000_k5001_pop1000_10fold": dict(n=5000, m=50000, k=5001, pop=1000, folds=10),
    # config 4 (large-n regime): full shape, population cut to what one short step needs; no CPU arm (minutes/genome)
    "c4_20000x500000_k50000_pop32": dict(n=20000, m=500000, k=50000, pop=32, folds=1, fast_synth=True, no_cpu=True),
    "tiny": dict(n=300, m=2000, k=400, pop=16, folds=1),
}
DEFAULT_WORKLOAD = "c2_5000x50000_k5001_pop1000"
# DRAM traffic per unit (matrix / genome) of the dominant kernels at the C2 shape, from the committed `ncu --set full`
# captures (profiles/README.md says how each was taken); filled in after every re-profile
NCU_TRAFFIC = {
    "solve_c32": {"bytes_per_unit": 164.06e6, "source": "profiles/r01p_solve_raw.csv"},
    "solve_c16": {"bytes_per_unit": 112.05e6, "source": "profiles/r01q_solve_raw.csv (1 000 matrices per launch)"},
    "gram_fused_c16": {"bytes_per_unit": 58.93e6, "source": "profiles/r01q_gram_raw.csv (1 000 genomes per launch)"},
    "gram_fp4_fused_c16": {"bytes_per_unit": 48.43e6, "source": "profiles/r01r_gram_raw.csv (1 000 genomes per launch)"},
}
H2 = 0.4
METRIC = "gblup_fitness_evals_per_sec"
UNIT = "evals/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--pop", type=int, default=0, help="override genomes per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="mixed", choices=["mixed", "fp64"],
                    help="mixed: TF32 tensor-core Cholesky preconditioner + fp64 refinement (default); fp64: fp64 Cholesky")
    ap.add_argument("--storage", default="packed2", choices=["int8", "packed2"],
                    help="resident genotype format: int8 dosages, or 2 bits per dosage (bit-identical results)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="individuals in the CPU sample (default: one per core)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks (sampled DURING the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), threading.Event()
        self.sm_max = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def result(self):
        self.stop_flag.set()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (numpy/scipy/sklearn calls of tblup/evaluator.py:244-314 as restated
# in oracle/gblup_oracle.py) on the host cores, one single-threaded BLAS worker per core -- the
# reference's own deployment (generate_sbs.py:25: OMP_NUM_THREADS=1, one process per core).
# ------------------------------------------------------------------------------------------------
def _cpu_worker(job):
    geno_path, y, train, valid, idx, h2 = job
    from oracle import gblup_oracle as O
    data = np.load(geno_path, mmap_mode="r")
    t0 = time.perf_counter()
    f = O.ref_blup(np.asarray(idx).astype(int), list(train), list(valid), data, y, h2)
    return float(f), time.perf_counter() - t0


class CpuArm:
    def __init__(self, x, y, train, valid, cores):
        import multiprocessing as mp
        import tempfile
        for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
            os.environ[v] = "1"
        base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
        self.tmp = tempfile.TemporaryDirectory(dir=base)
        self.path = os.path.join(self.tmp.name, "geno.npy")
        mm = np.lib.format.open_memmap(self.path, mode="w+", dtype=np.float64, shape=x.shape)
        for r0 in range(0, x.shape[0], 256):
            mm[r0:r0 + 256] = x[r0:r0 + 256]
        mm.flush()
        del mm
        self.y, self.train, self.valid, self.cores = y, np.asarray(train), np.asarray(valid), cores
        self.pool = mp.get_context("spawn").Pool(cores)

    def step(self, genomes):
        jobs = [(self.path, self.y, self.train, self.valid, g, H2) for g in genomes]
        t0 = time.perf_counter()
        out = self.pool.map(_cpu_worker, jobs, chunksize=1)
        return time.perf_counter() - t0, [o[0] for o in out]

    def close(self):
        self.pool.terminate()
        self.tmp.cleanup()


def chol_update_flops(ntp, nb=64):
    """Algorithmic flops of the (outer) Cholesky update launches of one matrix, lower triangle only: block column
    starting at c0 (width w = min(nb, ntp - c0)) contributes 2 * c0 * [w (w+1)/2 + (ntp - c0 - w) w]."""
    tot = 0
    for c0 in range(nb, ntp, nb):
        w = min(nb, ntp - c0)
        tot += 2 * c0 * (w * (w + 1) // 2 + (ntp - c0 - w) * w)
    return tot


def chol_update_bytes(ntp, ob=256):
    """Algorithmic DRAM bytes of the outer Cholesky updates of one matrix in mixed precision: the fp32 block column
    is read and written once (8 B per entry of the rows at and below the diagonal block) and the fp16 row operand
    L[rows, 0:c0] is streamed once per block column (2 B per entry); the 256-row column operand stays in L2."""
    tot = 0
    for c0 in range(ob, ntp, ob):
        w = min(ob, ntp - c0)
        rows = ntp - c0
        tot += rows * w * 8 + rows * c0 * 2
    return tot


def main():
    args = parse()
    wl = dict(WORKLOADS[args.workload])
    if args.pop:
        wl["pop"] = args.pop
    n, m, k, P, folds = wl["n"], wl["m"], wl["k"], wl["pop"], wl["folds"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from tblup_b200 import synth
    cores = len(os.sched_getaffinity(0))

    if args.impl == "reference":
        if rank != 0:
            return 0
        x, y = synth.synth_dataset(n, m, h2=H2, seed=0)
        train, valid, test = synth.split_indices(n, seed=0)
        sample = args.cpu_sample or cores
        arm = CpuArm(x, y, train, valid, cores)
        try:
            flat, off = synth.random_genomes(sample * (args.steps + args.warmup), m, k, seed=100)
            gens = [flat[off[i]:off[i + 1]] for i in range(off.size - 1)]
            pos = 0
            for _ in range(args.warmup):
                arm.step(gens[pos:pos + sample])
                pos += sample
            t = 0.0
            for _ in range(args.steps):
                dt, _f = arm.step(gens[pos:pos + sample])
                t += dt
                pos += sample
        finally:
            arm.close()
        value = sample * args.steps * 1 / t
        line = {
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "animals": n, "markers": m, "k": k, "pop_per_gpu": P, "folds": folds,
                       "h2": H2},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d genomes per step (one per core), single split; reference algorithm "
                                       "(numpy matmul + inv + pearsonr, oracle.ref_blup) with 1 BLAS thread per "
                                       "worker process" % sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    # ---------------------------------------------------------------- our arm
    import torch
    import torch.distributed as dist
    from tblup_b200 import GblupEngine, MODE_AUTO
    from tblup_b200 import engine as E

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the tblup_b200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    x, y = (synth.synth_dataset_fast if wl.get("fast_synth") else synth.synth_dataset)(n, m, h2=H2, seed=0)
    train, valid, test = synth.split_indices(n, seed=0)
    slots = [0]
    eng = GblupEngine(x, y, perm=np.concatenate([train, valid, test]), device=local_rank, storage=args.storage)
    if folds == 1:
        eng.set_rowset(0, train, valid)
    else:
        # intra-generation k-fold over the training animals (tblup/evaluator.py:455-483, :509-537)
        bounds = np.linspace(0, 0, 1)
        sizes = [len(train) // folds + (1 if f < len(train) % folds else 0) for f in range(folds)]
        bounds = np.concatenate([[0], np.cumsum(sizes)])
        slots = list(range(folds))
        for f in range(folds):
            va = train[bounds[f]:bounds[f + 1]]
            tr = np.concatenate([train[:bounds[f]], train[bounds[f + 1]:]])
            eng.set_rowset(f, tr, va)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    eng.set_precision(args.precision)

    n_batches = 2
    batches = [synth.random_genomes(P, m, k, seed=1000 + 17 * rank + b) for b in range(n_batches)]
    pinned = []
    for flat, off in batches:
        pf = torch.from_numpy(flat).pin_memory()
        po = torch.from_numpy(off).pin_memory()
        pinned.append((pf, po))
    fit_dev = torch.empty(P * len(slots), dtype=torch.float64, device="cuda")
    fit_all = torch.empty(world * P * len(slots), dtype=torch.float64, device="cuda") if world > 1 else None
    fit_host = torch.empty(P * len(slots), dtype=torch.float64).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(b):
        eng.evaluate_staged(slots, h2=H2, mode=MODE_AUTO, out_device_ptr=fit_dev.data_ptr())
        if world > 1:
            dist.all_gather_into_tensor(fit_all, fit_dev)

    def step_e2e(b):
        pf, po = pinned[b % n_batches]
        eng._check(eng._lib.tb_eval(eng._ctx, np.asarray(slots, dtype=np.int32).ctypes.data, len(slots),
                                    pf.data_ptr(), po.data_ptr(), P, H2, MODE_AUTO, fit_host.data_ptr()), "tb_eval")

    # -- value: inputs resident in HBM -----------------------------------------------------------
    dmma_peak = eng.microbench(0)
    eng.stage(flat=batches[0][0], off=batches[0][1])
    for i in range(args.warmup):
        step_resident(i)
    eng.reset_counters()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        step_resident(i)
    e1.record(stream)
    barrier()
    clocks = sampler.result()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count()

    # -- the same K steps again with per-stage CUDA events (stage split + roofline of the dominant kernel) ------
    eng.set_option("profile", 1)
    eng.reset_counters()
    barrier()
    e0.record(stream)
    for i in range(args.steps):
        step_resident(i)
    e1.record(stream)
    barrier()
    ms_instrumented = e0.elapsed_time(e1)
    stage = eng.stage_times()
    eng.set_option("profile", 0)
    wave = eng.last_wave()
    precision = eng.last_precision()
    mean_sweeps = None
    if precision == "mixed":
        last = (P - 1) % wave + 1            # jobs in the last wave (the one the debug view points at)
        mean_sweeps = float(np.mean([eng.debug_fetch(E.DBG_SWEEPS, j)[0] for j in range(min(last * len(slots), 64))]))

    # -- e2e: host buffers through the public C-ABI call, copies inside the timed region -----------
    for i in range(max(1, min(args.warmup, 2))):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    e0.record(stream)
    for i in range(args.steps):
        step_e2e(i)
    e1.record(stream)
    barrier()
    ms_e2e = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0))

    if world > 1:
        t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])

    # parity gate run with every benchmark: 64 genomes of the timed batch through the default path and
    # through the fp64 Cholesky path of the same library (two independent factorisations of the same exact integers)
    parity = None
    if rank == 0 and precision == "mixed":
        n_par = min(64, P)
        pf, po = batches[0]
        sub_flat, sub_off = pf[:po[n_par]], po[:n_par + 1]
        eng.stage(flat=sub_flat, off=sub_off)
        f_mixed = eng.evaluate_staged(slots, h2=H2, mode=MODE_AUTO)
        fallbacks = eng.info("last_fallbacks")
        eng.set_precision("fp64")
        f_fp64 = eng.evaluate_staged(slots, h2=H2, mode=MODE_AUTO)
        eng.set_precision(args.precision)
        parity = {"genomes": int(n_par), "folds": len(slots), "bar_abs": 1e-6,
                  "max_abs_fitness_diff_default_vs_fp64_path": float(np.nanmax(np.abs(f_mixed - f_fp64))),
                  "fp64_fallbacks_in_sample": int(fallbacks)}
        eng.stage(flat=batches[0][0], off=batches[0][1])

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not wl.get("no_cpu"):
        sample = args.cpu_sample or cores
        arm = CpuArm(x, y, train, valid, cores)
        try:
            flat, off = batches[0]
            gens = [flat[off[i]:off[i + 1]] for i in range(sample)]
            dt, cpu_fit = arm.step(gens)
        finally:
            arm.close()
        eng.stage(flat=batches[0][0], off=batches[0][1])
        gpu_fit = eng.evaluate_staged([0], h2=H2, mode=MODE_AUTO)[:sample, 0]
        cpu_baseline = {"value": sample / dt, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "%d genomes of the same batch (one per core), single split, oracle.ref_blup "
                                  "(the reference's numpy/scipy calls) with 1 BLAS thread per worker process; "
                                  "%.1f s wall" % (sample, dt),
                        "max_abs_fitness_diff_vs_gpu": float(np.abs(np.asarray(cpu_fit) - gpu_fit).max())}

    if rank == 0:
        evals = world * P * len(slots) * args.steps
        n_t = len(train) if folds == 1 else len(train) - len(train) // folds
        ntp = (n_t + 63) // 64 * 64
        upd_ms, upd_launches = stage["chol_update"]
        n_mats = P * len(slots) * args.steps
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        bf16 = peaks.get("bf16_tflops_sustained") or 1400.0
        bf16_src = ("bf16_tflops_sustained of MEASURED_PEAKS.json" if peaks.get("bf16_tflops_sustained")
                    else "fallback 1400 TFLOP/s sustained (B200_PROFILING.md)")
        c16 = bool(eng.info("last_c16"))
        fused = bool(eng.info("last_fused_scale"))
        fp4 = bool(eng.info("last_fp4"))
        if precision == "mixed":
            upd_flops = chol_update_flops(ntp, 256) * n_mats
            upd_kernel = ("tf32_gemm_kernel<F16> (outer left-looking Cholesky update on the fp16 copy of the factor, "
                          "tcgen05 kind::f16, M128 x N256, fp32 accumulate)")
            upd_peak = bf16
            upd_peak_src = bf16_src + " (fp16 operands run at the bf16 rate)"
        else:
            upd_flops = chol_update_flops(ntp, 64) * n_mats
            upd_kernel = "chol_gemm_kernel<0> (left-looking Cholesky update, fp64 DMMA)"
            upd_peak = dmma_peak
            upd_peak_src = ("fp64 mma.sync issue-rate probe run in this process (MEASURED_PEAKS.json has no fp64 "
                            "entry; B200 nominal fp64 is 37 TFLOP/s)")
        achieved = upd_flops / (upd_ms * 1e-3) / 1e12 if upd_ms > 0 else None
        gram_ms, gram_launches = stage["gram"]
        n_v = len(valid) if folds == 1 else len(train) // folds
        rows_t = len(train)   # Gram covers the union of the fold rows once per individual
        gram_ops = 2.0 * k * (rows_t * (rows_t + 1) / 2 + (len(valid) * rows_t if folds == 1 else 0)) * P * args.steps
        hbm = peaks.get("hbm_gbs") or 6650.0
        rl_update = {"bound": "tensor", "kernel": upd_kernel, "achieved": achieved, "peak": upd_peak, "unit": "TFLOP/s",
                     "frac": (achieved / upd_peak) if achieved and upd_peak else None, "traffic": None,
                     "peak_source": upd_peak_src, "launches": int(upd_launches),
                     "avg_launch_ms": upd_ms / max(1, upd_launches), "share_of_step": upd_ms / ms_instrumented}
        if precision == "mixed" and upd_ms > 0:
            # with fp16 operands the update is HBM-bound (ncu: 79 % of the copy bandwidth, tensor pipe 46 % active):
            # report it against the bandwidth roofline and keep the flop rate beside it
            upd_bytes = chol_update_bytes(ntp, 256)
            upd_gbs = upd_bytes * n_mats / (upd_ms * 1e-3) / 1e9
            rl_update = {"bound": "hbm", "kernel": upd_kernel, "achieved": upd_gbs,
                         "peak": peaks.get("hbm_gbs") or 6650.0, "unit": "GB/s",
                         "frac": upd_gbs / (peaks.get("hbm_gbs") or 6650.0), "traffic": None,
                         "algorithmic_bytes_per_matrix": upd_bytes,
                         "peak_source": "hbm_gbs of MEASURED_PEAKS.json" if peaks.get("hbm_gbs") else "fallback 6650 GB/s",
                         "tensor_tflops": achieved, "tensor_frac_of_bf16_sustained": achieved / upd_peak,
                         "launches": int(upd_launches), "avg_launch_ms": upd_ms / max(1, upd_launches),
                         "share_of_step": upd_ms / ms_instrumented}
        solve_ms, solve_launches = stage["solve"]
        csz = 2 if c16 else 4
        tri_c = n_t * (n_t + 1) / 2 * csz            # lower triangle of the stored cross-products
        tri_h = n_t * (n_t + 1) / 2 * 2              # lower triangle of the fp16 copy of the factor
        if precision == "mixed":
            # per matrix: (1 + sweeps) preconditioner applications (fp16 factor read forwards and backwards), `sweeps`
            # symmetric mat-vecs on the integer cross-products (lower triangle read by rows and by columns), one
            # pass over the validation rows
            per_mat = (1 + mean_sweeps) * 2 * tri_h + mean_sweeps * 2 * tri_c + n_v * n_t * csz
            solve_kernel = ("solve_mixed_kernel (blocked substitution with the fp16 copy of the TF32 factor + fp64 "
                            "refinement on the %s cross-products + predictions + Pearson)" % ("int16" if c16 else "int32"))
        else:
            per_mat = 2 * n_t * (n_t + 1) / 2 * 8 + n_v * n_t * 8
            solve_kernel = "solve_kernel (fp64 blocked substitution + predictions + Pearson)"
        solve_gbs = per_mat * n_mats / (solve_ms * 1e-3) / 1e9 if solve_ms > 0 else None
        mats_per_launch = n_mats / max(1, solve_launches)
        # measured DRAM bytes per launch from the committed ncu captures (dram__bytes_read.sum + dram__bytes_write.sum of
        # one `ncu --set full` launch, scaled to the matrices / genomes of one bench launch); None for other shapes
        ncu = NCU_TRAFFIC if (precision == "mixed" and n_t == 3200 and k == 5001 and folds == 1) else {}
        solve_traffic = ncu.get("solve_c16" if c16 else "solve_c32")
        rl_solve = {"bound": "hbm", "kernel": solve_kernel, "achieved": solve_gbs, "peak": hbm, "unit": "GB/s",
                    "frac": (solve_gbs / hbm) if solve_gbs else None,
                    "traffic": solve_traffic["bytes_per_unit"] * mats_per_launch if solve_traffic else None,
                    "traffic_source": solve_traffic["source"] if solve_traffic else None,
                    "algorithmic_bytes_per_launch": per_mat * mats_per_launch,
                    "peak_source": "hbm_gbs of MEASURED_PEAKS.json" if peaks.get("hbm_gbs") else "fallback 6650 GB/s",
                    "algorithmic_bytes_per_matrix": per_mat, "mean_refinement_sweeps": mean_sweeps,
                    "launches": int(solve_launches), "avg_launch_ms": solve_ms / max(1, solve_launches),
                    "share_of_step": solve_ms / ms_instrumented}
        gram_tops = gram_ops / (gram_ms * 1e-3) / 1e12 if gram_ms > 0 else None
        gram_traffic = ncu.get("gram_fp4_fused_c16" if fp4 else "gram_fused_c16") if (fused and c16) else None
        gram_peak = (4 if fp4 else 2) * bf16
        rl_gram = {"bound": "tensor",
                   "kernel": ("gram_tc_kernel (tcgen05 kind::mxf4 on E2M1 dosages, block scales 2^0, M128 x N224 x K64, "
                              "fp32 in TMEM holding exact integers%s)" if fp4 else
                              "gram_tc_kernel (tcgen05 kind::i8, M128 x N256 x K32, s32 in TMEM%s)")
                             % ("; epilogue also writes the scaled fp32 matrix" if fused else ""),
                   "achieved": gram_tops, "peak": gram_peak, "unit": "TOP/s",
                   "frac": gram_tops / gram_peak if gram_tops else None,
                   "traffic": gram_traffic["bytes_per_unit"] * P if gram_traffic else None,
                   "traffic_source": gram_traffic["source"] if gram_traffic else None,
                   "algorithmic_ops_per_launch": gram_ops / max(1, gram_launches),
                   "peak_source": ("4 x " + bf16_src + " (fp4 dense = 4 x bf16; nominal 9 000 TOP/s)") if fp4 else
                                  ("2 x " + bf16_src + " (int8 dense = 2 x bf16; nominal 4 500 TOP/s)"),
                   "launches": int(gram_launches), "avg_launch_ms": gram_ms / max(1, gram_launches),
                   "share_of_step": gram_ms / ms_instrumented}
        ranked = sorted([("gram", rl_gram, gram_ms), ("solve", rl_solve, solve_ms), ("cholesky_update", rl_update, upd_ms)],
                        key=lambda t: -t[2])
        dominant = ranked[0][1]
        line = {
            "metric": METRIC, "value": evals / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": ("%s Gram (exact integers) + tf32/f16 Cholesky preconditioner + f64 refinement/solve"
                      % ("e2m1" if fp4 else "s8") if precision == "mixed"
                      else "%s Gram (exact integers) + f64 Cholesky/solve" % ("e2m1" if fp4 else "s8")), "data": "synthetic",
            "config": {"workload": args.workload, "animals": n, "markers": m, "k": k, "pop_per_gpu": P, "folds": folds,
                       "h2": H2, "n_train": int(n_t), "n_valid": int(n_v), "individuals_per_wave": wave, "precision": precision,
                       "genotype_storage": args.storage, "cross_product_storage": "int16" if c16 else "int32",
                       "scaling_fused_into_gram": fused, "gram_operands": "e2m1 (fp4)" if fp4 else "int8", "genotype_bytes_resident": eng.resident_genotype_bytes(),
                       "l2": "inputs larger than L2 (each step streams >20 GB of per-genome panels and matrices)",
                       "parallelism": "replicated genotypes, population sharded, NCCL all-gather of fitness"},
            "e2e": {"value": evals / (ms_e2e * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(batches[0][0].nbytes + batches[0][1].nbytes),
                    "d2h_bytes_per_step": int(P * len(slots) * 8), "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": dominant,
            "roofline_" + ranked[1][0]: ranked[1][1],
            "roofline_" + ranked[2][0]: ranked[2][1],
            "stage_ms_per_step": {s: v[0] / args.steps for s, v in stage.items()},
            "ms_per_step_instrumented": ms_instrumented / args.steps,
            "cpu_baseline": cpu_baseline,
            "parity": parity,
        }
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
