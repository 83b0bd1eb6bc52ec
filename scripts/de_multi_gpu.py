"""Device DE with the evaluation sharded over the GPUs of one box (keys replicated, offspring regenerated on every
rank, one all-reduce of P doubles per generation) against the same run on one GPU.
usage: torchrun --nproc-per-node N scripts/de_multi_gpu.py [pop] [generations] [n] [m] [k]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from tblup_b200 import GblupEngine, synth
from tblup_b200.de import DeviceDE, mutation_intensity

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
G = int(sys.argv[2]) if len(sys.argv) > 2 else 4
n = int(sys.argv[3]) if len(sys.argv) > 3 else 5000
m = int(sys.argv[4]) if len(sys.argv) > 4 else 50000
k = int(sys.argv[5]) if len(sys.argv) > 5 else 5001
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
x, y = synth.synth_dataset(n, m, seed=0)
tr, va, te = synth.split_indices(n, seed=0)
eng = GblupEngine(x, y, perm=np.concatenate([tr, va, te]), device=local)
eng.set_rowset(0, tr, va)
de = DeviceDE(eng, P, k, seed=1)
f0 = de.evaluate_distributed()
ts, takes = [], []
for g in range(1, G + 1):
    dist.barrier()
    t0 = time.perf_counter()
    takes.append(de.step_distributed(mutation_intensity(g, 0.5), 0.8, seed=g))
    ts.append(time.perf_counter() - t0)
fit = de.fitness()
# every rank must hold the same population
all_fit = [torch.empty(P, dtype=torch.float64, device="cuda") for _ in range(world)]
dist.all_gather(all_fit, torch.from_numpy(fit).cuda())
same = all(torch.equal(all_fit[0], t) or bool(torch.all((all_fit[0] == t) | (torch.isnan(all_fit[0]) & torch.isnan(t))))
           for t in all_fit)
if rank == 0:
    print("%d GPUs, pop %d, %d x %d, k=%d: sharded DE generation median %.1f ms -> %.0f individuals/s; ranks agree: %s"
          % (world, P, n, m, k, 1e3 * np.median(ts), P / np.median(ts), same))
    ref = GblupEngine(x, y, perm=np.concatenate([tr, va, te]), device=local)
    ref.set_rowset(0, tr, va)
    one = DeviceDE(ref, P, k, seed=1)
    g0 = one.evaluate()
    ok = np.array_equal(np.nan_to_num(g0, nan=-1), np.nan_to_num(f0, nan=-1))
    for g in range(1, G + 1):
        tk = one.step(mutation_intensity(g, 0.5), 0.8, seed=g)
        ok = ok and np.array_equal(tk, takes[g - 1])
    ok = ok and np.array_equal(np.nan_to_num(one.fitness(), nan=-1), np.nan_to_num(fit, nan=-1))
    print("identical to the single-GPU run (generation-0 fitness, every selection, final fitness):", ok)
    assert same and ok
dist.barrier()
dist.destroy_process_group()
