"""Per-generation device time of ONE swarm row-sharded over the ranks (torchrun), by the slope
between a short and a long run, for both exchange modes:
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/prof_sharded.py"""
import os
import sys
import time

sys.path.insert(0, os.getcwd())
import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
import stochopy_b200 as sb
from stochopy_b200 import parallel

off = dict(xtol=-1.0, ftol=-1.0e300)
b64 = [[-5.12, 5.12]] * 64


def barrier():
    if world > 1:
        dist.barrier()


def run(it, **kw):
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    parallel.cpso_sharded(sb.factory.styblinski_tang, b64, maxiter=it, seed=0, dtype="float32", **off, **kw)
    torch.cuda.synchronize()
    barrier()
    return time.perf_counter() - t0


for P in (int(a) for a in (sys.argv[1:] or ["32768", "262144"])):
    for comp in (None, 1.0):
        for exchange in ("peer", "nccl"):
            kw = dict(popsize=P, competitivity=comp, exchange=exchange)
            run(5, **kw)
            a, b = 100, 1100
            ta = min(run(a, **kw) for _ in range(2))
            tb = min(run(b, **kw) for _ in range(2))
            per = (tb - ta) / (b - a)
            if rank == 0:
                print(f"{'cpso' if comp else 'pso'} P={P} N=64 fp32 over {world} GPUs, {exchange} exchange: {per * 1e6:.1f} us/gen, "
                      f"{P / per:.3e} evals/s (fixed {1e3 * (ta - a * per):.2f} ms)", flush=True)
if world > 1:
    dist.destroy_process_group()
