"""Where a generation of the row-sharded CPSO goes (torchrun): the device-timed generation loop of
parallel.cpso_sharded with stages of the eager restart sequence switched off one after the other
(SP_SHARD_SKIP, timing only -- the results are then wrong), plain PSO (no restart) as the floor.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/prof_sharded_stages.py SKIP"""
import os
import sys

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
skip = os.environ.get("SP_SHARD_SKIP", "0")
for P in (32768, 262144):
    for comp in ((None, 1.0) if skip == "0" else (1.0,)):
        kw = dict(popsize=P, competitivity=comp, exchange="peer", seed=0, dtype="float32", **off)
        parallel.cpso_sharded(sb.factory.styblinski_tang, b64, maxiter=8, **kw)
        r = parallel.cpso_sharded(sb.factory.styblinski_tang, b64, maxiter=300, **kw)
        t = torch.tensor([r.loop_ms * 1e3 / r.loop_generations], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"skip={skip} {'cpso' if comp else 'pso '} P={P} over {world} GPUs: {t.item():.1f} us/gen", flush=True)
if world > 1:
    dist.destroy_process_group()
