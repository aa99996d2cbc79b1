"""Run under torchrun on >= 2 GPUs (NCCL): the row-sharded swarm must reproduce the
single-GPU run bit for bit, and seed sharding must agree on the winner.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    import stochopy_b200 as sb
    from stochopy_b200 import parallel

    b = [[-5.12, 5.12]] * 64
    for opts in (dict(competitivity=None, constraints="Shrink"), dict(competitivity=1.0), dict(competitivity=1.0, dtype="float32")):
        o = dict(opts, maxiter=80, popsize=4099, seed=5)
        one = sb.optimize.minimize(sb.factory.styblinski_tang, b, method="cpso", options=dict(o, updating="deferred"))
        for exchange in ("peer", "nccl"):  # in-kernel NVLink mailboxes / host-driven NCCL all-gather
            r = parallel.cpso_sharded(sb.factory.styblinski_tang, b, exchange=exchange, **o)
            assert np.array_equal(r.x, one.x) and r.fun == one.fun and (r.nit, r.status) == (one.nit, one.status), (rank, opts, exchange)
    seeds = list(range(2 * world + 1))
    res = parallel.minimize_seeds(sb.factory.rastrigin, [[-5.12, 5.12]] * 16, seeds, method="de",
                                  options=dict(maxiter=60, popsize=256, dtype="float64", updating="deferred"))
    ref = [sb.optimize.minimize(sb.factory.rastrigin, [[-5.12, 5.12]] * 16, method="de",
                                options=dict(maxiter=60, popsize=256, seed=s, updating="deferred")).fun for s in seeds]
    assert np.array_equal(res["funs"], np.array(ref)) and res["fun"] == min(ref), rank
    dist.barrier()
    if rank == 0:
        print(f"multi_gpu_check ok on {world} GPUs: sharded swarm == single GPU (bitwise), seeds agree", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
