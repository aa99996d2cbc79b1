#!/usr/bin/env python
"""Throughput of the other BASELINE.json configs (C2-C5) through the public API.

Not the driver's headline (that is bench.py); this records evals/s = nfev / wall for
the remaining configurations so DESIGN.md / profiles/ can quote them.  Termination is
disabled (xtol=-1, ftol=-1e300) so exactly `maxiter` generations run.  Under torchrun
(N ranks) C3 runs ONE swarm sharded over the ranks and C5 runs independent seeds.

    python bench_configs.py [--quick]
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 bench_configs.py
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def timed(fn, repeat=3):
    best = None
    for _ in range(repeat):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if best is None or dt < best[0]:
            best = (dt, r)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    import stochopy_b200 as sb
    from stochopy_b200 import parallel

    off = dict(xtol=-1.0, ftol=-1.0e300)
    q = 4 if args.quick else 1
    out = []

    def report(name, nfev, dt, extra=None):
        if rank == 0:
            rec = dict(config=name, evals_per_s=nfev / dt, seconds=dt, nfev=nfev, n_gpus=world, **(extra or {}))
            out.append(rec)
            print(json.dumps(rec), flush=True)

    b128 = [[-5.12, 5.12]] * 128
    if world == 1:
        # warm-up (module load, allocator)
        sb.optimize.minimize(sb.factory.sphere, b128, method="de", options=dict(maxiter=3, popsize=1024, seed=0, dtype="float32"))
        # C2: DE on Rastrigin ndim=128, popsize=65536, maxiter=1000, fp32
        for strat in ("best1bin", "rand1bin"):
            o = dict(maxiter=1000 // q, popsize=65536, seed=0, dtype="float32", strategy=strat, updating="deferred", **off)
            dt, r = timed(lambda: sb.optimize.minimize(sb.factory.rastrigin, b128, method="de", options=o))
            report(f"C2 de/{strat} rastrigin N=128 P=65536 fp32", r.nfev, dt, dict(nit=r.nit, fun=r.fun))
        # headline objective for reference
        o = dict(maxiter=1000 // q, popsize=65536, seed=0, dtype="float32", updating="deferred", **off)
        dt, r = timed(lambda: sb.optimize.minimize(sb.factory.rosenbrock, b128, method="de", options=o))
        report("headline de/best1bin rosenbrock N=128 P=65536 fp32", r.nfev, dt, dict(nit=r.nit, fun=r.fun))
        # C3 single GPU: PSO / CPSO on Styblinski-Tang ndim=64, popsize=32768
        b64 = [[-5.12, 5.12]] * 64
        for method, extra in (("pso", {}), ("cpso", dict(competitivity=1.0))):
            for dt_name in ("float32", "float64"):
                o = dict(maxiter=1000 // q, popsize=32768, seed=0, dtype=dt_name, updating="deferred", **extra, **off)
                dt, r = timed(lambda: sb.optimize.minimize(sb.factory.styblinski_tang, b64, method=method, options=o))
                report(f"C3 {method} styblinski_tang N=64 P=32768 {dt_name}", r.nfev, dt, dict(nit=r.nit, fun=r.fun))
        # C4: CMA-ES on Rosenbrock ndim=256, popsize=4096, fp64
        o = dict(maxiter=40 // q, popsize=4096, seed=0, **off)
        dt, r = timed(lambda: sb.optimize.minimize(sb.factory.rosenbrock, [[-5.12, 5.12]] * 256, method="cmaes", options=o), 2)
        report("C4 cmaes rosenbrock N=256 P=4096 fp64", r.nfev, dt, dict(nit=r.nit, fun=r.fun, status=r.status))
        o = dict(maxiter=100 // q, popsize=16384, seed=0, **off)
        dt, r = timed(lambda: sb.optimize.minimize(sb.factory.rosenbrock, b128, method="cmaes", options=o), 2)
        report("cmaes rosenbrock N=128 P=16384 fp64", r.nfev, dt, dict(nit=r.nit, fun=r.fun, status=r.status))
        # C5: VD-CMA on Ackley ndim=1024, popsize=16384
        for dt_name in ("float32", "float64"):
            o = dict(maxiter=100 // q, popsize=16384, seed=0, dtype=dt_name, **off)
            dt, r = timed(lambda: sb.optimize.minimize(sb.factory.ackley, [[-5.12, 5.12]] * 1024, method="vdcma", options=o), 2)
            report(f"C5 vdcma ackley N=1024 P=16384 {dt_name}", r.nfev, dt, dict(nit=r.nit, fun=r.fun, status=r.status))
    else:
        b64 = [[-5.12, 5.12]] * 64
        # C3: ONE swarm row-sharded over the GPUs; "peer" = exchange fused into the kernels over
        # NVLink peer memory, "nccl" = one host-driven all-gather per generation (the baseline)
        for method, comp in (("cpso", 1.0), ("pso", None)):
            for exchange in ("peer", "nccl"):
                o = dict(maxiter=300 // q, popsize=32768, seed=0, dtype="float32", competitivity=comp,
                         exchange=exchange, **off)
                parallel.cpso_sharded(sb.factory.styblinski_tang, b64, **dict(o, maxiter=5))
                dist.barrier()
                dt, r = timed(lambda: parallel.cpso_sharded(sb.factory.styblinski_tang, b64, **o), 2)
                report(f"C3 {method} ONE swarm sharded over {world} GPUs ({exchange} exchange), styblinski_tang N=64 "
                       f"P=32768 fp32", r.nfev, dt, dict(nit=r.nit, fun=r.fun))
        # warm-up of the vdcma kernels (module load, allocator) before the timed seeds
        parallel.minimize_seeds(sb.factory.ackley, [[-5.12, 5.12]] * 1024, [0] * world, method="vdcma",
                                options=dict(maxiter=3, popsize=16384, dtype="float32", **off))
        # C5: 8 seeds per GPU, independent
        seeds = list(range(8 * world if not args.quick else world))
        o = dict(maxiter=50 // q, popsize=16384, dtype="float32", **off)
        dist.barrier()
        t0 = time.perf_counter()
        r = parallel.minimize_seeds(sb.factory.ackley, [[-5.12, 5.12]] * 1024, seeds, method="vdcma", options=o)
        torch.cuda.synchronize()
        dist.barrier()
        dt = time.perf_counter() - t0
        report(f"C5 vdcma ackley N=1024 P=16384 fp32, {len(seeds)} seeds over {world} GPUs",
               len(seeds) * o["maxiter"] * 16384, dt, dict(best_fun=r["fun"]))
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
