"""Multi-rank logic.  CPU: world_size-2 gloo tests of the host side
(seed sharding, final min-loc gather).  GPU: a 2-rank sharded swarm on one device over
gloo must equal the single-process run bit for bit (draws are keyed by global rows)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stochopy_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _spawn(fn, world, *args):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_entry, args=(fn, r, world, port, q) + args) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get() for _ in range(world)]
    for p in procs:
        p.join(60)
    errs = [o for o in out if isinstance(o, str)]
    assert not errs, errs
    return sorted(out, key=lambda o: o[0])


def _entry(fn, rank, world, port, q, *args):
    try:
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        q.put((rank, fn(rank, world, *args)))
        dist.destroy_process_group()
    except Exception as e:  # pragma: no cover
        import traceback

        q.put(f"rank {rank}: {e}\n{traceback.format_exc()}")


def test_shard_range_covers_everything():
    for total in (1, 7, 8, 64, 65536, 100003):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert parallel.shard_seeds(range(5), 1, 2) == [3, 4] and parallel.shard_seeds(range(5), 0, 2) == [0, 1, 2]


def _fake_run(rank, world, seeds):
    # objective value and solution are a pure function of the seed: every rank must agree on the winner
    def runner(seed):
        rs = np.random.RandomState(seed)
        return dict(fun=float(rs.uniform(0, 10)), x=rs.uniform(-1, 1, 4))

    r = parallel.minimize_seeds(None, [[-1, 1]] * 4, seeds, runner=runner)
    return r["fun"], r["x"].tolist(), r["funs"].tolist(), len(r["local"])


@pytest.mark.parametrize("nseeds", [1, 5, 8])
def test_minimize_seeds_gloo_world2(nseeds):
    seeds = list(range(100, 100 + nseeds))
    out = _spawn(_fake_run, 2, seeds)
    want = [np.random.RandomState(s).uniform(0, 10) for s in seeds]
    for rank, (fun, x, funs, nlocal) in out:
        assert np.allclose(funs, want) and np.isclose(fun, min(want))
        rs = np.random.RandomState(seeds[int(np.argmin(want))])
        rs.uniform(0, 10)
        assert np.allclose(x, rs.uniform(-1, 1, 4))
    assert sum(o[1][3] for o in out) == nseeds


def _split_eval(rank, world, nrows):
    rows = np.random.RandomState(4).uniform(-1, 1, (nrows, 3))
    calls = []

    def fun(x, shift):
        calls.append(1)
        return float(np.sum(x * x) + shift)

    f = parallel.evaluate_split(fun, (0.5,), rows)
    return f.tolist(), len(calls)


@pytest.mark.parametrize("nrows", [1, 7, 10])
def test_evaluate_split_gloo_world2(nrows):
    """8f-4: rank-strided host evaluation + all-reduce (the reference's mpi backend)."""
    rows = np.random.RandomState(4).uniform(-1, 1, (nrows, 3))
    want = (rows * rows).sum(axis=1) + 0.5
    out = _spawn(_split_eval, 2, nrows)
    for rank, (f, ncalls) in out:
        assert np.allclose(f, want) and ncalls == len(range(rank, nrows, 2))
    assert np.allclose(parallel.evaluate_split(lambda x: float(x.sum()), (), rows), rows.sum(axis=1))  # no group: serial


def _mpi_backend(rank, world, method):
    import stochopy_b200 as sb

    torch.cuda.set_device(0)
    calls = []

    def fun(x):
        calls.append(1)
        return float(np.sum((x - 0.3) ** 2))

    r = sb.optimize.minimize(fun, [[-2.0, 2.0]] * 4, method=method,
                             options=dict(maxiter=12, popsize=9, seed=2, backend="mpi", **({"updating": "deferred"} if method in ("de", "pso") else {})))
    return r.x.tolist(), r.fun, r.nit, len(calls)


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["de", "pso", "cmaes"])
def test_mpi_backend_splits_host_evaluations(method):
    import stochopy_b200 as sb

    def fun(x):
        return float(np.sum((x - 0.3) ** 2))

    one = sb.optimize.minimize(fun, [[-2.0, 2.0]] * 4, method=method,
                               options=dict(maxiter=12, popsize=9, seed=2, **({"updating": "deferred"} if method in ("de", "pso") else {})))
    out = _spawn(_mpi_backend, 2, method)
    for rank, (x, f, nit, ncalls) in out:
        assert np.array_equal(np.array(x), one.x) and f == one.fun and nit == one.nit
    assert sum(o[1][3] for o in out) == one.nfev and out[0][1][3] > out[1][1][3] > 0  # 5 + 4 rows per generation


def _sharded(rank, world, opts):
    import stochopy_b200 as sb

    torch.cuda.set_device(0)
    r = parallel.cpso_sharded(sb.factory.styblinski_tang, [[-5.12, 5.12]] * 10, **opts)
    return r.x.tolist(), r.fun, r.nit, r.status


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", ["peer", "nccl"])
@pytest.mark.parametrize("opts", [dict(competitivity=None, constraints="Shrink"), dict(competitivity=1.0),
                                  dict(competitivity=1.2, dtype="float32")])
def test_sharded_swarm_equals_single_process(opts, exchange):
    """exchange="peer": gbest / radius / pbestfit travel through CUDA-IPC mapped mailboxes inside
    the kernels (csrc/peer.cuh); here both ranks sit on one GPU, so the spin-waits resolve by
    context time-slicing.  exchange="nccl": host-driven collectives (gloo in this test)."""
    import stochopy_b200 as sb

    o = dict(opts, maxiter=60, popsize=101, seed=11, exchange=exchange)
    single = {k: v for k, v in o.items() if k != "exchange"}
    one = sb.optimize.minimize(sb.factory.styblinski_tang, [[-5.12, 5.12]] * 10, method="cpso",
                               options=dict(single, updating="deferred"))
    solo = parallel.cpso_sharded(sb.factory.styblinski_tang, [[-5.12, 5.12]] * 10, **o)
    assert np.array_equal(solo.x, one.x) and solo.fun == one.fun and (solo.nit, solo.status) == (one.nit, one.status)
    for rank, (x, fun, nit, status) in _spawn(_sharded, 2, o):  # two ranks, odd split 51 + 50
        assert np.array_equal(np.array(x), one.x) and fun == one.fun and (nit, status) == (one.nit, one.status)


# ---- the mpi-backend analogue must evaluate rank 0's population on every rank -------------
def _split_rows(rank, world):
    rs = np.random.RandomState(100 + rank)  # every rank holds DIFFERENT rows, like unseeded optimisers would
    rows = rs.uniform(-1, 1, (11, 3))
    f = parallel.evaluate_split(lambda x: float(np.sum(x * x)), (), rows)
    seed = parallel.shared_seed(None)
    return f.tolist(), seed, parallel.shared_seed(7)


def test_evaluate_split_broadcasts_rank0_rows_and_shares_a_seed():
    """reference _common.py:58-72: Bcast(x, root=0), strided evaluation, Allreduce -- all ranks end up with
    the fitness of rank 0's rows; with seed=None the ranks agree on rank 0's seed."""
    out = _spawn(_split_rows, 2)
    want = np.sum(np.random.RandomState(100).uniform(-1, 1, (11, 3)) ** 2, axis=1)
    for _, (f, seed, given) in out:
        assert np.allclose(f, want, rtol=1e-15)
        assert given == 7
    assert out[0][1][1] == out[1][1][1] and out[0][1][1] is not None


# ---- real multi-GPU run (driver-visible): sharded swarm == single GPU bitwise, seeds agree ---------------
@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_sharded_swarm_equals_single_gpu(world):
    """tests/multi_gpu_check.py under torchrun on `world` real GPUs (NCCL + NVLink peer mailboxes); skipped
    when the box has fewer devices."""
    import subprocess
    import sys

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, box has {torch.cuda.device_count()}")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(here, "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "multi_gpu_check ok" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])


@pytest.mark.gpu
@pytest.mark.parametrize("method,n,opts", [("vdcma", 300, dict(popsize=600, maxiter=12, dtype="float32")),
                                           ("de", 128, dict(popsize=900, maxiter=15, dtype="float32", updating="deferred")),
                                           ("cpso", 10, dict(popsize=500, maxiter=30, competitivity=1.0, updating="deferred"))])
def test_minimize_seeds_concurrent_streams_equal_sequential(method, n, opts):
    """A rank's seeds on several host threads with their own CUDA streams (minimize_seeds(concurrent=k)):
    every run has its own buffers, control block and Philox key, so the results are those of the sequential
    loop, bit for bit."""
    import stochopy_b200 as sb

    b = [[-5.12, 5.12]] * n
    seeds = list(range(6))
    a = parallel.minimize_seeds(sb.factory.rastrigin, b, seeds, method=method, options=opts)
    c = parallel.minimize_seeds(sb.factory.rastrigin, b, seeds, method=method, options=opts, concurrent=3)
    assert np.array_equal(a["funs"], c["funs"]) and a["fun"] == c["fun"] and np.array_equal(a["x"], c["x"])
    for ra, rc in zip(a["local"], c["local"]):
        assert (ra.nit, ra.status, ra.nfev) == (rc.nit, rc.status, rc.nfev) and np.array_equal(ra.x, rc.x)
