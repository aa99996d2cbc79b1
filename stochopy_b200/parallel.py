"""Multi-GPU use of the engine, one process per GPU.

Two ways the population path spreads over GPUs:

* ``minimize_seeds``: independent restarts / seeds.  Each rank runs its share of the
  seeds on its own GPU with no data-path collective; one all-gather of
  ``(fun, x[N])`` per seed at the end, then a local argmin (NCCL has no MINLOC).
  This is what ``bench.py --gpus N`` scales.
* ``cpso_sharded``: ONE swarm row-sharded over the ranks (reference analogue: the mpi
  backend that splits a population's evaluation over ranks, ``_common.py:53-72``).
  Per generation the ranks exchange their local best ``[fit, x]`` (N+1 scalars,
  latency bound) and every rank reduces them with ``sp_gbest_reduce``; the
  competitive restart needs a max-reduce of the swarm radius and the global rank of
  ``pbestfit``.  Draws are Philox keyed by the *global* row index, so a sharded run is
  identical, bit for bit, to the single-GPU run of the same seed.

``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests of the host logic) is the
plumbing; the collectives carry device tensors.
"""
import ctypes as C

import os

import numpy as np
import torch
import torch.distributed as dist

__all__ = ["shard_range", "shard_seeds", "reduce_best", "minimize_seeds", "cpso_sharded", "PeerMailboxes", "evaluate_split",
           "shared_seed"]


def _all_gather_flat(out, part, group):
    """out (world * k) <- concatenation of every rank's `part` (k); list form on gloo."""
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(out, part, group=group)
    else:
        world = dist.get_world_size(group)
        chunks = list(out.view(world, -1).unbind(0))
        dist.all_gather(chunks, part, group=group)


class PeerMailboxes:
    """One device mailbox per rank, every mailbox mapped into every process (CUDA IPC;
    NVLink peer memory between GPUs).  ``torch.distributed`` only carries the 64-byte
    handles once; afterwards the kernels store into the peers' mailboxes themselves
    (csrc/peer.cuh).  ``table`` is the device array of the `world` mailbox pointers as
    addressed from this process; ``own`` is this rank's mailbox."""

    def __init__(self, nbytes, group=None, device=None):
        from . import _lib as L

        self._L = L
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        own, handle = C.c_void_p(), C.create_string_buffer(L.PEER_HANDLE_BYTES)
        L.call("sp_peer_alloc", int(nbytes), C.byref(own), handle)
        self.own = own.value
        handles = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(handles, handle.raw, group=group)
        else:
            handles[0] = handle.raw
        self.ptrs, self._opened = [], []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.ptrs.append(self.own)
                continue
            q = C.c_void_p()
            L.call("sp_peer_open", C.create_string_buffer(h, L.PEER_HANDLE_BYTES), C.byref(q))
            self.ptrs.append(q.value)
            self._opened.append(q.value)
        device = device or torch.device("cuda", torch.cuda.current_device())
        self.table = torch.tensor(self.ptrs, dtype=torch.int64).to(device)
        if self.world > 1:
            dist.barrier(group=group)  # every mailbox is mapped everywhere before the first store

    def close(self):
        """Collective: unmap the peers' mailboxes, then free the own one."""
        if self.own is None:
            return
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier(group=self.group)  # nobody stores into a mailbox that is about to go away
        for q in self._opened:
            self._L.call("sp_peer_close", C.c_void_p(q))
        if self.world > 1:
            dist.barrier(group=self.group)
        self._L.call("sp_peer_free", C.c_void_p(self.own))
        self.own, self._opened = None, []


def _group_root(group):
    """Global rank of the group's rank 0 (the broadcast source)."""
    return dist.get_global_rank(group, 0) if group is not None else 0


def evaluate_split(fun, args, rows, group=None):
    """The reference's ``backend="mpi"`` evaluation (``_common.py:58-72``) over
    ``torch.distributed`` -- for objectives that are expensive *on the host* (SURVEY.md 8f-4):
    rank 0's population is broadcast (the reference's ``Bcast(x, root=0)``), every rank
    evaluates the rows ``rank::size`` and an Allreduce(SUM) completes the fitness vector, so
    all ranks get the fitness of rank 0's rows whatever their own random streams did.  (The
    front-ends additionally share rank 0's seed when none was given -- ``shared_seed`` -- so
    the ranks hold the same population in the first place.)  Without an initialised process
    group this is the serial loop."""
    rows = np.ascontiguousarray(rows, dtype=np.float64)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return np.array([fun(r, *args) for r in rows], dtype=np.float64)
    rank = dist.get_rank(group)
    on_gpu = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
    x = torch.from_numpy(rows).to(dev)
    dist.broadcast(x, src=_group_root(group), group=group)
    rows = x.cpu().numpy()
    f = np.zeros(len(rows), dtype=np.float64)
    f[rank::world] = [fun(r, *args) for r in rows[rank::world]]
    t = torch.from_numpy(f).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


def shared_seed(seed, group=None):
    """``backend="mpi"`` with ``seed=None``: every rank would draw its own random seed and hold a
    different population.  Rank 0 draws one and broadcasts it."""
    if seed is not None or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return seed
    on_gpu = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
    t = torch.from_numpy(np.random.SeedSequence().generate_state(1, dtype=np.uint32).astype(np.int64)).to(dev)
    dist.broadcast(t, src=_group_root(group), group=group)
    return int(t.cpu().item())


_SEED_STREAMS = {}  # (device, k) -> the k worker streams of minimize_seeds(concurrent=k)


def shard_range(total, rank, world):
    """Contiguous, balanced [start, stop) of `total` items for `rank` (first ranks get the extras)."""
    q, r = divmod(int(total), int(world))
    start = rank * q + min(rank, r)
    return start, start + q + (1 if rank < r else 0)


def shard_seeds(seeds, rank, world):
    """Seeds handled by `rank`: contiguous blocks in seed order."""
    seeds = list(seeds)
    a, b = shard_range(len(seeds), rank, world)
    return seeds[a:b]


def reduce_best(funs, xs, group=None, device=None):
    """All-gather per-seed results and pick the best on every rank.

    funs: (k,) local objective values, xs: (k, N) local solutions (k may differ by rank).
    Returns (best_fun, best_x, all_funs) with ties resolved to the lowest global seed
    position, like np.argmin."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    funs = np.atleast_1d(np.asarray(funs, dtype=np.float64))
    xs = np.atleast_2d(np.asarray(xs, dtype=np.float64))
    if world == 1:
        b = int(np.argmin(funs))
        return float(funs[b]), xs[b].copy(), funs.copy()
    device = device or (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else "cpu")
    n = xs.shape[1]
    counts = torch.zeros(world, dtype=torch.int64, device=device)
    counts[rank] = len(funs)
    dist.all_reduce(counts, group=group)
    kmax = int(counts.max().item())
    rec = torch.full((kmax, n + 1), float("inf"), dtype=torch.float64, device=device)
    rec[: len(funs), 0] = torch.from_numpy(funs).to(device)
    rec[: len(funs), 1:] = torch.from_numpy(xs).to(device)
    out = torch.empty((world, kmax, n + 1), dtype=torch.float64, device=device)
    _all_gather_flat(out.view(-1), rec.view(-1), group)
    out = out.cpu().numpy()
    rows = np.concatenate([out[r, : int(counts[r].item())] for r in range(world)], axis=0)
    b = int(np.argmin(rows[:, 0]))
    return float(rows[b, 0]), rows[b, 1:].copy(), rows[:, 0].copy()


def minimize_seeds(fun, bounds, seeds, method="de", options=None, group=None, runner=None, concurrent=1):
    """Run one optimisation per seed, seeds sharded over the ranks; returns the best
    OptimizeResult-like dict plus every seed's final value (identical on all ranks).

    concurrent > 1: this rank's seeds run on that many host threads, each on its own CUDA stream, so the
    latency-bound stages of one run (ranking, single-CTA updates, status reads) overlap the wide kernels of the
    others (seeds batched per GPU).  Results do not depend on it (every run has its own buffers,
    control block and Philox key)."""
    from .optimize import minimize

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    runner = runner or (lambda seed: minimize(fun, bounds, method=method, options=dict(options or {}, seed=seed)))
    local_seeds = shard_seeds(seeds, rank, world)
    if concurrent > 1 and len(local_seeds) > 1 and torch.cuda.is_available():
        from concurrent.futures import ThreadPoolExecutor

        dev = torch.cuda.current_device()
        k = min(int(concurrent), len(local_seeds))
        # one stream per worker, kept for the life of the process: PyTorch's caching allocator pools blocks per
        # stream, so a fresh stream per run would pay a cudaMalloc (and its device-wide sync) for every buffer
        streams = _SEED_STREAMS.setdefault((dev, k), [torch.cuda.Stream(device=dev) for _ in range(k)])
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(dev))

        def worker(w):
            torch.cuda.set_device(dev)  # a new host thread starts on device 0
            st = streams[w]
            st.wait_event(ready)
            out = []
            with torch.cuda.stream(st):
                for j in range(w, len(local_seeds), k):
                    out.append((j, runner(local_seeds[j])))
                st.synchronize()
            return out

        with ThreadPoolExecutor(max_workers=k) as pool:
            done = [r for part in pool.map(worker, range(k)) for r in part]
        mine = [r for _, r in sorted(done, key=lambda t: t[0])]
    else:
        mine = [runner(s) for s in local_seeds]
    n = len(bounds)
    funs = [r["fun"] for r in mine]
    xs = [r["x"] for r in mine] if mine else np.empty((0, n))
    best_fun, best_x, all_funs = reduce_best(funs, xs, group)
    return dict(x=best_x, fun=best_fun, funs=all_funs, seeds=list(seeds), local=mine)


def cpso_sharded(fun, bounds, maxiter=100, popsize=10, inertia=0.7298, cognitivity=1.49618, sociability=1.49618,
                 competitivity=1.0, seed=None, xtol=1.0e-8, ftol=1.0e-8, constraints=None, dtype="float64",
                 group=None, exchange="peer"):
    """One (C)PSO swarm of `popsize` particles row-sharded over the process group.

    Same algorithm, options and result as ``optimize.cpso`` (reference
    ``cpso/_cpso.py:182-321``, synchronous); device objectives only; every rank returns
    the same result.  With world size 1 it is exactly ``optimize.cpso``'s Philox path.

    exchange="peer" (default): the per-generation gbest / radius / pbestfit exchanges run
    inside the kernels over peer-mapped mailboxes (``PeerMailboxes``, ``sp_pso_run_sharded``):
    generations are enqueued in chunks with no collective call and no host round trip in
    between.  exchange="nccl": one all-gather + ``sp_gbest_reduce`` per generation driven
    from the host (the baseline; also what the gloo CPU tests of the host logic use)."""
    from . import _lib as L
    from .optimize._common import Engine, device_objective, fresh_seed, messages
    from .optimize._helpers import OptimizeResult

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    obj = device_objective(fun, ())
    if obj is None:
        raise ValueError("cpso_sharded needs one of the factory objectives")
    cons = {None: L.CONS_NONE, "Shrink": L.CONS_SHRINK}[constraints]
    if seed is None:
        raise ValueError("a sharded swarm needs an explicit seed (all ranks must draw the same stream)")
    if exchange not in {"peer", "nccl"}:
        raise ValueError()
    mode = exchange

    eng = Engine(dtype)
    bounds = np.asarray(bounds, dtype=np.float64)
    N, Ptot = len(bounds), int(popsize)
    row0, row1 = shard_range(Ptot, rank, world)
    P = row1 - row0
    if P < 1:
        raise ValueError("more ranks than particles")
    lower, upper = bounds[:, 0].copy(), bounds[:, 1].copy()
    restart = bool(competitivity)
    ld = eng.ld(N)
    X, V, pbest = eng.rows(P, N), eng.rows(P, N), eng.rows(P, N)
    pbestfit, pfit = eng.empty(P), eng.empty(P)
    gbest = eng.zeros(ld)
    d_lower, d_upper = eng.upload_vec(lower, ld), eng.upload_vec(upper, ld)
    ctrl, scratch = eng.new_ctrl()
    rec_ld = ld + 2  # [fit, x_0..x_{N-1}] padded
    xch = eng.zeros(rec_ld)
    recs = eng.zeros(world, rec_ld)
    rank_all = eng.zeros(Ptot, dtype=torch.int32)
    fit_all = eng.zeros(Ptot)

    st = L.PsoState()
    st.dtype, st.objective, st.constraint = eng.sp_dt, obj, cons
    st.P, st.N, st.maxiter, st.ld = P, N, int(maxiter), ld
    st.w, st.c1, st.c2 = float(inertia), float(cognitivity), float(sociability)
    st.xtol, st.ftol = float(xtol), float(ftol)
    st.gamma = float(competitivity) if restart else -1.0
    st.delta = float(np.log(1.0 + 0.003 * Ptot) / np.max((0.2, np.log(0.01 * maxiter)))) if restart else 0.0
    st.seed = fresh_seed(seed)
    st.X, st.V, st.pbest = X.data_ptr(), V.data_ptr(), pbest.data_ptr()
    st.pbestfit, st.pfit, st.gbest = pbestfit.data_ptr(), pfit.data_ptr(), gbest.data_ptr()
    st.lower, st.upper = d_lower.data_ptr(), d_upper.data_ptr()
    st.ctrl, st.scratch = ctrl.data_ptr(), scratch.data_ptr()
    st.row0, st.P_total, st.xch, st.shard = row0, Ptot, xch.data_ptr(), 1

    def exchange_best(it):
        """all-gather the local bests, reduce on every rank (status from the global best)."""
        if world > 1:
            _all_gather_flat(recs.view(-1), xch, group)
        else:
            recs[0].copy_(xch)
        L.call("sp_gbest_reduce", eng.sp_dt, recs.data_ptr(), world, N, rec_ld, gbest.data_ptr(), ctrl.data_ptr(), it,
               int(maxiter), float(xtol), float(ftol), eng.stream)

    L.call("sp_lhs_init_shard", eng.sp_dt, X.data_ptr(), P, N, ld, st.lower, st.upper, st.seed, Ptot, row0, eng.stream)
    pbest.copy_(X)
    L.call("sp_eval", obj, eng.sp_dt, X.data_ptr(), P, N, ld, None, None, pbestfit.data_ptr(), eng.stream)
    pfit.copy_(pbestfit)
    # initial best: local argmin -> record -> exchange (no status test: it = -1)
    b = int(torch.argmin(pbestfit).item())
    xch[0] = pbestfit[b]
    xch[1:1 + N] = X[b, :N]
    exchange_best(-1)

    ctrl64 = ctrl.view(torch.float64)  # aux[0] (local max squared radius) sits at byte 40
    it = 1
    c = eng.read_ctrl(ctrl)
    # the generation loop timed on the device (CUDA events on the launching stream): the IPC mailbox set-up in
    # front of it is a fixed cost of tens to hundreds of milliseconds that a wall clock would mix into the rate
    ev_loop = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    it_loop0, started = 1, False
    if mode == "peer":
        box = PeerMailboxes(L.load().sp_peer_bytes(eng.sp_dt, world, ld, Ptot), group, eng.device)
        st.shard, st.world, st.rank = 2, world, rank
        st.mailbox, st.peers = box.own, box.table.data_ptr()
        try:
            eager_left, last_restart = 0, -(1 << 30)
            if os.environ.get("SP_SHARD_EAGER"):  # profiling switch: gated in-chunk restarts from the first generation on
                eager_left = 1 << 30
            if world > 1:
                dist.barrier(group=group)  # every rank has mapped every mailbox and is about to enqueue
            ev_loop[0].record()
            started = True
            while c.status == L.SP_RUNNING:
                n = min(32 if it < 64 else 128, max(int(maxiter), 2) - it)
                if restart and eager_left <= 0:
                    # restart out of the common path (see optimize/_cpso.py): every rank parks at the same
                    # generation (the decision comes from the exchanged radius) and resumes collectively
                    L.call("sp_pso_run_lazy", C.byref(st), it + 1, min(n, 32), eng.stream)
                    c = eng.read_ctrl(ctrl)
                    if c.status == L.SP_STATUS_RESTART_PENDING:
                        L.call("sp_cpso_restart_resume", C.byref(st), c.nit, rank_all.data_ptr(), eng.stream)
                        if c.nit - last_restart < 16:
                            eager_left = 64
                        last_restart = c.nit
                        c.status = L.SP_RUNNING
                    it = c.nit
                    continue
                L.call("sp_pso_run_sharded", C.byref(st), it + 1, n, rank_all.data_ptr(), eng.stream)
                eager_left -= n
                c = eng.read_ctrl(ctrl)
                it = c.nit
        finally:
            box.close()
        if c.status == L.SP_STATUS_PEER_TIMEOUT:
            raise L.EngineError("cpso_sharded: a peer did not answer within the exchange timeout")
    if not started:
        ev_loop[0].record()
    while c.status == L.SP_RUNNING:
        it += 1
        L.call("sp_pso_generation", C.byref(st), it, eng.stream)
        exchange_best(it)
        c = eng.read_ctrl(ctrl)
        if c.status == L.SP_RUNNING and restart:  # _cpso.py:304-307, 405-426
            L.call("sp_cpso_radius", C.byref(st), it, eng.stream)
            if world > 1:
                dist.all_reduce(ctrl64[5:6], op=dist.ReduceOp.MAX, group=group)
            L.call("sp_cpso_decide", C.byref(st), it, eng.stream)
            if eng.read_ctrl(ctrl).flag > 0:
                if world > 1:
                    counts = [shard_range(Ptot, r, world) for r in range(world)]
                    if len({b - a for a, b in counts}) == 1:
                        _all_gather_flat(fit_all, pbestfit, group)
                    else:
                        parts = [eng.empty(b - a) for a, b in counts]
                        dist.all_gather(parts, pbestfit, group=group)
                        fit_all.copy_(torch.cat(parts))
                else:
                    fit_all.copy_(pbestfit)
                L.call("sp_fitness_rank", eng.sp_dt, fit_all.data_ptr(), Ptot, rank_all.data_ptr(), eng.stream)
                L.call("sp_cpso_restart_apply", C.byref(st), it, rank_all[row0:].data_ptr(), None, eng.stream)

    it = c.nit
    ev_loop[1].record()
    ev_loop[1].synchronize()
    return OptimizeResult(
        x=gbest[:N].to("cpu").numpy().astype(np.float64),
        success=c.status >= 0,
        status=int(c.status),
        message=messages[int(c.status)],
        fun=float(c.gfit),
        nfev=it * Ptot,
        nit=it,
        loop_ms=float(ev_loop[0].elapsed_time(ev_loop[1])),  # generations 2..nit on this rank (device time)
        loop_generations=int(it - it_loop0),
    )
