import sys, os
sys.path.insert(0, os.getcwd())
import stochopy_b200 as sb
which = sys.argv[1]
off = dict(xtol=-1.0, ftol=-1.0e300)
if which == "vd":
    sb.optimize.minimize(sb.factory.ackley, [[-5.12, 5.12]] * 1024, method="vdcma", options=dict(maxiter=4, popsize=16384, seed=0, dtype="float32", **off))
elif which == "cma":
    sb.optimize.minimize(sb.factory.rosenbrock, [[-5.12, 5.12]] * 256, method="cmaes", options=dict(maxiter=4, popsize=4096, seed=0, **off))
elif which == "pso":
    sb.optimize.minimize(sb.factory.styblinski_tang, [[-5.12, 5.12]] * 64, method="pso", options=dict(maxiter=6, popsize=32768, seed=0, dtype="float32", updating="deferred", **off))
elif which == "cpso":
    sb.optimize.minimize(sb.factory.styblinski_tang, [[-5.12, 5.12]] * 64, method="cpso", options=dict(maxiter=6, popsize=32768, seed=0, dtype="float32", updating="deferred", **off))
elif which == "vd64":
    sb.optimize.minimize(sb.factory.ackley, [[-5.12, 5.12]] * 1024, method="vdcma", options=dict(maxiter=4, popsize=16384, seed=0, **off))
if which == "cma_time":
    import time, torch
    for n, p, it in ((256, 4096, 30), (128, 16384, 30), (512, 2048, 10)):
        o = dict(maxiter=it, popsize=p, seed=0, **off)
        sb.optimize.minimize(sb.factory.rosenbrock, [[-5.12, 5.12]] * n, method="cmaes", options=dict(o, maxiter=3))
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = sb.optimize.minimize(sb.factory.rosenbrock, [[-5.12, 5.12]] * n, method="cmaes", options=o)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"cmaes N={n} P={p}: {dt / it * 1e3:.2f} ms/gen, {r.nfev / dt:.3e} evals/s", flush=True)
if which == "eigh_time":
    import time, torch, numpy as np, ctypes as C
    from stochopy_b200 import _lib as L
    from stochopy_b200.optimize._common import Engine
    eng = Engine("float64")
    for N in (64, 113, 128, 256, 512):
        rs = np.random.RandomState(N)
        A = rs.normal(0, 1, (N, N)); Cm = A @ A.T / N + np.diag(rs.uniform(0.01, 3.0, N))
        w, B = eng.zeros(N), eng.zeros(N, N)
        work = eng.zeros(int(L.load().sp_sym_eigh_work_scalars(N)))
        sw = eng.zeros(1, dtype=torch.int32)
        for warm, pert in ((0, 0.0), (1, 0.03), (1, 0.001)):
            Cp = Cm + pert * (lambda E: E @ E.T / N)(rs.normal(0, 1, (N, N)))
            dC = torch.from_numpy(Cp).to(eng.device)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            L.call("sp_sym_eigh", eng.sp_dt, dC.data_ptr(), N, w.data_ptr(), B.data_ptr(), work.data_ptr(), warm, sw.data_ptr(), eng.stream)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            nsw = int(sw.item())
            offb = (2 * N * N + N) * 8  # per-sweep largest |cos| (fp32 bit patterns) sit behind W and lambda in `work`
            worst = work.view(torch.uint8)[offb:offb + 4 * nsw].view(torch.float32).cpu().numpy()
            print(f"eigh N={N} warm={warm} pert={pert}: {dt*1e3:.3f} ms, sweeps={nsw}, per round {dt*1e6/max(1,nsw)/(N-1+N%2):.2f} us, worst |cos| per sweep {['%.1e' % v for v in worst]}", flush=True)

if which == "slopes":
    # per-generation device time in a real run (no profiler): slope between a short and a long run
    import time, torch
    def run(method, fun, n, p, it, **kw):
        o = dict(maxiter=it, popsize=p, seed=0, **off, **kw)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        sb.optimize.minimize(fun, [[-5.12, 5.12]] * n, method=method, options=o)
        torch.cuda.synchronize(); return time.perf_counter() - t0
    cases = [("vdcma", sb.factory.ackley, 1024, 16384, dict(dtype="float32")), ("vdcma", sb.factory.ackley, 1024, 16384, {}),
             ("cpso", sb.factory.styblinski_tang, 64, 32768, dict(dtype="float32", updating="deferred")),
             ("pso", sb.factory.styblinski_tang, 64, 32768, dict(dtype="float32", updating="deferred")),
             ("de", sb.factory.rastrigin, 128, 65536, dict(dtype="float32", updating="deferred")),
             ("cmaes", sb.factory.rosenbrock, 256, 4096, {}), ("cmaes", sb.factory.rosenbrock, 128, 16384, {})]
    for method, fun, n, p, kw in cases:
        a, b = (10, 40) if method == "cmaes" else (50, 450)
        run(method, fun, n, p, 5, **kw)
        ta = min(run(method, fun, n, p, a, **kw) for _ in range(2))
        tb = min(run(method, fun, n, p, b, **kw) for _ in range(2))
        per = (tb - ta) / (b - a)
        print(f"{method} {fun.__name__} N={n} P={p} {kw.get('dtype', 'float64')}: {per * 1e6:.1f} us/gen, {p / per:.3e} evals/s (fixed {1e3 * (ta - a * per):.2f} ms)", flush=True)
if which in ("shard_pso", "shard_cpso"):
    from stochopy_b200 import parallel
    parallel.cpso_sharded(sb.factory.styblinski_tang, [[-5.12, 5.12]] * 64, maxiter=6, popsize=32768, seed=0, dtype="float32",
                          competitivity=(None if which == "shard_pso" else 1.0), exchange="peer", **off)
if which == "eigh256":
    import torch, numpy as np, ctypes as C
    from stochopy_b200 import _lib as L
    from stochopy_b200.optimize._common import Engine
    eng = Engine("float64")
    N = 256
    rs = np.random.RandomState(N)
    A = rs.normal(0, 1, (N, N)); Cm = A @ A.T / N + np.diag(rs.uniform(0.01, 3.0, N))
    w, B = eng.zeros(N), eng.zeros(N, N)
    work = eng.zeros(int(L.load().sp_sym_eigh_work_scalars(N)))
    sw = eng.zeros(1, dtype=torch.int32)
    for warm, pert in ((0, 0.0), (1, 0.03)):
        Cp = Cm + pert * (lambda E: E @ E.T / N)(rs.normal(0, 1, (N, N)))
        dC = torch.from_numpy(Cp).to(eng.device)
        L.call("sp_sym_eigh", eng.sp_dt, dC.data_ptr(), N, w.data_ptr(), B.data_ptr(), work.data_ptr(), warm, sw.data_ptr(), eng.stream)
        torch.cuda.synchronize()
