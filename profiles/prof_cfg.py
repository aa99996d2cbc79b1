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
            print(f"eigh N={N} warm={warm} pert={pert}: {dt*1e3:.3f} ms, sweeps={int(sw.item())}, per round {dt*1e6/max(1,int(sw.item()))/(N-1+N%2):.2f} us", flush=True)
