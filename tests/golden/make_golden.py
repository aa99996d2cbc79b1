"""Generate the golden fixtures in this directory from the REAL reference.

Run in the dev container only (the reference lives at /root/reference and does
not exist on the GPU box):

    python tests/golden/make_golden.py

Outputs (committed):
  reference_cases.json  the 24 (options, xref) cases of the reference's
                        tests/test_optimize.py with the literals from that file,
                        plus what the reference returns here (x, fun, nit, nfev,
                        status), plus the README run
  factory.json          tests/test_factory.py known answers + random rows
  steps.npz             inputs/outputs of the reference's *step functions*
                        (lhs, selection_sync, de_sync, mutation/Shrink, restart,
                        Penalize, converge, pvec_and_qvec, ngv_ngd) at sizes the
                        reference's tests never reach, with the numpy random
                        state recorded so the draws can be replayed
  trajectories.json     full minimize() runs at N in {5..8} (deferred modes)
"""
import importlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
import stochopy  # noqa: E402
from stochopy.factory import rosenbrock  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
BOUNDS2 = [[-5.12, 5.12], [-5.12, 5.12]]

# literals of /root/reference/tests/test_optimize.py:9-132 (method, options, xref)
CASES = [
    ("cmaes", {"constraints": None, "sigma": 0.1, "muperc": 0.5}, [0.29967256, 0.0803311]),
    ("cmaes", {"x0": [-5.0, -5.0], "constraints": None, "sigma": 0.1, "muperc": 0.5}, [0.99998135, 0.99995618]),
    ("cmaes", {"constraints": "Penalize", "sigma": 0.1, "muperc": 0.5}, [0.18765786, 0.05858025]),
    ("cmaes", {"x0": [-5.0, -5.0], "constraints": "Penalize", "sigma": 0.1, "muperc": 0.5}, [0.99998135, 0.99995618]),
    ("cpso", {"inertia": 0.7298, "constraints": None, "updating": "deferred"}, [0.95990315, 0.92082304]),
    ("cpso", {"inertia": 0.7298, "constraints": None, "updating": "immediate"}, [0.93258856, 0.86919435]),
    ("cpso", {"inertia": 0.91, "constraints": "Shrink", "updating": "deferred"}, [0.73752093, 0.54625484]),
    ("cpso", {"inertia": 0.91, "constraints": "Shrink", "updating": "immediate"}, [0.76668308, 0.58381385]),
    ("de", {"strategy": "rand1bin", "constraints": None, "updating": "deferred"}, [0.83228338, 0.68910339]),
    ("de", {"strategy": "rand2bin", "constraints": None, "updating": "deferred"}, [0.79409325, 0.60743767]),
    ("de", {"strategy": "best1bin", "constraints": None, "updating": "deferred"}, [1.00025932, 1.00051521]),
    ("de", {"strategy": "best2bin", "constraints": None, "updating": "deferred"}, [1.00515037, 1.01055037]),
    ("de", {"strategy": "rand1bin", "constraints": None, "updating": "immediate"}, [0.85658185, 0.726094]),
    ("de", {"strategy": "rand1bin", "constraints": "Random", "updating": "deferred"}, [1.02340815, 1.04590782]),
    ("de", {"strategy": "rand1bin", "constraints": "Random", "updating": "immediate"}, [0.99438151, 0.9944796]),
    ("na", {"nrperc": 0.5}, [1.14849912, 1.31885465]),
    ("pso", {"inertia": 0.7298, "constraints": None, "updating": "deferred"}, [0.95990315, 0.92082304]),
    ("pso", {"inertia": 0.7298, "constraints": None, "updating": "immediate"}, [0.95909508, 0.91977272]),
    ("pso", {"inertia": 0.91, "constraints": "Shrink", "updating": "deferred"}, [0.73752093, 0.54625484]),
    ("pso", {"inertia": 0.91, "constraints": "Shrink", "updating": "immediate"}, [0.76668308, 0.58381385]),
    ("vdcma", {"constraints": None, "sigma": 0.1, "muperc": 0.5}, [0.90013445, 0.85037782]),
    ("vdcma", {"x0": [-5.0, -5.0], "constraints": None, "sigma": 0.1, "muperc": 0.5}, [0.84059993, 0.69998341]),
    ("vdcma", {"constraints": "Penalize", "sigma": 0.1, "muperc": 0.5}, [0.90013445, 0.85037782]),
    ("vdcma", {"x0": [-5.0, -5.0], "constraints": "Penalize", "sigma": 0.1, "muperc": 0.5}, [0.82405114, 0.61993136]),
]
EXTRA = {
    "cpso": {"cognitivity": 1.49618, "sociability": 1.49618, "competitivity": 1.0},
    "pso": {"cognitivity": 1.49618, "sociability": 1.49618},
    "de": {"recombination": 0.1, "mutation": 0.5},
}


def _res(r):
    return dict(x=[float(v) for v in r.x], fun=float(r.fun), nit=int(r.nit), nfev=int(r.nfev), status=int(r.status))


def reference_cases():
    out = []
    for method, opts, xref in CASES:
        o = dict(opts)
        o.update(EXTRA.get(method, {}))
        o.update({"maxiter": 128, "popsize": 8, "seed": 42, "return_all": True})  # tests/helpers.py:14
        x0 = o.pop("x0", None)
        r = stochopy.optimize.minimize(rosenbrock, BOUNDS2, x0=x0, options=dict(o), method=method)
        assert np.allclose(xref, r.x), (method, opts)
        o.pop("return_all")
        out.append(dict(method=method, options=o, x0=x0, xref=xref, got=_res(r),
                        xall_shape=list(r.xall.shape), funall_last=[float(v) for v in r.funall[-1]]))
    r = stochopy.optimize.minimize(rosenbrock, BOUNDS2, method="cmaes",
                                   options={"maxiter": 100, "popsize": 10, "seed": 0})
    readme = dict(method="cmaes", options={"maxiter": 100, "popsize": 10, "seed": 0}, x0=None,
                  xref=[0.99997096, 0.99993643], got=_res(r),
                  readme=dict(fun=3.862267657514075e-09, nit=49, nfev=490, status=1))
    return dict(cases=out, readme=readme)


def factory():
    rs = np.random.RandomState(7)
    known = {  # tests/test_factory.py:7-17 at x = ones(10)
        "ackley": 3.625384938440362, "griewank": 0.8067591547236139, "quartic": 55.0,
        "rastrigin": 10.0, "rosenbrock": 0.0, "sphere": 10.0, "styblinski_tang": 341.6599,
    }
    rows = {}
    for n in (2, 7, 64, 128, 1000):
        X = rs.uniform(-5.12, 5.12, (5, n))
        rows[str(n)] = dict(X=X.tolist(), f={k: [float(getattr(stochopy.factory, k)(x)) for x in X] for k in known})
    ones = {k: float(getattr(stochopy.factory, k)(np.ones(10))) for k in known}
    return dict(known=known, ones=ones, rows=rows)


def mod(name):
    return importlib.import_module(name)


def steps():
    common = mod("stochopy.optimize._common")
    de = mod("stochopy.optimize.de._de")
    destrat = mod("stochopy.optimize.de._strategy")
    decons = mod("stochopy.optimize.de._constraints")
    cpso = mod("stochopy.optimize.cpso._cpso")
    pcons = mod("stochopy.optimize.cpso._constraints")
    cma = mod("stochopy.optimize.cmaes._cmaes")
    ccons = mod("stochopy.optimize.cmaes._constraints")
    vd = mod("stochopy.optimize.vdcma._vdcma")
    fac = stochopy.factory
    out = {}

    # lhs ------------------------------------------------------------------
    P, N = 24, 6
    bounds = np.stack([np.linspace(-5, -1, N), np.linspace(2, 7, N)], axis=1)
    np.random.seed(11)
    out["lhs_P"], out["lhs_N"], out["lhs_seed"] = P, N, 11
    out["lhs_bounds"] = bounds
    out["lhs_out"] = common.lhs(P, N, bounds)

    # de_sync, all strategies x constraints ------------------------------------
    P, N = 16, 6
    lower, upper = -np.ones(N) * 2.0, np.ones(N) * 2.0
    batched = lambda f: (lambda X: np.array([f(x) for x in X]))
    for s, strat in enumerate(["rand1bin", "rand2bin", "best1bin", "best2bin"]):
        for c, cons in enumerate([None, "Random"]):
            tag = f"de_{strat}_{cons}"
            rs = np.random.RandomState(100 + 10 * s + c)
            X = rs.uniform(-2.5, 2.5, (P, N))
            fun = batched(fac.rastrigin)
            pbestfit = fun(X)
            g = int(np.argmin(pbestfit))
            gbest = X[g].copy()
            out[tag + "_X0"], out[tag + "_pbestfit0"], out[tag + "_gbest0"] = X.copy(), pbestfit.copy(), gbest.copy()
            seed = 1000 + 10 * s + c
            out[tag + "_seed"] = seed
            np.random.seed(seed)
            r1 = np.random.rand(P, N)
            U = np.empty((P, N))
            Xn, gb, pbf, gfit, pfit, status = de.de_sync(
                5, X, U, gbest, pbestfit, pbestfit[g], None, 0.6, 0.7, r1, 100, 1e-8, 1e-8, fun,
                destrat._strategy_map[strat], decons._constraints_map[cons](lower, upper))
            out[tag + "_U"], out[tag + "_X1"], out[tag + "_pbestfit1"] = U.copy(), Xn.copy(), pbf.copy()
            out[tag + "_gbest1"], out[tag + "_gfit1"], out[tag + "_pfit"] = gb.copy(), gfit, pfit.copy()
            out[tag + "_status"] = -99 if status is None else status

    # pso_sync with NoConstraint / Shrink -----------------------------------
    P, N = 20, 5
    lower, upper = -np.ones(N) * 3.0, np.ones(N) * 3.0
    for c, cons in enumerate([None, "Shrink"]):
        tag = f"pso_{cons}"
        rs = np.random.RandomState(200 + c)
        X = rs.uniform(-3, 3, (P, N))
        V = rs.uniform(-4, 4, (P, N))
        pbest = X + rs.normal(0, 0.3, (P, N))
        fun = batched(fac.styblinski_tang)
        pbestfit = fun(pbest)
        g = int(np.argmin(pbestfit))
        gbest = pbest[g].copy()
        for k, v in dict(X0=X, V0=V, pbest0=pbest, pbestfit0=pbestfit, gbest0=gbest).items():
            out[f"{tag}_{k}"] = v.copy()
        r1, r2 = rs.rand(P, N), rs.rand(P, N)
        out[tag + "_r1"], out[tag + "_r2"] = r1, r2
        Xn, Vn, pb, gb, pbf, gfit, pfit, status = cpso.pso_sync(
            7, X, V, pbest, gbest, pbestfit, pbestfit[g], None, 0.8, 1.4, 1.6, r1, r2, 100, 1e-8, 1e-8, fun,
            pcons._constraints_map[cons](lower, upper, True))
        for k, v in dict(X1=Xn, V1=Vn, pbest1=pb, pbestfit1=pbf, gbest1=gb, pfit=pfit).items():
            out[f"{tag}_{k}"] = np.array(v, copy=True)
        out[tag + "_gfit1"] = gfit

    # restart (fires) ----------------------------------------------------------
    P, N = 32, 4
    rs = np.random.RandomState(300)
    gbest = rs.uniform(-1, 1, N)
    X = gbest + rs.normal(0, 1e-3, (P, N))
    V = rs.normal(0, 1, (P, N))
    pbest = X + rs.normal(0, 1e-4, (P, N))
    pbestfit = rs.uniform(0, 1, P)
    lower, upper = -np.ones(N) * 2, np.ones(N) * 2
    for k, v in dict(X0=X, V0=V, pbest0=pbest, pbestfit0=pbestfit, gbest=gbest).items():
        out[f"restart_{k}"] = v.copy()
    delta = np.log(1.0 + 0.003 * P) / np.max((0.2, np.log(0.01 * 50)))
    out["restart_delta"], out["restart_seed"] = delta, 301
    np.random.seed(301)
    Xn, Vn, pb, pbf = cpso.restart(10, X, V, pbest, gbest, pbestfit, lower, upper, 1.0, delta, 50)
    for k, v in dict(X1=Xn, V1=Vn, pbest1=pb, pbestfit1=pbf).items():
        out[f"restart_{k}"] = v.copy()

    # Penalize, three consecutive calls ----------------------------------------
    P, N = 12, 5
    rs = np.random.RandomState(400)
    fun = lambda X: np.array([fac.sphere(x * 5.12) for x in X])
    bw, hist, valid, ini = np.zeros(N), np.ones(1), False, True
    xold = rs.uniform(-1, 1, N)
    for call in range(3):
        xmean = xold + rs.normal(0, 0.4, N) + (0.5 if call else 0.0)
        arx = xmean + 0.3 * rs.normal(0, 1, (P, N))
        diagC = rs.uniform(0.5, 2.0, N)
        tag = f"pen{call}"
        for k, v in dict(arx=arx, xmean=xmean, xold=xold, diagC=diagC, bw0=bw, hist0=hist).items():
            out[f"{tag}_{k}"] = np.array(v, copy=True)
        out[tag + "_valid0"], out[tag + "_ini0"] = valid, ini
        fit, xv, bw, hist, valid, ini = ccons.Penalize(arx.copy(), arx, xmean, xold, 0.3, diagC, 3.2, call + 2,
                                                       bw.copy(), hist.copy(), valid, ini, fun)
        for k, v in dict(fit=fit, xvalid=xv, bw1=bw, hist1=hist).items():
            out[f"{tag}_{k}"] = np.array(v, copy=True)
        out[tag + "_valid1"], out[tag + "_ini1"] = valid, ini
        xold = xmean

    # pvec_and_qvec / ngv_ngd -------------------------------------------------------
    N, mu = 9, 6
    rs = np.random.RandomState(500)
    vvec = rs.normal(0, 1, N) / np.sqrt(N)
    dvec = rs.uniform(0.5, 1.5, N)
    nv2 = vvec @ vvec
    vn = vvec / np.sqrt(nv2)
    y = rs.normal(0, 1, (mu, N))
    w = np.log(mu + 0.5) - np.log(np.arange(1, mu + 1))
    w /= w.sum()
    p_mu, q_mu = vd.pvec_and_qvec(vn, nv2, y, w)
    p_1, q_1 = vd.pvec_and_qvec(vn, nv2, y[0])
    for k, v in dict(vvec=vvec, dvec=dvec, y=y, w=w, p_mu=p_mu, q_mu=q_mu, p_1=p_1, q_1=q_1).items():
        out[f"vd_{k}"] = v

    # converge ladder -----------------------------------------------------------
    rs = np.random.RandomState(600)
    conv_in, conv_out = [], []
    N, P = 4, 8
    for trial in range(60):
        it = int(rs.randint(1, 30))
        xmean = rs.normal(0, 1, N)
        xold = xmean + rs.normal(0, 10.0 ** rs.uniform(-12, 0), N)
        hist = np.zeros(40)
        hist[:it] = rs.uniform(0, 10.0 ** rs.uniform(-13, 1), it) + (0.0 if trial % 3 else 1.0)
        arfit = rs.uniform(0, 10.0 ** rs.uniform(-13, 1), P) + (0.0 if trial % 3 else 1.0)
        order = np.argsort(arfit)
        sigma = 10.0 ** rs.uniform(-12, 4)
        pc = rs.normal(0, 10.0 ** rs.uniform(-13, 0), N)
        diagC = rs.uniform(0.1, 10, N) * 10.0 ** rs.uniform(-3, 3)
        Q, _ = np.linalg.qr(rs.normal(0, 1, (N, N)))
        D = np.sort(rs.uniform(1e-6, 1, N) * 10.0 ** rs.uniform(0, 8, N))
        st = cma.converge(it, N, 25, xmean, xold, hist, arfit, order, sigma, 0.1, 12, pc, 1e-8, 1e-8, diagC, Q, D)
        st2 = cma.converge(it, N, 25, xmean, xold, hist, arfit, order, sigma, 0.1, 12, pc, 1e-8, 1e-8, diagC)
        conv_in.append(np.concatenate([[it], xmean, xold, hist, arfit, [sigma], pc, diagC, Q.ravel(), D]))
        conv_out.append([-99 if st is None else st, -99 if st2 is None else st2])
    out["conv_in"], out["conv_out"] = np.array(conv_in), np.array(conv_out)
    return out


def trajectories():
    fac = stochopy.factory
    runs = []

    def run(method, fun, N, P, maxiter, seed, **opts):
        bounds = [[-5.12, 5.12]] * N
        o = dict(opts, maxiter=maxiter, popsize=P, seed=seed)
        r = stochopy.optimize.minimize(getattr(fac, fun), bounds, method=method, options=dict(o))
        runs.append(dict(method=method, fun=fun, N=N, options=o, got=_res(r)))

    for strat in ("rand1bin", "rand2bin", "best1bin", "best2bin"):
        for cons in (None, "Random"):
            run("de", "rastrigin", 6, 24, 40, 3, strategy=strat, constraints=cons, updating="deferred",
                mutation=0.5, recombination=0.9)
    for cons in (None, "Shrink"):
        run("pso", "styblinski_tang", 5, 20, 40, 4, constraints=cons, updating="deferred")
        run("cpso", "sphere", 5, 24, 60, 5, constraints=cons, updating="deferred", competitivity=1.0)
    run("cpso", "rosenbrock", 3, 16, 200, 9, constraints=None, updating="deferred", competitivity=1.5)
    for cons in (None, "Penalize"):
        run("cmaes", "rosenbrock", 6, 12, 60, 6, constraints=cons)
        run("cmaes", "ackley", 5, 10, 80, 16, constraints=cons, sigma=0.5)
        run("vdcma", "rosenbrock", 8, 12, 60, 7, constraints=cons)
        run("vdcma", "sphere", 7, 10, 80, 17, constraints=cons, sigma=0.5)
    run("na", "sphere", 3, 8, 12, 8, nrperc=0.5)
    return runs


def main():
    with open(os.path.join(HERE, "reference_cases.json"), "w") as f:
        json.dump(reference_cases(), f, indent=1)
    with open(os.path.join(HERE, "factory.json"), "w") as f:
        json.dump(factory(), f)
    np.savez_compressed(os.path.join(HERE, "steps.npz"), **steps())
    with open(os.path.join(HERE, "trajectories.json"), "w") as f:
        json.dump(trajectories(), f, indent=1)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
