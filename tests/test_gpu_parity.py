"""GPU parity: the CUDA path (through the C ABI / the minimize() front-end)
against the oracle on the same seeded inputs and the same random draws.

Tolerances: integer / index work bit-exact; fp64 positions 1e-12 relative (the
north-star's contract is 1e-6), fp64 fitness 1e-13 * sum|terms|; fp32 positions
2e-6, fitness 4e-6 * sum|terms| (see gpu_util.tol_for)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import de as ode  # noqa: E402
from oracle import objectives as oobj  # noqa: E402
from oracle import pso as opso  # noqa: E402
from oracle.common import lhs_from_draws, select_sync  # noqa: E402
from oracle.streams import MTStream, PhiloxStream  # noqa: E402

G = os.path.join(os.path.dirname(__file__), "golden")
STEPS = np.load(os.path.join(G, "steps.npz"))
CASES = json.load(open(os.path.join(G, "reference_cases.json")))
TRAJ = json.load(open(os.path.join(G, "trajectories.json")))
FACT = json.load(open(os.path.join(G, "factory.json")))

DTYPES = ["float64", "float32"]


def close_fit(name, X, got, want, dtype, extra=1.0):
    from gpu_util import tol_for

    scale = oobj.term_magnitude(name, X)
    return np.all(np.abs(got - want) <= extra * tol_for(dtype)[1] * scale * max(1.0, np.sqrt(X.shape[1]) / 4) + 1e-300)


# ---- a1/a2 objectives ------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name", oobj.NAMES)
@pytest.mark.parametrize("N", [1, 2, 3, 7, 10, 16, 33, 64, 100, 128, 129, 256, 500, 1024, 2048])
def test_eval_matches_oracle(name, N, dtype):
    from gpu_util import device_eval

    if dtype == "float64" and N > 1024:
        pytest.skip("fp64 rows above 1024 are outside the compiled shapes")
    rs = np.random.RandomState(N * 7 + len(name))
    P = 257 if N <= 256 else 37
    X = rs.uniform(-5.12, 5.12, (P, N)).astype(dtype).astype(np.float64)
    X[0] = 1.0
    X[1] = 0.0
    want = oobj.evaluate_rows(name, X)
    got = device_eval(name, X, dtype)
    assert close_fit(name, X, got, want, dtype), np.abs(got - want).max()


@pytest.mark.parametrize("name", oobj.NAMES)
def test_factory_known_answers(name):
    import stochopy_b200

    f = getattr(stochopy_b200.factory, name)
    assert np.allclose(FACT["known"][name], f(np.ones(10)))  # tests/test_factory.py:20-23
    for n, blk in FACT["rows"].items():
        got = f.batch(np.array(blk["X"]))
        assert np.allclose(got, blk["f"][name], rtol=1e-11, atol=1e-9)


def test_eval_unstandardise():
    from gpu_util import device_eval

    rs = np.random.RandomState(3)
    X = rs.uniform(-1, 1, (100, 37))
    xs, xm = rs.uniform(1, 5, 37), rs.uniform(-1, 1, 37)
    got = device_eval("rastrigin", X, "float64", xs, xm)
    assert np.allclose(got, oobj.evaluate_rows("rastrigin", X * xs + xm), rtol=1e-13)


def test_eval_full_size_properties():
    """BASELINE size (P=65536, N=128, fp32): oracle on a row sample + invariances."""
    from gpu_util import device_eval

    rs = np.random.RandomState(0)
    X = rs.uniform(-5.12, 5.12, (65536, 128)).astype(np.float32).astype(np.float64)
    got = device_eval("rosenbrock", X, "float32")
    idx = rs.choice(65536, 512, replace=False)
    assert close_fit("rosenbrock", X[idx], got[idx], oobj.evaluate_rows("rosenbrock", X[idx]), "float32")
    perm = rs.permutation(65536)  # row order does not matter
    assert np.array_equal(device_eval("rosenbrock", X[perm], "float32"), got[perm])
    assert np.all(device_eval("sphere", X, "float32") >= 0)


# ---- a3 LHS ---------------------------------------------------------------------------
def test_lhs_reference_fixture():
    from gpu_util import device_lhs

    P, N, seed = int(STEPS["lhs_P"]), int(STEPS["lhs_N"]), int(STEPS["lhs_seed"])
    jitter, perms = MTStream(seed).lhs(P, N)
    got = device_lhs(P, N, STEPS["lhs_bounds"], 0, "float64", jitter, perms)
    assert np.array_equal(got, STEPS["lhs_out"])


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("P,N", [(8, 2), (100, 7), (4096, 64), (65536, 16)])
def test_lhs_philox(P, N, dtype):
    from gpu_util import device_lhs

    bounds = np.stack([np.linspace(-5, -1, N), np.linspace(2, 7, N)], axis=1)
    got = device_lhs(P, N, bounds, 1234567, dtype)
    st = PhiloxStream(1234567, dtype)
    want = lhs_from_draws(*st.lhs(P, N), bounds.astype(dtype))
    if dtype == "float64":
        assert np.array_equal(got, want)
    else:
        assert np.allclose(got, want, rtol=1e-6, atol=1e-6)
    if dtype == "float32":
        return  # fp32 rounding of the scaled value can cross a 2/P stratum edge; the fp64 check covers the property
    lo, hi = bounds.T  # latin property: every stratum hit once per column
    strata = np.floor(((got - 0.5 * (hi + lo)) / (0.5 * (hi - lo)) + 1.0) * P / 2.0 + 1e-6).astype(int)
    assert all(np.array_equal(np.sort(strata[:, j]), np.arange(P)) for j in range(N))


# ---- a6-a10 DE --------------------------------------------------------------------------
@pytest.mark.parametrize("strategy", ["rand1bin", "rand2bin", "best1bin", "best2bin"])
@pytest.mark.parametrize("cons", [None, "Random"])
def test_de_generation_reference_fixture(strategy, cons):
    """Reference de_sync outputs (recorded) reproduced with the reference's own draws."""
    from gpu_util import DeRig

    t = f"de_{strategy}_{cons}_"
    X0, N = STEPS[t + "X0"], STEPS[t + "X0"].shape[1]
    lower, upper = -2.0 * np.ones(N), 2.0 * np.ones(N)
    rig = DeRig(X0, STEPS[t + "pbestfit0"], STEPS[t + "gbest0"], "rastrigin", strategy, cons, 0.6, 0.7, lower, upper,
                maxiter=100, it0=5)
    r1, donors, irand, rep = MTStream(int(STEPS[t + "seed"])).de(5, X0.shape[0], N, ode.DONORS[strategy], lower, upper,
                                                                 cons == "Random")
    rig.draws(r1, donors, irand, rep)
    rig.step(5)
    out = rig.get(5)
    assert np.array_equal(out["X"], STEPS[t + "X1"])  # positions bit-exact (same op order, no FMA)
    assert np.allclose(out["pbestfit"], STEPS[t + "pbestfit1"], rtol=1e-13)
    assert np.allclose(out["pfit"], STEPS[t + "pfit"], rtol=1e-13)
    assert np.array_equal(out["gbest"], STEPS[t + "gbest1"])
    assert np.isclose(out["ctrl"].gfit, float(STEPS[t + "gfit1"]), rtol=1e-13)
    assert out["ctrl"].status == -1000 and out["ctrl"].nit == 5


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("strategy,cons,obj,P,N", [
    ("best1bin", None, "rosenbrock", 4096, 128),
    ("rand1bin", "Random", "rastrigin", 1000, 128),
    ("rand2bin", None, "styblinski_tang", 333, 64),
    ("best2bin", "Random", "ackley", 257, 10),
    ("best1bin", None, "griewank", 64, 2),
    ("rand1bin", None, "sphere", 50, 300),
    ("best1bin", "Random", "quartic", 40, 1000),
])
def test_de_generation_philox_vs_oracle(strategy, cons, obj, P, N, dtype):
    """In-kernel Philox draws: the oracle is fed the same draws (oracle/philox.py)."""
    from gpu_util import DeRig, tol_for

    dt = np.dtype(dtype)
    rs = np.random.RandomState(P + N)
    lower, upper = -5.12 * np.ones(N), 5.12 * np.ones(N)
    X = rs.uniform(-5.5, 5.5, (P, N)).astype(dt)
    pbestfit = oobj.evaluate_rows(obj, X.astype(np.float64)).astype(dt)
    gbest = X[int(np.argmin(pbestfit))].copy()
    seed = 0xC0FFEE12345 + P
    rig = DeRig(X, pbestfit, gbest, obj, strategy, cons, 0.5, 0.9, lower, upper, seed=seed, dtype=dtype)
    stream = PhiloxStream(seed, dt)
    Xo, po, go = X.copy(), pbestfit.copy(), gbest.copy()
    for it in (2, 3, 4):
        r1, donors, irand, rep = stream.de(it, P, N, ode.DONORS[strategy], lower, upper, cons == "Random")
        U = ode.trial_population(Xo, go, strategy, dt.type(0.5), dt.type(0.9), r1, donors, irand, rep,
                                 lower.astype(dt), upper.astype(dt))
        rig.step(it)
        out = rig.get(it)
        fU = oobj.evaluate_rows(obj, U.astype(np.float64))
        assert close_fit(obj, U.astype(np.float64), out["pfit"], fU, dtype), it
        # selection decided by the device fitness (ties in the last bit may flip it): replay it in the oracle
        go, gf, _ = select_sync(it, U, out["pfit"].astype(dt), go, Xo, po, 1000, 1e-8, 1e-8)
        if dtype == "float64":
            assert np.array_equal(out["X"], Xo), it
        else:
            assert np.allclose(out["X"], Xo, rtol=tol_for(dtype)[0], atol=1e-6), it
        assert np.allclose(out["pbestfit"], po, rtol=1e-6) and np.allclose(out["gbest"], go, rtol=1e-6, atol=1e-6)
        assert out["ctrl"].gbest_row == int(np.argmin(po)) and out["ctrl"].nit == it
        if cons == "Random":
            assert (out["X"] >= -5.5).all() and (U >= lower - 1e-6).all() and (U <= upper + 1e-6).all()


def test_de_full_size_properties():
    """C2 / headline size (P=65536, N=128, fp32): invariants of a generation."""
    from gpu_util import DeRig, device_eval

    P, N = 65536, 128
    rs = np.random.RandomState(1)
    X = rs.uniform(-5.12, 5.12, (P, N)).astype(np.float32)
    fit = device_eval("rosenbrock", X.astype(np.float64), "float32").astype(np.float32)
    gbest = X[int(np.argmin(fit))]
    rig = DeRig(X, fit, gbest, "rosenbrock", "best1bin", None, 0.5, 0.9, -5.12 * np.ones(N), 5.12 * np.ones(N),
                seed=99, dtype="float32")
    prev, Xprev = fit.astype(np.float64), X.astype(np.float64)
    for it in range(2, 8):
        rig.step(it)
        out = rig.get(it)
        assert np.all(out["pbestfit"] <= prev)  # greedy selection never worsens an individual
        kept = out["pbestfit"] == prev
        assert np.array_equal(out["X"][kept], Xprev[kept])  # unchanged rows are copied verbatim
        assert np.all(out["pfit"][~kept] == out["pbestfit"][~kept])
        assert np.array_equal(device_eval("rosenbrock", out["X"], "float32"), out["pbestfit"])  # fitness belongs to rows
        b = int(np.argmin(out["pbestfit"]))
        assert out["ctrl"].gbest_row == b and out["ctrl"].gfit == out["pbestfit"][b]
        assert np.array_equal(out["gbest"], out["X"][b])
        prev, Xprev = out["pbestfit"], out["X"]
    assert (~kept).sum() > 0


# ---- a11-a14 PSO / CPSO ---------------------------------------------------------------
@pytest.mark.parametrize("cons", [None, "Shrink"])
def test_pso_generation_reference_fixture(cons):
    from gpu_util import PsoRig

    t = f"pso_{cons}_"
    N = STEPS[t + "X0"].shape[1]
    rig = PsoRig(STEPS[t + "X0"], STEPS[t + "V0"], STEPS[t + "pbest0"], STEPS[t + "pbestfit0"], STEPS[t + "gbest0"],
                 "styblinski_tang", cons, 0.8, 1.4, 1.6, -3.0 * np.ones(N), 3.0 * np.ones(N), maxiter=100)
    rig.draws(STEPS[t + "r1"], STEPS[t + "r2"])
    rig.step(7)
    out = rig.get()
    for k in ("X", "V", "pbest", "gbest"):
        assert np.array_equal(out[k], STEPS[t + k + "1"]), k
    assert np.allclose(out["pbestfit"], STEPS[t + "pbestfit1"], rtol=1e-13)
    assert np.allclose(out["pfit"], STEPS[t + "pfit"], rtol=1e-13)
    if cons == "Shrink":  # lands on the bound within the reference's own tolerance (tests/helpers.py:23-25)
        assert (out["X"] + 1e-15 >= -3.0).all() and (out["X"] - 1e-15 <= 3.0).all()


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("cons,obj,P,N", [(None, "styblinski_tang", 4096, 64), ("Shrink", "rosenbrock", 1000, 128),
                                          ("Shrink", "rastrigin", 100, 7), (None, "ackley", 33, 513)])
def test_pso_generation_philox_vs_oracle(cons, obj, P, N, dtype):
    from gpu_util import PsoRig, tol_for

    dt = np.dtype(dtype)
    rs = np.random.RandomState(P * 3 + N)
    lower, upper = (-5.12 * np.ones(N)).astype(dt), (5.12 * np.ones(N)).astype(dt)
    X = rs.uniform(-5.12, 5.12, (P, N)).astype(dt)
    V = rs.uniform(-3, 3, (P, N)).astype(dt)
    pbest = (X + rs.normal(0, 0.5, (P, N))).astype(dt)
    pbestfit = oobj.evaluate_rows(obj, pbest.astype(np.float64)).astype(dt)
    gbest = pbest[int(np.argmin(pbestfit))].copy()
    seed = 777 + N
    rig = PsoRig(X, V, pbest, pbestfit, gbest, obj, cons, 0.7298, 1.49618, 1.49618, lower, upper, seed=seed, dtype=dtype)
    stream = PhiloxStream(seed, dt)
    w, c1, c2 = dt.type(0.7298), dt.type(1.49618), dt.type(1.49618)
    for it in (2, 3, 4):
        r1, r2 = stream.pso(it, P, N)
        X, V = opso.move(X, V, pbest, gbest, w, c1, c2, r1, r2, cons, lower, upper)
        X, V = X.astype(dt), V.astype(dt)
        rig.step(it)
        out = rig.get()
        tol = tol_for(dtype)[0]
        if dtype == "float64":
            assert np.array_equal(out["X"], X) and np.array_equal(out["V"], V), it
        else:
            assert np.allclose(out["X"], X, rtol=tol, atol=1e-5) and np.allclose(out["V"], V, rtol=tol, atol=1e-5), it
            X, V = out["X"].astype(dt), out["V"].astype(dt)  # fp32: continue from the device state
        assert close_fit(obj, X.astype(np.float64), out["pfit"], oobj.evaluate_rows(obj, X.astype(np.float64)), dtype)
        gbest, gf, _ = select_sync(it, X, out["pfit"].astype(dt), gbest, pbest, pbestfit, 1000, 1e-8, 1e-8)
        assert np.allclose(out["pbest"], pbest, rtol=tol, atol=1e-5) and np.allclose(out["pbestfit"], pbestfit, rtol=1e-6)
        assert out["ctrl"].gbest_row == int(np.argmin(pbestfit))
        if cons == "Shrink":
            assert (out["X"] >= -5.12 - 1e-5).all() and (out["X"] <= 5.12 + 1e-5).all()


def test_cpso_restart_reference_fixture():
    from gpu_util import PsoRig

    t = "restart_"
    X0, N = STEPS[t + "X0"], STEPS[t + "X0"].shape[1]
    lower, upper = -2.0 * np.ones(N), 2.0 * np.ones(N)
    rig = PsoRig(X0, STEPS[t + "V0"], STEPS[t + "pbest0"], STEPS[t + "pbestfit0"], STEPS[t + "gbest"], "sphere", None,
                 0.7, 1.5, 1.5, lower, upper, maxiter=50, gamma=1.0, delta=float(STEPS[t + "delta"]))
    ms = MTStream(int(STEPS[t + "seed"]))
    nw = rig.restart(10, lambda n: ms.pso_restart(10, np.arange(n), N, lower, upper))
    out = rig.get()
    assert nw == int((STEPS[t + "pbestfit1"] == 1e30).sum()) and nw > 0
    for k in ("X", "V", "pbest", "pbestfit"):
        assert np.array_equal(out[k], STEPS[t + k + "1"]), k


def test_cpso_restart_philox_and_wide_swarm():
    from gpu_util import PsoRig

    P, N = 3000, 24
    rs = np.random.RandomState(5)
    lower, upper = -5.0 * np.ones(N), 5.0 * np.ones(N)
    gbest = rs.uniform(-1, 1, N)
    X = gbest + rs.normal(0, 1e-3, (P, N))
    pbestfit = rs.uniform(0, 1, P)
    delta = opso.swarm_delta(P, 100)
    rig = PsoRig(X, rs.normal(0, 1, (P, N)), X.copy(), pbestfit, gbest, "sphere", None, 0.7, 1.5, 1.5, lower, upper,
                 seed=42, maxiter=100, gamma=1.2, delta=delta)
    rows = opso.restart_plan(20, X, gbest, pbestfit.copy(), 1.2, delta, 100)
    nw = rig.restart(20)
    out = rig.get()
    assert nw == len(rows) > 0
    hit = np.where(out["pbestfit"] == 1e30)[0]
    assert np.array_equal(np.sort(hit), np.sort(rows))  # same worst-nw set as argsort()[:-nw-1:-1]
    fresh = PhiloxStream(42).pso_restart(20, hit, N, lower, upper)
    assert np.array_equal(out["X"][hit], fresh) and np.array_equal(out["pbest"][hit], fresh)
    assert np.all(out["V"][hit] == 0) and np.array_equal(np.delete(out["X"], hit, 0), np.delete(X, hit, 0))
    # a wide swarm must not restart (long run -> small threshold, _cpso.py:216)
    rig2 = PsoRig(rs.uniform(-5, 5, (P, N)), X, X, pbestfit, gbest, "sphere", None, 0.7, 1.5, 1.5, lower, upper,
                  maxiter=100000, gamma=1.2, delta=opso.swarm_delta(P, 100000))
    assert rig2.restart(20) == 0


# ---- whole runs through minimize() ------------------------------------------------------
def _sync_cases():
    out = []
    for c in CASES["cases"]:
        if c["method"] in ("de", "pso", "cpso") and c["options"].get("updating") == "deferred":
            out.append(c)
    return out


@pytest.mark.parametrize("case", _sync_cases(), ids=lambda c: c["method"] + "-" + str(c["options"].get("strategy")) +
                         "-" + str(c["options"].get("constraints")))
def test_reference_known_answers_through_minimize(case):
    """The reference's own golden xref (tests/test_optimize.py) with rng='numpy'."""
    import stochopy_b200 as sb

    o = dict(case["options"], rng="numpy", return_all=True)
    r = sb.optimize.minimize(sb.factory.rosenbrock, [[-5.12, 5.12]] * 2, x0=case["x0"], options=o, method=case["method"])
    assert np.allclose(case["xref"], r.x)  # the reference's assertion, tests/helpers.py:22
    got = case["got"]
    assert (r.nit, r.nfev, r.status) == (got["nit"], got["nfev"], got["status"])
    assert np.allclose(r.x, got["x"], rtol=1e-9) and np.isclose(r.fun, got["fun"], rtol=1e-7, atol=1e-14)
    assert list(r.xall.shape) == case["xall_shape"]
    assert np.allclose(r.funall[-1], case["funall_last"], rtol=1e-7, atol=1e-12)
    if o.get("constraints"):
        assert np.all(r.xall + 1e-15 >= -5.12) and np.all(r.xall - 1e-15 <= 5.12)


@pytest.mark.parametrize("run", [t for t in TRAJ if t["method"] in ("de", "pso", "cpso")],
                         ids=lambda r: f"{r['method']}-{r['fun']}-{r['options'].get('strategy')}-{r['options'].get('constraints')}-{r['options'].get('competitivity')}")
def test_reference_trajectories_through_minimize(run):
    import stochopy_b200 as sb

    o = dict(run["options"], rng="numpy")
    r = sb.optimize.minimize(getattr(sb.factory, run["fun"]), [[-5.12, 5.12]] * run["N"], options=o, method=run["method"])
    got = run["got"]
    assert (r.nit, r.nfev, r.status) == (got["nit"], got["nfev"], got["status"])
    assert np.allclose(r.x, got["x"], rtol=1e-7, atol=1e-10) and np.isclose(r.fun, got["fun"], rtol=1e-6, atol=1e-12)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("strategy,cons,ftol,maxiter", [("best1bin", None, 5.0e2, 400), ("rand1bin", "Random", -1.0, 23),
                                                        ("best2bin", None, 1.0e9, 50)])
def test_de_chained_generations_equal_unchained(strategy, cons, ftol, maxiter, dtype):
    """sp_de_run chains its generations (SP_CHAIN_OUT / SP_CHAIN_IN: the per-generation
    argmin / gbest / status resolution moves into the next launch's prologue).  State,
    nit and status must equal the one-launch-per-generation path bit for bit, including
    a stop in the middle of a chunk (ftol), at the chunk's first generation (1e9) and by maxiter."""
    import ctypes as C

    from gpu_util import DeRig, device_eval
    from stochopy_b200 import _lib as L

    P, N = 3001, 128
    rs = np.random.RandomState(3)
    X = rs.uniform(-5.12, 5.12, (P, N)).astype(dtype)
    fit = device_eval("sphere", X.astype(np.float64), dtype).astype(dtype)
    gbest = X[int(np.argmin(fit))]
    lo, hi = -5.12 * np.ones(N), 5.12 * np.ones(N)

    def rig():
        return DeRig(X, fit, gbest, "sphere", strategy, cons, 0.5, 0.9, lo, hi, seed=11, dtype=dtype, maxiter=maxiter,
                     xtol=1e-8, ftol=ftol)

    a = rig()
    assert L.load().sp_de_chainable(C.byref(a.st)) == 1
    it = 1
    while a.eng.read_ctrl(a.ctrl).status == L.SP_RUNNING:
        it += 1
        a.step(it)
    ca = a.eng.read_ctrl(a.ctrl)
    assert ca.nit == it and (ftol < 0 or ca.nit < maxiter or ftol > 1e8)

    b = rig()
    nxt = 2
    while b.eng.read_ctrl(b.ctrl).status == L.SP_RUNNING:
        L.call("sp_de_run", C.byref(b.st), nxt, 7, b.eng.stream)
        nxt += 7
    cb = b.eng.read_ctrl(b.ctrl)
    assert (cb.nit, cb.status, cb.gbest_row, cb.gfit) == (ca.nit, ca.status, ca.gbest_row, ca.gfit)
    assert np.isclose(cb.dist, ca.dist, rtol=1e-12, atol=0)
    oa, ob = a.get(ca.nit), b.get(cb.nit)
    for k in ("X", "pbestfit", "pfit", "gbest"):
        assert np.array_equal(oa[k], ob[k]), k


@pytest.mark.parametrize("method,opts", [
    ("de", dict(strategy="best1bin")), ("de", dict(strategy="rand1bin", constraints="Random")),
    ("pso", dict(constraints="Shrink")), ("cpso", dict(competitivity=1.0)),
])
def test_philox_runs_match_oracle(method, opts):
    """Performance mode (in-kernel Philox), fp64: whole runs against the oracle
    driven by the same counter-based stream."""
    import stochopy_b200 as sb

    N, P, seed = 12, 96, 2024
    bounds = [[-5.12, 5.12]] * N
    o = dict(opts, maxiter=60, popsize=P, seed=seed, updating="deferred")
    r = sb.optimize.minimize(sb.factory.rastrigin, bounds, options=dict(o), method=method)
    drv = ode.minimize if method == "de" else opso.minimize
    if method == "pso":
        o["competitivity"] = None
    w = drv(oobj.rastrigin, bounds, stream=PhiloxStream(seed), **o)
    assert (r.nit, r.status) == (w["nit"], w["status"])
    assert np.allclose(r.x, w["x"], rtol=1e-9, atol=1e-12) and np.isclose(r.fun, w["fun"], rtol=1e-9)


def test_host_objective_follows_the_same_path():
    """Arbitrary Python fun(x, *args): propose/select on device, fun on host."""
    import stochopy_b200 as sb

    calls = []

    def fun(x, shift):
        calls.append(x.shape)
        return oobj.rosenbrock(x) + shift

    b = [[-5.12, 5.12]] * 3
    for method in ("de", "pso"):
        o = dict(maxiter=25, popsize=12, seed=5, rng="numpy", updating="deferred")
        a = sb.optimize.minimize(fun, b, args=(0.0,), options=dict(o), method=method)
        d = sb.optimize.minimize(sb.factory.rosenbrock, b, options=dict(o), method=method)
        assert np.allclose(a.x, d.x, rtol=1e-12) and a.nit == d.nit and a.nfev == d.nfev == 25 * 12
    assert set(calls) == {(3,)}


@pytest.mark.parametrize("method", ["de", "pso", "cpso", "na", "cmaes", "vdcma"])
def test_callback_contract(method):
    """callback(X, state) once per iteration incl. the initial population (tests/test_optimize.py:135-152)."""
    import stochopy_b200 as sb

    seen = []
    r = sb.optimize.minimize(sb.factory.rosenbrock, [[-5.12, 5.12]] * 2, method=method, options={"maxiter": 7, "seed": 1},
                             callback=lambda X, s: seen.append((X.shape, s.nit, s.nfev)))
    assert len(seen) == 7 and seen[0] == ((10, 2), 1, 10) and seen[-1][1] == r.nit == 7


def test_termination_codes():
    import stochopy_b200 as sb

    b = [[-5.12, 5.12]] * 2
    r = sb.optimize.minimize(sb.factory.sphere, b, method="de", options=dict(maxiter=500, popsize=20, seed=0, ftol=1e-3))
    assert r.status in (0, 1) and r.success and r.fun <= 1e-3 and r.nit < 500
    r = sb.optimize.minimize(sb.factory.sphere, b, method="pso", options=dict(maxiter=5, popsize=20, seed=0, ftol=-1.0))
    assert r.status == -1 and not r.success and r.nit == 5 and r.nfev == 100
    assert r.message == "maximum number of iterations is reached"


# ---- a19 NA -----------------------------------------------------------------------------------
def test_na_reference_known_answer():
    import stochopy_b200 as sb

    case = [c for c in CASES["cases"] if c["method"] == "na"][0]
    o = dict(case["options"], rng="numpy", return_all=True)
    r = sb.optimize.minimize(sb.factory.rosenbrock, [[-5.12, 5.12]] * 2, options=o, method="na")
    assert np.allclose(case["xref"], r.x)
    got = case["got"]
    assert (r.nit, r.nfev, r.status) == (got["nit"], got["nfev"], got["status"])
    assert np.allclose(r.x, got["x"], rtol=1e-9) and np.isclose(r.fun, got["fun"], rtol=1e-7)
    assert list(r.xall.shape) == case["xall_shape"] and np.allclose(r.funall[-1], case["funall_last"], rtol=1e-7)


def test_na_reference_trajectory_and_philox():
    import stochopy_b200 as sb
    from oracle import na as ona

    run = [t for t in TRAJ if t["method"] == "na"][0]
    r = sb.optimize.minimize(getattr(sb.factory, run["fun"]), [[-5.12, 5.12]] * run["N"],
                             options=dict(run["options"], rng="numpy"), method="na")
    got = run["got"]
    assert (r.nit, r.status) == (got["nit"], got["status"]) and np.allclose(r.x, got["x"], rtol=1e-8)
    # in-kernel Philox draws vs the oracle fed with the same stream; one fixed (zero-span) axis
    bounds = [[-5.12, 5.12], [1.5, 1.5], [-2.0, 3.0], [-5.12, 5.12]]
    o = dict(maxiter=10, popsize=12, seed=77, nrperc=0.4)
    w = ona.minimize(oobj.sphere, bounds, stream=PhiloxStream(77), **o)
    r = sb.optimize.minimize(sb.factory.sphere, bounds, options=dict(o), method="na")
    assert (r.nit, r.status) == (w["nit"], w["status"])
    assert np.allclose(r.x, w["x"], rtol=1e-9, atol=1e-12) and np.isclose(r.fun, w["fun"], rtol=1e-9)
    assert r.x[1] == 1.5


def test_na_direct_call_default_callback_quirk():
    import stochopy_b200 as sb

    with pytest.raises(ValueError):  # callback=True default is not callable (_na.py:26,113-114)
        sb.optimize.na(sb.factory.sphere, [[-1, 1]] * 2)


# ---- 8f-1: return_all streamed through a device ring + side stream -------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("method,opts", [("de", dict(strategy="best1bin")), ("pso", {}), ("cpso", dict(competitivity=1.0))])
def test_return_all_streaming_equals_synchronous_history(method, opts, dtype):
    """Without a callback the per-generation snapshots leave through HistoryStreamer (no host
    sync per generation); with a callback the synchronous path records them.  Same arrays,
    including a run that terminates inside a chunk (ftol) -- xall/funall are cut at nit."""
    import stochopy_b200 as sb

    b = [[-5.12, 5.12]] * 6
    for ftol in (-1.0, 0.5):
        o = dict(opts, maxiter=90, popsize=40, seed=17, dtype=dtype, updating="deferred", return_all=True,
                 verbosity=0.5, ftol=ftol)
        a = sb.optimize.minimize(sb.factory.sphere, b, method=method, options=dict(o))
        d = sb.optimize.minimize(sb.factory.sphere, b, method=method, options=dict(o), callback=lambda X, s: None)
        assert (a.nit, a.status) == (d.nit, d.status) and (ftol < 0) == (a.nit == 90)
        assert a.xall.shape == d.xall.shape == (a.nit, 20, 6) and a.funall.shape == (a.nit, 20)
        assert np.array_equal(a.xall, d.xall) and np.array_equal(a.funall, d.funall)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("N,P,cons,ftol,maxiter", [(64, 3001, None, 2.0e2, 300), (200, 777, "Shrink", -1.0, 23),
                                                   (64, 5000, None, 1.0e9, 40)])
def test_pso_chained_generations_equal_unchained(N, P, cons, ftol, maxiter, dtype):
    """sp_pso_run chains the generations of a plain PSO (per-CTA minima + best rows left for the
    next launch's prologue).  The callback path launches one unchained generation at a time:
    same x / fun / nit / status, incl. stops by ftol inside a chunk, at once, and by maxiter."""
    import stochopy_b200 as sb

    b = [[-5.12, 5.12]] * N
    o = dict(maxiter=maxiter, popsize=P, seed=4, dtype=dtype, constraints=cons, ftol=ftol, updating="deferred")
    a = sb.optimize.minimize(sb.factory.sphere, b, method="pso", options=dict(o))
    d = sb.optimize.minimize(sb.factory.sphere, b, method="pso", options=dict(o), callback=lambda X, s: None)
    assert (a.nit, a.status, a.nfev) == (d.nit, d.status, d.nfev)
    assert np.array_equal(a.x, d.x) and a.fun == d.fun
    assert (ftol < 0) == (a.status == -1)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("maxiter,P,N", [(30, 500, 10), (400, 300, 6), (150, 4000, 64)])
def test_cpso_lazy_restart_equals_eager(maxiter, P, N, dtype):
    """Fast path: sp_pso_run_lazy parks the run when a restart fires and the host resumes it
    (sp_cpso_restart_resume), falling back to gated in-chunk restarts when they come in runs.
    The callback path restarts eagerly after every generation.  maxiter=30 makes delta large
    (restart every generation), 400 / 150 give runs with rare and clustered restarts."""
    import stochopy_b200 as sb

    b = [[-5.12, 5.12]] * N
    o = dict(maxiter=maxiter, popsize=P, seed=8, dtype=dtype, competitivity=1.0, updating="deferred", xtol=-1.0,
             ftol=-1.0e300)
    a = sb.optimize.minimize(sb.factory.rastrigin, b, method="cpso", options=dict(o))
    d = sb.optimize.minimize(sb.factory.rastrigin, b, method="cpso", options=dict(o), callback=lambda X, s: None)
    assert (a.nit, a.status, a.nfev) == (d.nit, d.status, d.nfev) == (maxiter, -1, maxiter * P)
    assert np.array_equal(a.x, d.x) and a.fun == d.fun


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("method,opts", [("cmaes", {}), ("cmaes", dict(constraints="Penalize")), ("vdcma", {}),
                                         ("vdcma", dict(constraints="Penalize"))])
def test_return_all_streaming_es_methods(method, opts, dtype):
    """CMA-ES / VD-CMA: the streamed history (standardised rows un-standardised and clipped on the
    host at the end) equals the synchronous per-generation history."""
    import stochopy_b200 as sb

    b = [[-3.0, 5.0]] * 7
    for ftol in (-1.0, 1.0e-3):
        o = dict(opts, maxiter=60, popsize=16, seed=5, dtype=dtype, return_all=True, verbosity=0.5, ftol=ftol)
        a = sb.optimize.minimize(sb.factory.sphere, b, method=method, options=dict(o))
        d = sb.optimize.minimize(sb.factory.sphere, b, method=method, options=dict(o), callback=lambda X, s: None)
        assert (a.nit, a.status) == (d.nit, d.status)
        assert a.xall.shape == d.xall.shape == (a.nit, 8, 7)
        assert np.array_equal(a.xall, d.xall) and np.array_equal(a.funall, d.funall)
        assert np.array_equal(a.x, d.x) and a.fun == d.fun


@pytest.mark.parametrize("method,opts", [("de", dict(strategy="best1bin", updating="deferred")), ("vdcma", {})])
def test_return_all_streaming_window_wraps(method, opts, monkeypatch):
    """The pinned host side of HistoryStreamer is a ring: with a window of a few generations (PIN_LIMIT
    forced tiny) the history still equals the synchronous one -- no fallback to per-generation syncs."""
    import stochopy_b200 as sb
    from stochopy_b200.optimize import _common

    monkeypatch.setattr(_common.HistoryStreamer, "PIN_LIMIT", 1)  # window = SLOTS generations
    b = [[-3.0, 5.0]] * 7
    for ftol in (-1.0, 1.0e-3):
        o = dict(opts, maxiter=70, popsize=16, seed=5, return_all=True, verbosity=0.5, ftol=ftol)
        a = sb.optimize.minimize(sb.factory.sphere, b, method=method, options=dict(o))
        d = sb.optimize.minimize(sb.factory.sphere, b, method=method, options=dict(o), callback=lambda X, s: None)
        assert (a.nit, a.status) == (d.nit, d.status)
        assert np.array_equal(a.xall, d.xall) and np.array_equal(a.funall, d.funall)


@pytest.mark.parametrize("dtype", DTYPES)
def test_cpso_undecided_bound_takes_the_exact_radius_path(dtype, monkeypatch):
    """The lazy CPSO run decides the restart from a bound on the swarm radius inside the generation kernel and parks
    with flag = -1 when the bound cannot tell; sp_cpso_restart_resume then decides with the exact radius kernel.
    SP_CPSO_FORCE_AMBIGUOUS sends EVERY generation down that path: the trajectory must still be the eager one's."""
    import stochopy_b200 as sb

    b = [[-5.12, 5.12]] * 6
    o = dict(maxiter=60, popsize=200, seed=8, dtype=dtype, competitivity=1.0, updating="deferred", xtol=-1.0, ftol=-1.0e300)
    d = sb.optimize.minimize(sb.factory.rastrigin, b, method="cpso", options=dict(o), callback=lambda X, s: None)
    monkeypatch.setenv("SP_CPSO_FORCE_AMBIGUOUS", "1")
    a = sb.optimize.minimize(sb.factory.rastrigin, b, method="cpso", options=dict(o))
    assert (a.nit, a.status, a.nfev) == (d.nit, d.status, d.nfev) == (60, -1, 60 * 200)
    assert np.array_equal(a.x, d.x) and a.fun == d.fun
