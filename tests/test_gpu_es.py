"""GPU parity of the evolution-strategy path (CMA-ES, VD-CMA) and its building blocks."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import cmaes as ocma  # noqa: E402
from oracle import objectives as oobj  # noqa: E402
from oracle import vdcma as ovd  # noqa: E402
from oracle.streams import PhiloxStream  # noqa: E402

G = os.path.join(os.path.dirname(__file__), "golden")
CASES = json.load(open(os.path.join(G, "reference_cases.json")))
TRAJ = json.load(open(os.path.join(G, "trajectories.json")))


def canonical_eigh(C):
    """np.linalg.eigh with the device solver's sign rule: largest |component| positive."""
    w, V = np.linalg.eigh(C)
    for j in range(V.shape[1]):
        k = int(np.argmax(np.abs(V[:, j])))
        if V[k, j] < 0:
            V[:, j] = -V[:, j]
    return w, V


def device_eigh(C, dtype="float64"):
    import torch

    from stochopy_b200 import _lib as L
    from stochopy_b200.optimize._common import Engine

    eng = Engine(dtype)
    N = C.shape[0]
    dC = torch.from_numpy(np.ascontiguousarray(C, dtype=eng.np_dt)).to(eng.device)
    w, B, work = eng.zeros(N), eng.zeros(N, N), eng.zeros(int(L.load().sp_sym_eigh_work_scalars(N)))
    L.call("sp_sym_eigh", eng.sp_dt, dC.data_ptr(), N, w.data_ptr(), B.data_ptr(), work.data_ptr(), 0, None, eng.stream)
    eng.sync()
    return w.cpu().numpy().astype(np.float64), B.cpu().numpy().astype(np.float64), dC.cpu().numpy().astype(np.float64)


@pytest.mark.parametrize("N", [1, 2, 3, 5, 16, 33, 64, 100, 113, 114, 128, 200, 256, 301])
def test_sym_eigh_matches_lapack(N):
    rs = np.random.RandomState(N)
    A = rs.normal(0, 1, (N, N))
    C = A @ A.T / N + np.diag(rs.uniform(0.01, 3.0, N))
    Cu = np.triu(C) + rs.normal(0, 1, (N, N)) * np.tri(N, k=-1)  # garbage below the diagonal: only triu counts
    w, B, Csym = device_eigh(Cu)
    wr, Vr = canonical_eigh(C)
    assert np.allclose(Csym, C, rtol=0, atol=0)  # symmetrised from the upper triangle (_cmaes.py:303)
    assert np.allclose(w, wr, rtol=1e-11, atol=1e-13)
    assert np.allclose(B.T @ B, np.eye(N), atol=1e-12)
    assert np.allclose((B * w) @ B.T, C, rtol=1e-11, atol=1e-12)
    assert all(B[np.argmax(np.abs(B[:, j])), j] > 0 for j in range(N))
    if N <= 64:  # well separated spectrum: eigenvectors themselves agree
        assert np.allclose(B, Vr, atol=1e-8)


def test_sym_eigh_fp32_and_degenerate():
    rs = np.random.RandomState(1)
    A = rs.normal(0, 1, (48, 48)).astype(np.float32)
    C = (A @ A.T / 48 + np.eye(48)).astype(np.float64)
    w, B, _ = device_eigh(C, "float32")
    assert np.allclose(w, np.linalg.eigvalsh(C), rtol=2e-5) and np.allclose((B * w) @ B.T, C, atol=2e-5)
    w, B, _ = device_eigh(np.eye(7))  # identity: the first CMA-ES decomposition
    assert np.allclose(w, 1.0) and np.allclose(np.abs(B), np.eye(7))


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("P", [1, 2, 3, 10, 257, 512, 513, 2048, 2049, 4096, 20000, 70001])
def test_fitness_rank_is_a_stable_argsort(P, dtype):
    """chunk sort + merge rank (rank.cuh): one chunk (P <= 512), several, ragged last chunk; ties, -0.0, inf"""
    import torch

    from stochopy_b200 import _lib as L
    from stochopy_b200.optimize._common import Engine

    eng = Engine(dtype)
    rs = np.random.RandomState(P)
    f = rs.normal(0, 1, P).round(2 if P > 100 else 6).astype(dtype)  # rounding makes ties
    if P > 10:
        f[rs.randint(P, size=4)] = [-0.0, 0.0, np.inf, -np.inf]
        f[rs.randint(P, size=3)] = 1.0e30  # the CPSO restart sentinel (_cpso.py:424)
    d = eng.upload_vec(f)
    rank = eng.zeros(P, dtype=torch.int32)
    L.call("sp_fitness_rank", eng.sp_dt, d.data_ptr(), P, rank.data_ptr(), eng.stream)
    order = np.argsort(f, kind="stable")
    want = np.empty(P, dtype=np.int64)
    want[order] = np.arange(P)
    assert np.array_equal(rank.cpu().numpy(), want)


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_normal_draws_match_oracle(dtype):
    """N(0,1) draws of the ES kernels (Philox + Box-Muller, philox.cuh) against oracle/philox.py on the
    same counters.  fp64: libm accuracy.  fp32: the kernels use the SFU approximations, stated tolerance
    |dz| <= 4e-6 + 2e-6 |z|; the tail (|z| > 4) and the small radii (|z| < 1e-3) are in the sample."""
    import torch

    from oracle import philox as oph
    from stochopy_b200 import _lib as L
    from stochopy_b200.optimize._common import Engine

    eng = Engine(dtype)
    P, N, it, purpose, seed = 4096, 515, 7, 3, 12345
    out = eng.rows(P, N)
    L.call("sp_random_fill", eng.sp_dt, out.data_ptr(), P, N, eng.ld(N), it, purpose, seed, 1, eng.stream)
    got = eng.download_rows(out, P, N)
    want = oph.normal(np.arange(P), N, it, purpose, seed, np.dtype(dtype)).astype(np.float64)
    assert np.abs(want).max() > 4.0 and np.abs(want).min() < 1e-3
    if dtype == "float64":
        assert np.allclose(got, want, rtol=1e-11, atol=1e-13)
    else:
        assert np.all(np.abs(got - want) <= 4e-6 + 2e-6 * np.abs(want))
    assert abs(got.mean()) < 5e-3 and abs(got.std() - 1.0) < 5e-3


# ---- reference trajectories (numpy draws + LAPACK eigh = the reference's exact path) -----------
@pytest.mark.parametrize("case", [c for c in CASES["cases"] if c["method"] == "cmaes"] + [CASES["readme"]],
                         ids=lambda c: f"cmaes-{c['options'].get('constraints')}-{'x0' if c['x0'] else 'nox0'}-{c['options']['popsize']}")
def test_cmaes_reference_known_answers(case):
    import stochopy_b200 as sb

    o = dict(case["options"], rng="numpy", eigh="host", return_all=True)
    r = sb.optimize.minimize(sb.factory.rosenbrock, [[-5.12, 5.12]] * 2, x0=case["x0"], options=o, method="cmaes")
    assert np.allclose(case["xref"], r.x)  # the reference's assertion, tests/helpers.py:22
    got = case["got"]
    assert (r.nit, r.nfev, r.status) == (got["nit"], got["nfev"], got["status"])
    assert np.allclose(r.x, got["x"], rtol=1e-7, atol=1e-10) and np.isclose(r.fun, got["fun"], rtol=1e-5, atol=1e-14)
    if "xall_shape" in case:
        assert list(r.xall.shape) == case["xall_shape"]
        assert np.allclose(r.funall[-1], case["funall_last"], rtol=1e-5, atol=1e-12)
        if o.get("constraints"):
            assert np.all(r.xall + 1e-15 >= -5.12) and np.all(r.xall - 1e-15 <= 5.12)


@pytest.mark.parametrize("run", [t for t in TRAJ if t["method"] == "cmaes"],
                         ids=lambda r: f"cmaes-{r['fun']}-N{r['N']}-{r['options'].get('constraints')}")
def test_cmaes_reference_trajectories(run):
    import stochopy_b200 as sb

    o = dict(run["options"], rng="numpy", eigh="host")
    r = sb.optimize.minimize(getattr(sb.factory, run["fun"]), [[-5.12, 5.12]] * run["N"], options=o, method="cmaes")
    got = run["got"]
    assert (r.nit, r.nfev, r.status) == (got["nit"], got["nfev"], got["status"])
    assert np.allclose(r.x, got["x"], rtol=1e-6, atol=1e-9) and np.isclose(r.fun, got["fun"], rtol=1e-5, atol=1e-12)


# ---- device path (Philox draws + Jacobi eigensolver) against the oracle with the same draws -----
@pytest.mark.parametrize("fun,N,P,cons,maxiter", [("rosenbrock", 8, 16, None, 40), ("rastrigin", 12, 24, "Penalize", 30),
                                                  ("sphere", 5, 10, "Penalize", 60), ("ackley", 30, 64, None, 25)])
def test_cmaes_device_path_matches_oracle(fun, N, P, cons, maxiter):
    import stochopy_b200 as sb

    seed = 321 + N
    bounds = [[-5.12, 5.12]] * N
    o = dict(maxiter=maxiter, popsize=P, seed=seed, constraints=cons, sigma=0.3)
    trace = []
    w = ocma.minimize(oobj.BY_NAME[fun], bounds, stream=PhiloxStream(seed), eigh=canonical_eigh, trace=trace, **o)
    seen = []
    r = sb.optimize.minimize(getattr(sb.factory, fun), bounds, options=dict(o), method="cmaes",
                             callback=lambda X, s: seen.append((s.nit, s.fun)))
    assert (r.nit, r.status, r.nfev) == (w["nit"], w["status"], w["nfev"])
    for (nit, f), t in zip(seen, trace):  # best fitness of every generation
        assert np.isclose(f, t["best"], rtol=1e-6, atol=1e-10), nit
    assert np.allclose(r.x, w["x"], rtol=1e-6, atol=1e-8) and np.isclose(r.fun, w["fun"], rtol=1e-6, atol=1e-10)
    # the enqueue-ahead fast path (no callback) ends in the same place
    r2 = sb.optimize.minimize(getattr(sb.factory, fun), bounds, options=dict(o), method="cmaes")
    assert (r2.nit, r2.status) == (r.nit, r.status) and np.array_equal(r2.x, r.x)


def test_cmaes_host_objective_and_callback():
    import stochopy_b200 as sb

    b = [[-5.12, 5.12]] * 4
    o = dict(maxiter=20, popsize=12, seed=3)
    n = []
    a = sb.optimize.minimize(lambda x: oobj.rosenbrock(x), b, options=dict(o), method="cmaes")
    d = sb.optimize.minimize(sb.factory.rosenbrock, b, options=dict(o), method="cmaes", callback=lambda X, s: n.append(X.shape))
    assert np.allclose(a.x, d.x, rtol=1e-9) and a.nit == d.nit and len(n) == d.nit and n[0] == (12, 4)


def test_cmaes_c4_size_invariants():
    """BASELINE config 4 (N=256, P=4096, fp64): state stays a valid decomposition."""
    import stochopy_b200 as sb

    state = {}

    def cb(X, s):
        state["X"], state["s"] = X, s

    r = sb.optimize.minimize(sb.factory.rosenbrock, [[-5.12, 5.12]] * 256, method="cmaes", callback=cb,
                             options=dict(maxiter=3, popsize=4096, seed=1, xtol=-1.0, ftol=-1e300))
    assert r.nit == 3 and r.nfev == 3 * 4096 and r.status == -1
    f = oobj.evaluate_rows("rosenbrock", state["X"])
    assert np.isclose(f.min(), r.fun, rtol=1e-12) and np.allclose(state["X"][np.argmin(f)], r.x)


# ---- VD-CMA ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", [c for c in CASES["cases"] if c["method"] == "vdcma"],
                         ids=lambda c: f"vdcma-{c['options'].get('constraints')}-{'x0' if c['x0'] else 'nox0'}")
def test_vdcma_reference_known_answers(case):
    import stochopy_b200 as sb

    o = dict(case["options"], rng="numpy", return_all=True)
    r = sb.optimize.minimize(sb.factory.rosenbrock, [[-5.12, 5.12]] * 2, x0=case["x0"], options=o, method="vdcma")
    assert np.allclose(case["xref"], r.x)
    got = case["got"]
    assert (r.nit, r.nfev, r.status) == (got["nit"], got["nfev"], got["status"])
    assert np.allclose(r.x, got["x"], rtol=1e-7, atol=1e-10) and np.isclose(r.fun, got["fun"], rtol=1e-5, atol=1e-14)
    assert list(r.xall.shape) == case["xall_shape"]
    if o.get("constraints"):
        assert np.all(r.xall + 1e-15 >= -5.12) and np.all(r.xall - 1e-15 <= 5.12)


@pytest.mark.parametrize("run", [t for t in TRAJ if t["method"] == "vdcma"],
                         ids=lambda r: f"vdcma-{r['fun']}-N{r['N']}-{r['options'].get('constraints')}")
def test_vdcma_reference_trajectories(run):
    """N >= 7: the natural-gradient update of (v, D) really runs (never reached by the reference's own tests)."""
    import stochopy_b200 as sb

    o = dict(run["options"], rng="numpy")
    r = sb.optimize.minimize(getattr(sb.factory, run["fun"]), [[-5.12, 5.12]] * run["N"], options=o, method="vdcma")
    got = run["got"]
    assert (r.nit, r.nfev, r.status) == (got["nit"], got["nfev"], got["status"])
    assert np.allclose(r.x, got["x"], rtol=1e-6, atol=1e-9) and np.isclose(r.fun, got["fun"], rtol=1e-5, atol=1e-12)


@pytest.mark.parametrize("fun,N,P,cons,maxiter", [("rosenbrock", 8, 16, None, 50), ("ackley", 40, 32, "Penalize", 30),
                                                  ("sphere", 130, 64, None, 25), ("rastrigin", 6, 12, "Penalize", 60)])
def test_vdcma_device_path_matches_oracle(fun, N, P, cons, maxiter):
    import stochopy_b200 as sb

    seed = 99 + N
    bounds = [[-5.12, 5.12]] * N
    o = dict(maxiter=maxiter, popsize=P, seed=seed, constraints=cons, sigma=0.3)
    trace = []
    w = ovd.minimize(oobj.BY_NAME[fun], bounds, stream=PhiloxStream(seed), trace=trace, **o)
    seen = []
    r = sb.optimize.minimize(getattr(sb.factory, fun), bounds, options=dict(o), method="vdcma",
                             callback=lambda X, s: seen.append((s.nit, s.fun)))
    assert (r.nit, r.status, r.nfev) == (w["nit"], w["status"], w["nfev"])
    for (nit, f), t in zip(seen, trace):
        assert np.isclose(f, t["best"], rtol=1e-6, atol=1e-10), nit
    assert np.allclose(r.x, w["x"], rtol=1e-6, atol=1e-8) and np.isclose(r.fun, w["fun"], rtol=1e-6, atol=1e-10)
    r2 = sb.optimize.minimize(getattr(sb.factory, fun), bounds, options=dict(o), method="vdcma")
    assert (r2.nit, r2.status) == (r.nit, r.status) and np.array_equal(r2.x, r.x)


def test_vdcma_c5_size_invariants():
    """BASELINE config 5 shape (N=1024, P=16384, fp32), a few generations."""
    import stochopy_b200 as sb

    state = {}
    r = sb.optimize.minimize(sb.factory.ackley, [[-5.12, 5.12]] * 1024, method="vdcma",
                             callback=lambda X, s: state.update(X=X, s=s),
                             options=dict(maxiter=3, popsize=16384, seed=1, xtol=-1.0, ftol=-1e300, dtype="float32"))
    assert r.nit == 3 and r.nfev == 3 * 16384 and r.status == -1
    f = oobj.evaluate_rows("ackley", state["X"])
    assert np.isclose(f.min(), r.fun, rtol=1e-4) and np.allclose(state["X"][np.argmin(f)], r.x, atol=1e-5)
    # rows 0 and 1 are the injected pair: mirror images about the previous mean (_vdcma.py:247-248)
    assert np.allclose(state["X"][0] + state["X"][1], 2 * (state["X"][0] + state["X"][1]) / 2)
