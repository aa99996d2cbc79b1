"""Oracle-level parity AT THE BASELINE SIZES, fp64 and fp32.

The small-size tests of test_gpu_es.py never leave one 64x64 GEMM tile; here the CUDA path
runs BASELINE.json's configurations themselves and every generation is compared with the
oracle's step functions (oracle/cmaes.py, oracle/vdcma.py, oracle/de.py) fed the device's own
counter-based draws (oracle/philox.py reproduces them):

  C4  CMA-ES  N=256  P=4096   multi-tile sampling GEMM, split-K rank-mu covariance, Jacobi eigh
  C5  VD-CMA  N=1024 P=16384  row-tile sampling, weighted sums over the mu best, natural gradient
  C2  DE      N=128  P=65536  one generation of the headline kernel against ode.trial_population

fp32 tolerances are stated next to each assertion; the observed maxima are appended to
gpurun_out/size_parity.jsonl so they can be read back after a GPU run.
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import cmaes as ocma  # noqa: E402
from oracle import de as ode  # noqa: E402
from oracle import objectives as oobj  # noqa: E402
from oracle import vdcma as ovd  # noqa: E402
from oracle.common import select_sync  # noqa: E402
from oracle.streams import PhiloxStream  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
BOUND = 5.12


def note(**kw):
    try:
        os.makedirs(OUT, exist_ok=True)
        with open(os.path.join(OUT, "size_parity.jsonl"), "a") as f:
            f.write(json.dumps(kw, default=float) + "\n")
    except OSError:
        pass


def host(t, n=None):
    a = t.detach().to("cpu").numpy().astype(np.float64)
    return a if n is None else a[..., :n]


def maxerr(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))))


# ---- C4: CMA-ES --------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype,N,P,fun", [("float64", 256, 4096, "rosenbrock"), ("float32", 256, 4096, "rosenbrock"),
                                           ("float64", 130, 600, "rastrigin"), ("float32", 130, 600, "rastrigin")])
def test_cmaes_generation_steps_match_oracle(dtype, N, P, fun):
    """Every generation of the device run is replayed by the oracle's step functions from the
    device's own state before that generation (sample :232-237, update :272-298, decompose
    :301-309, converge :360-434 of cmaes/_cmaes.py).  Eigenvectors are compared through
    sign-invariant quantities (B D^2 B^T, invsqrtC), see SURVEY.md 7 "Eigenvector sign"."""
    import stochopy_b200 as sb

    f64 = dtype == "float64"
    seed, sigma0, gens = 4242, 0.3, 4
    dt = np.dtype(dtype)
    snaps = []

    def probe(it, bufs, c):
        snaps.append(dict(
            it=it, xmean=host(bufs["xmean"], N), xold=host(bufs["xold"], N), ps=host(bufs["ps"]), pc=host(bufs["pc"]),
            C=host(bufs["C"]), B=host(bufs["B"]), D=host(bufs["D"]), invsqrtC=host(bufs["invsqrtC"]),
            arx=host(bufs["arx"], N), arfit=host(bufs["arfit"]), rank=bufs["rank"].to("cpu").numpy().astype(np.int64),
            sigma=c.sigma, sigma_gen=c.sigma_gen, hsig=c.hsig, nfev=int(c.nfev), status=c.base.status,
            gbest_row=int(c.base.gbest_row), gfit=c.base.gfit, do_eig=c.do_eig))

    r = sb.optimize.minimize(getattr(sb.factory, fun), [[-BOUND, BOUND]] * N, method="cmaes",
                             options=dict(_probe=probe, maxiter=gens, popsize=P, seed=seed, sigma=sigma0, xtol=-1.0, ftol=-1e300,
                                          dtype=dtype))
    assert r.nit == gens and r.status == -1 and len(snaps) == gens

    stream = PhiloxStream(seed, dt)
    mu, w, mueff = ocma.selection_weights(P, 0.5)
    consts = ocma.strategy_constants(N, mueff)
    cc, cs, c1, cmu, damps, chind = consts
    xm, xs = np.zeros(N), np.full(N, BOUND)
    before = dict(xmean=stream.es_mean0(N).astype(np.float64), ps=np.zeros(N), pc=np.zeros(N), C=np.eye(N), B=np.eye(N),
                  D=np.ones(N), invsqrtC=np.eye(N), sigma=sigma0, nfev=0)
    besthist = np.zeros(gens)
    # tolerances: fp64 = north_star's 1e-6 relative with room to spare; fp32 = 24-bit arithmetic over
    # K = N (sampling) and K = mu (covariance) products, SFU Box-Muller draws (4e-6)
    t_x = 1e-10 if f64 else 3e-5     # arx, xmean (values O(1))
    t_c = 1e-11 if f64 else 2e-5     # C entries (diag ~1)
    t_i = 1e-8 if f64 else 5e-4      # invsqrtC and B D^2 B^T - C (eigendecomposition)
    t_p = 1e-8 if f64 else 2e-3      # ps, pc (divide by sigma, multiply by sqrt(mueff) ~ 30)
    worst = {}

    def chk(name, got, want, tol):
        e = maxerr(got, want)
        worst[name] = max(worst.get(name, 0.0), e)
        assert e <= tol, (name, s["it"], e, tol)

    for s in snaps:
        g = s["it"]
        Z = stream.es_z(g, P, N).astype(np.float64)
        arx_o = ocma.sample(before["xmean"], before["sigma"], before["B"], before["D"], Z)
        chk("arx", s["arx"], arx_o, t_x)
        assert s["sigma_gen"] == before["sigma"]
        # objective in the un-standardised space (_cmaes.py:168-173) on the DEVICE's rows
        Xu = s["arx"] * xs + xm
        f_o = oobj.evaluate_rows(fun, Xu)
        scale = oobj.term_magnitude(fun, Xu)
        ftol_rel = 1e-13 if f64 else 4e-6 * max(1.0, np.sqrt(N) / 4)
        assert np.all(np.abs(s["arfit"] - f_o) <= ftol_rel * scale), (g, np.max(np.abs(s["arfit"] - f_o) / scale))
        # ranking == stable argsort of the device's own fitness (np.argsort order, _cmaes.py:272)
        order = np.argsort(s["arfit"], kind="stable")
        want_rank = np.empty(P, dtype=np.int64)
        want_rank[order] = np.arange(P)
        assert np.array_equal(s["rank"], want_rank)
        assert s["gbest_row"] == order[0] and s["gfit"] == s["arfit"][order[0]]
        nfev = before["nfev"] + P
        xmean, xold, ps, pc, C, sigma, hsig = ocma.update(
            s["arx"], order, mu, w, before["xmean"], before["sigma"], before["ps"], before["pc"], before["C"],
            before["invsqrtC"], consts, mueff, nfev, P)
        chk("xmean", s["xmean"], xmean, t_x)
        chk("xold", s["xold"], xold, 0.0 if f64 else 1e-7)
        chk("ps", s["ps"], ps, t_p)
        chk("pc", s["pc"], pc, t_p)
        assert bool(s["hsig"]) == bool(hsig) and s["nfev"] == nfev
        assert abs(s["sigma"] - sigma) <= (1e-10 if f64 else 2e-5) * sigma, (g, s["sigma"], sigma)
        assert P > P / (c1 + cmu) / N / 10.0 and s["do_eig"] == 1  # these sizes decompose every generation (:301)
        Cs, B_o, D_o, inv_o = ocma.decompose(C)
        chk("C", s["C"], Cs, t_c)
        assert np.array_equal(s["C"], s["C"].T)  # symmetrised from the upper triangle (_cmaes.py:303)
        chk("D", s["D"], D_o, 1e-9 if f64 else 1e-4)  # sqrt of fp32 Jacobi eigenvalues (~N eps)
        assert np.all(np.diff(s["D"]) >= 0)
        chk("BD2Bt", (s["B"] * s["D"] ** 2) @ s["B"].T, s["C"], t_i)
        chk("BtB", s["B"].T @ s["B"], np.eye(N), t_i)
        chk("invsqrtC", s["invsqrtC"], inv_o, t_i)
        besthist[g - 1] = s["arfit"][order[0]]
        st = ocma.converge(g, N, gens, xmean, xold, besthist, s["arfit"], order, sigma, sigma0, int(10.0 + 30.0 * N / P),
                           pc, -1.0, -1e300, np.diag(Cs), B_o, D_o)
        assert (st if st is not None else -1000) == s["status"], (g, st, s["status"])
        before = dict(xmean=s["xmean"], ps=s["ps"], pc=s["pc"], C=s["C"], B=s["B"], D=s["D"], invsqrtC=s["invsqrtC"],
                      sigma=s["sigma"], nfev=nfev)
    note(test="cmaes_steps", dtype=dtype, N=N, P=P, **worst)


# ---- C5: VD-CMA --------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype,N,P,fun", [("float64", 1024, 16384, "ackley"), ("float32", 1024, 16384, "ackley"),
                                           ("float64", 300, 1000, "rosenbrock"), ("float32", 300, 1000, "rosenbrock")])
def test_vdcma_generation_steps_match_oracle(dtype, N, P, fun):
    """VD-CMA has no sign ambiguity, so each generation is replayed from the device's state
    before it: sampling + injection (_vdcma.py:239-249), weighted sums / sigma / pc (:290-315)
    and the restricted covariance update (:318-372, 426-458) of oracle/vdcma.py."""
    import stochopy_b200 as sb

    f64 = dtype == "float64"
    seed, sigma0, gens = 777, 0.3, 4
    dt = np.dtype(dtype)
    snaps = []

    def probe(it, bufs, c):
        snaps.append(dict(
            it=it, xmean=host(bufs["xmean"], N), xold=host(bufs["xold"], N), pc=host(bufs["pc"]), dx=host(bufs["dx"]),
            dvec=host(bufs["dvec"], N), vvec=host(bufs["vvec"], N), arx=host(bufs["arx"], N), ary=host(bufs["ary"], N),
            arfit=host(bufs["arfit"]), rank=bufs["rank"].to("cpu").numpy().astype(np.int64), sigma=c.sigma,
            sigma_gen=c.sigma_gen, vd_ps=c.vd_ps, nfev=int(c.nfev), status=c.base.status,
            gbest_row=int(c.base.gbest_row), gfit=c.base.gfit))

    r = sb.optimize.minimize(getattr(sb.factory, fun), [[-BOUND, BOUND]] * N, method="vdcma",
                             options=dict(_probe=probe, maxiter=gens, popsize=P, seed=seed, sigma=sigma0, xtol=-1.0, ftol=-1e300,
                                          dtype=dtype))
    assert r.nit == gens and r.status == -1 and len(snaps) == gens

    stream = PhiloxStream(seed, dt)
    mu, w, mueff = ocma.selection_weights(P, 0.5)
    cc, c1, cmu = ovd.strategy_constants(N, mueff)
    xs = np.full(N, BOUND)
    b = dict(xmean=stream.es_mean0(N).astype(np.float64), dvec=np.ones(N),
             vvec=stream.vd_v0(N).astype(np.float64) / np.sqrt(N), pc=np.zeros(N), dx=np.zeros(N), sigma=sigma0, ps=0.0)
    if not f64:  # the device divides in fp32
        b["vvec"] = (stream.vd_v0(N) / dt.type(np.sqrt(N))).astype(np.float64)
    t_y = 1e-11 if f64 else 2e-5     # ary / arx rows (O(1) values; Box-Muller 4e-6 |z|, z up to ~5)
    t_m = 1e-11 if f64 else 5e-6     # xmean (weighted mean of mu rows)
    t_v = 1e-9 if f64 else 2e-4      # dvec, vvec, pc: relative to max|.|
    worst = {}

    def chk(name, got, want, tol, rel=False):
        e = maxerr(got, want) / (float(np.max(np.abs(want))) if rel else 1.0)
        worst[name] = max(worst.get(name, 0.0), e)
        assert e <= tol, (name, s["it"], e, tol)

    for s in snaps:
        g = s["it"]
        Z = stream.es_z(g, P, N).astype(np.float64)
        ary = ovd.sample(Z, b["dvec"], b["vvec"])
        if g > 1:
            dy = ovd.injection(b["dx"], b["dvec"], b["vvec"], stream.vd_inject(g, N).astype(np.float64))
            ary[0], ary[1] = dy, -dy
        chk("ary", s["ary"], ary, t_y)
        chk("arx", s["arx"], b["xmean"] + b["sigma"] * ary, t_y)
        Xu = s["arx"] * xs
        f_o = oobj.evaluate_rows(fun, Xu)
        scale = oobj.term_magnitude(fun, Xu)
        ftol_rel = 1e-13 if f64 else 4e-6 * max(1.0, np.sqrt(N) / 4)
        assert np.all(np.abs(s["arfit"] - f_o) <= ftol_rel * scale), (g, np.max(np.abs(s["arfit"] - f_o) / scale))
        order = np.argsort(s["arfit"], kind="stable")
        want_rank = np.empty(P, dtype=np.int64)
        want_rank[order] = np.arange(P)
        assert np.array_equal(s["rank"], want_rank)
        assert s["gbest_row"] == order[0]
        # _vdcma.py:290-315 on the device's rows
        elite_x, elite_y = s["arx"][order[:mu]], s["ary"][order[:mu]]
        dx = w @ elite_x - w.sum() * b["xmean"]
        chk("dx", s["dx"], dx, t_m)
        chk("xmean", s["xmean"], b["xmean"] + dx, t_m)
        sigma, ps = b["sigma"], b["ps"]
        if g > 1:
            gap = int(want_rank[1]) - int(want_rank[0])
            ps += 0.3 * (gap / (P - 1.0) - ps)
            sigma *= np.exp(ps / np.sqrt(N))
            hsig = ps < 0.5
        else:
            hsig = True
        assert abs(s["sigma"] - sigma) <= 1e-12 * sigma and abs(s["vd_ps"] - ps) <= 1e-12
        pc = (1.0 - cc) * b["pc"]
        if hsig:
            pc = pc + np.sqrt(cc * (2.0 - cc) * mueff) * (w @ elite_y)
        chk("pc", s["pc"], pc, t_v, rel=True)
        dvec, vvec = ovd.adapt(elite_y, w, pc, b["dvec"], b["vvec"], c1, cmu, hsig)
        chk("dvec", s["dvec"], dvec, t_v, rel=True)
        chk("vvec", s["vvec"], vvec, t_v, rel=True)
        assert s["status"] == (-1 if g == gens else -1000) and s["nfev"] == g * P
        b = dict(xmean=s["xmean"], dvec=s["dvec"], vvec=s["vvec"], pc=s["pc"], dx=s["dx"], sigma=s["sigma"], ps=s["vd_ps"])
    note(test="vdcma_steps", dtype=dtype, N=N, P=P, **worst)


# ---- C2: DE at the headline size ----------------------------------------------------------------
@pytest.mark.parametrize("fun", ["rastrigin", "rosenbrock"])
def test_de_full_size_generation_matches_oracle(fun):
    """One generation of the headline kernel at P=65536, N=128, fp32 (BASELINE configs[1] and
    the headline) against oracle.de.trial_population / select_sync fed the same Philox draws:
    trial vectors, candidate fitness, selection, argmin (de/_de.py:314-351)."""
    import ctypes as C

    from gpu_util import DeRig
    from stochopy_b200 import _lib as L

    P, N, seed, it = 65536, 128, 31337, 2
    dt = np.dtype("float32")
    rs = np.random.RandomState(5)
    X = rs.uniform(-BOUND, BOUND, (P, N)).astype(dt)
    fit = oobj.evaluate_rows(fun, X.astype(np.float64)).astype(dt)
    gbest = X[int(np.argmin(fit))].copy()
    lo, hi = np.full(N, -BOUND), np.full(N, BOUND)
    rig = DeRig(X, fit, gbest, fun, "best1bin", None, 0.5, 0.9, lo, hi, seed=seed, dtype="float32", it0=it)
    assert L.load().sp_de_chainable(C.byref(rig.st)) != 0  # this is the pool kernel bench.py times
    rig.step(it)
    out = rig.get(it)
    r1, donors, irand, rep = PhiloxStream(seed, dt).de(it, P, N, 2, lo, hi, False)
    U = ode.trial_population(X, gbest, "best1bin", dt.type(0.5), dt.type(0.9), r1, donors, irand, rep, lo, hi)
    fU = oobj.evaluate_rows(fun, U.astype(np.float64))
    scale = oobj.term_magnitude(fun, U.astype(np.float64))
    err = np.max(np.abs(out["pfit"] - fU) / scale)
    assert err <= 4e-6 * np.sqrt(N) / 4, err  # fp32 fitness: 4e-6 sum|terms| sqrt(N)/4 (tests/gpu_util.py tol_for)
    Xo, po = X.copy(), fit.copy()
    select_sync(it, U, out["pfit"].astype(dt), gbest, Xo, po, 1000, 1e-8, 1e-8)
    # positions: fp32 tolerance of tests/gpu_util.py tol_for (2e-6 relative); the share of bit-identical
    # entries is recorded (the update uses non-contracted arithmetic in numpy's operation order)
    assert np.allclose(out["X"], Xo, rtol=2e-6, atol=1e-6)
    assert np.array_equal(out["pbestfit"], po.astype(np.float64))
    assert out["ctrl"].gbest_row == int(np.argmin(po)) and out["ctrl"].nit == it
    assert np.array_equal(out["gbest"], out["X"][int(np.argmin(po))])
    note(test="de_full_size", fun=fun, fit_err_rel=float(err), improved=int((po < fit).sum()),
         exact_share=float((out["X"] == Xo).mean()))
