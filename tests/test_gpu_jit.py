"""8f-3: objectives compiled from CUDA source at run time stay on the device and follow
exactly the path of the same objective given as a Python callable (the reference's
per-individual contract, _common.py:27-106)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# separate roundings (no FMA contraction) so that a Python loop reproduces it bit for bit
SRC = """
__device__ real objective(const real* x, int n) {
  real s = 0;
  for (int i = 0; i + 1 < n; ++i) {
    const real a = x[i + 1] - x[i] * x[i];
    const real b = (real)1 - x[i];
    s = s + ((real)100 * (a * a) + b * b);
  }
  return s;
}
"""
SRC_EXACT = """
__device__ real objective(const real* x, int n) {
  double s = 0;
  for (int i = 0; i + 1 < n; ++i) {
    const double a = __dadd_rn((double)x[i + 1], -__dmul_rn((double)x[i], (double)x[i]));
    const double b = __dadd_rn(1.0, -(double)x[i]);
    s = __dadd_rn(s, __dadd_rn(__dmul_rn(100.0, __dmul_rn(a, a)), __dmul_rn(b, b)));
  }
  return (real)s;
}
"""


def py_rosen(x):
    s = 0.0
    for i in range(len(x) - 1):
        a = x[i + 1] - x[i] * x[i]
        b = 1.0 - x[i]
        s = s + (100.0 * (a * a) + b * b)
    return s


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("P,N", [(1, 2), (7, 33), (1000, 128), (257, 1000), (70000, 16)])
def test_jit_eval_matches_numpy(P, N, dtype):
    import stochopy_b200 as sb
    from stochopy_b200.optimize._common import Engine

    f = sb.jit_objective(SRC)
    eng = Engine(dtype)
    rs = np.random.RandomState(P + N)
    X = rs.uniform(-2, 2, (P, N)).astype(dtype)
    scale, shift = rs.uniform(0.5, 2.0, N).astype(dtype), rs.uniform(-1, 1, N).astype(dtype)
    dX, out = eng.upload_rows(X), eng.empty(P)
    f.evaluate_rows(eng, dX, P, N, out)
    X64 = X.astype(np.float64)
    want = (100.0 * (X64[:, 1:] - X64[:, :-1] ** 2) ** 2 + (1 - X64[:, :-1]) ** 2).sum(axis=1)
    tol = 1e-12 if dtype == "float64" else 2e-5
    assert np.allclose(out.cpu().numpy(), want, rtol=tol, atol=tol)
    ld = dX.shape[1]
    f.evaluate_rows(eng, dX, P, N, out, eng.upload_vec(scale, ld), eng.upload_vec(shift, ld))
    Y = (X * scale + shift).astype(np.float64)
    want = (100.0 * (Y[:, 1:] - Y[:, :-1] ** 2) ** 2 + (1 - Y[:, :-1]) ** 2).sum(axis=1)
    assert np.allclose(out.cpu().numpy(), want, rtol=tol * 10, atol=tol * 10)


def test_jit_objective_is_still_a_callable():
    import stochopy_b200 as sb

    f = sb.jit_objective(SRC_EXACT)
    x = np.array([0.3, -1.2, 0.7, 2.0])
    assert f(x) == py_rosen(x)
    with pytest.raises(ValueError):
        f(x, 1.0)


@pytest.mark.parametrize("method,opts", [
    ("de", dict(strategy="rand1bin", constraints="Random", updating="deferred")), ("pso", dict(updating="deferred")),
    ("cpso", dict(updating="deferred")), ("cmaes", dict(constraints="Penalize")), ("cmaes", {}), ("vdcma", {}), ("na", {}),
])
def test_jit_objective_follows_the_python_callable_path(method, opts):
    """Same generation kernels, same draws; the only difference is where fun is evaluated.
    The source uses explicit round-to-nearest ops, so the Python loop is bit-identical."""
    import stochopy_b200 as sb

    b = [[-2.0, 2.0]] * 5
    o = dict(opts, maxiter=25, popsize=12, seed=3)
    a = sb.optimize.minimize(sb.jit_objective(SRC_EXACT), b, method=method, options=dict(o))
    d = sb.optimize.minimize(py_rosen, b, method=method, options=dict(o))
    assert (a.nit, a.status, a.nfev) == (d.nit, d.status, d.nfev)
    assert np.array_equal(a.x, d.x) and a.fun == d.fun
