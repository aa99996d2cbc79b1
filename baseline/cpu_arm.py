"""CPU arm of bench.py: the reference's own DE on the box's host cores.

Runs the UNMODIFIED reference (keurfonluu/stochopy 2.3.0, installed by
``baseline/fetch_ref.sh`` into the git-ignored ``baseline/_ref``) through its public
API, ``stochopy.optimize.minimize(fun, bounds, x0, method="de", options=...)``
(stochopy/optimize/_helpers.py:44-94 -> de/_de.py:13-301), with
``updating="deferred"`` -- the synchronous variant the GPU path implements
(de/_de.py:314-351).  Legs (SURVEY.md 8d "Reference CPU timing beside it"):

  serial      workers=1 (optimizer wrapper's plain loop, _common.py:79-80)
  loky        workers=-1, backend="loky"      (_common.py:38-43, 94-97)
  threading   workers=-1, backend="threading"
  raw         [fun(x) for x in X] at the full P -- the fun(x) contract's CPU ceiling

The reference's DE draws a (P-1) x P donor index matrix every generation
(de/_de.py:304-311): 34 GB of int64 at P=65536, so it is timed on population
samples of 4096 and 8192 rows (evals/s falls with P: the O(P^2) term).

When ``baseline/_ref`` is missing the oracle port (oracle/de.py) is timed instead
and the result says ``kind: "port"``.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")

N = 128
BOUND = 5.12


def load_reference():
    """The installed reference package, or None."""
    if not os.path.isdir(os.path.join(REF, "stochopy")):
        return None
    if REF not in sys.path:
        sys.path.insert(0, REF)
    try:
        import stochopy  # noqa: F401
        import stochopy.optimize  # noqa: F401
        from stochopy.factory import rosenbrock  # noqa: F401
    except Exception:
        return None
    return sys.modules["stochopy"]


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def de_options(P, gens, seed=0, **extra):
    o = dict(maxiter=gens + 1, popsize=P, mutation=0.5, recombination=0.9, strategy="best1bin", seed=seed,
             xtol=-1.0, ftol=-1.0e300, updating="deferred")
    o.update(extra)
    return o


def reference_de_rate(ref, P, steps, warmup, seed=0, **extra):
    """(evals/s, s per generation) of the reference's DE over `steps` generations after
    `warmup` untimed ones; generation boundaries are the reference's own callback calls
    (de/_de.py:242 after the initial population, :287 after every generation)."""
    from stochopy.factory import rosenbrock

    rs = np.random.RandomState(seed)
    x0 = rs.uniform(-BOUND, BOUND, (P, N))
    stamps = []
    ref.optimize.minimize(rosenbrock, [[-BOUND, BOUND]] * N, x0=x0, method="de",
                          options=de_options(P, steps + warmup, seed, **extra),
                          callback=lambda X, s: stamps.append(time.perf_counter()))
    assert len(stamps) >= steps + 1, (len(stamps), steps, warmup)
    t = stamps[-1] - stamps[-1 - steps]
    return P * steps / t, t / steps


def port_de_rate(P, steps, warmup, seed=0):
    """Fallback: the oracle port of the same algorithm (oracle/de.py)."""
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import de as ode
    from oracle import objectives as oobj

    rs = np.random.RandomState(seed)
    x0 = rs.uniform(-BOUND, BOUND, (P, N))
    stamps = []
    ode.minimize(oobj.rosenbrock, [[-BOUND, BOUND]] * N, x0=x0, maxiter=steps + warmup + 1, popsize=P,
                 mutation=0.5, recombination=0.9, strategy="best1bin", seed=seed, xtol=-1.0, ftol=-1.0e300,
                 updating="deferred", callback=lambda X, s: stamps.append(time.perf_counter()))
    t = stamps[-1] - stamps[-1 - steps]
    return P * steps / t, t / steps


def raw_eval_rate(ref, P=65536, seed=0):
    """[fun(x) for x in X] at the full population: what the per-individual fun(x)
    contract costs on one core with nothing else around it (_common.py:79-80)."""
    if ref is not None:
        from stochopy.factory import rosenbrock as fun
    else:
        from oracle.objectives import rosenbrock as fun
    X = np.random.RandomState(seed).uniform(-BOUND, BOUND, (P, N))
    [fun(x) for x in X[:256]]
    t0 = time.perf_counter()
    f = [fun(x) for x in X]
    dt = time.perf_counter() - t0
    assert len(f) == P
    return P / dt


def run_legs(steps, warmup, full=True, budget_s=150.0):
    """Time the legs; returns a dict with per-leg evals/s and the description strings.
    `full`: all legs (the --impl reference run); otherwise the serial P=4096 leg and the
    raw loop only (the cpu_baseline key of our own arm)."""
    ref = load_reference()
    cpus = os.cpu_count() or 1
    out = {"kind": "reference" if ref is not None else "port", "host_cpus": cpus, "cpu_model": cpu_model(),
           "legs": {}}
    t_start = time.perf_counter()
    if ref is None:
        r, per = port_de_rate(4096, steps, warmup)
        out["legs"]["port_serial_P4096"] = {"evals_per_s": r, "s_per_generation": per, "cores": 1, "generations": steps}
        out["best"] = ("port_serial_P4096", r, per, 1, 4096, steps)
        out["legs"]["raw_fun_loop_P65536"] = {"evals_per_s": raw_eval_rate(None), "cores": 1}
        return out

    r, per = reference_de_rate(ref, 4096, steps, warmup)
    out["legs"]["serial_P4096"] = {"evals_per_s": r, "s_per_generation": per, "cores": 1, "generations": steps}
    best = ("serial_P4096", r, per, 1, 4096, steps)
    if full:
        short = max(3, min(steps, 6))
        for name, kw in (("loky_P4096", dict(workers=-1, backend="loky")),
                         ("threading_P4096", dict(workers=-1, backend="threading"))):
            if time.perf_counter() - t_start > budget_s:
                break
            try:
                r2, per2 = reference_de_rate(ref, 4096, short, 1, **kw)
                out["legs"][name] = {"evals_per_s": r2, "s_per_generation": per2, "cores": cpus, "generations": short}
                if r2 > best[1]:
                    best = (name, r2, per2, cpus, 4096, short)
            except Exception as e:  # a broken joblib backend must not lose the run
                out["legs"][name] = {"error": repr(e)[:200]}
        if time.perf_counter() - t_start < budget_s:
            r3, per3 = reference_de_rate(ref, 8192, 3, 1)
            out["legs"]["serial_P8192"] = {"evals_per_s": r3, "s_per_generation": per3, "cores": 1, "generations": 3}
    out["legs"]["raw_fun_loop_P65536"] = {"evals_per_s": raw_eval_rate(ref), "cores": 1}
    out["best"] = best
    return out


if __name__ == "__main__":
    import json

    print(json.dumps(run_legs(6, 1), indent=1))
