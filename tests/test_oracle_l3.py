"""CPU half of the level-3 gate: the oracle fed the DEVICE's counter-based draws
(oracle/philox.py reproduces them bit for bit) against the real reference's outcome
distribution.  What this pins without a GPU is the draw *distribution* of the Philox mode --
O(k) donors, Feistel LHS permutations, Box-Muller normals -- the one thing the step-level
parity tests cannot see.  tests/test_gpu_l3.py repeats it with the CUDA path."""
import numpy as np
import pytest

import l3_util
from oracle import cmaes as ocma
from oracle import de as ode
from oracle import objectives as oobj
from oracle import pso as opso
from oracle import vdcma as ovd
from oracle.streams import PhiloxStream

REF = l3_util.load()
SEEDS = 32  # half of the fixture's 64 to keep the CPU suite short; the GPU test uses all 64


def canonical_eigh(C):
    w, V = np.linalg.eigh(C)
    for j in range(V.shape[1]):
        if V[int(np.argmax(np.abs(V[:, j]))), j] < 0:
            V[:, j] = -V[:, j]
    return w, V


def oracle_run(cfg, seed):
    o = dict(cfg["options"])
    fun = oobj.BY_NAME[cfg["fun"]]
    b = [[-REF["bound"], REF["bound"]]] * cfg["N"]
    st = PhiloxStream(seed)
    m = cfg["method"]
    if m == "de":
        r = ode.minimize(fun, b, stream=st, **o)
    elif m == "pso":
        r = opso.minimize(fun, b, stream=st, competitivity=None, **o)
    elif m == "cpso":
        r = opso.minimize(fun, b, stream=st, **o)
    elif m == "cmaes":
        r = ocma.minimize(fun, b, stream=st, eigh=canonical_eigh, **o)
    else:
        r = ovd.minimize(fun, b, stream=st, **o)
    return [int(r["status"]), int(r["nit"]), int(r["nfev"]), float(r["fun"])]


@pytest.mark.parametrize("name", ["headline_de_rosenbrock", "de_rand1bin_random_sphere", "c3_pso_styblinski",
                                  "c3_cpso_styblinski", "c4_cmaes_rosenbrock", "c5_vdcma_ackley"])
def test_philox_oracle_matches_reference_distribution(name):
    cfg = REF["configs"][name]
    got = [oracle_run(cfg, 1000 + s) for s in range(SEEDS)]
    stats = l3_util.compare(name, cfg["runs"], got)
    print(stats)
