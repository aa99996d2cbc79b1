"""Level-3 gate on the device: the default mode (rng="philox", device
eigensolver) against the real reference's outcome distribution over 64 seeds per scaled-down
BASELINE config (tests/golden/l3_reference.json, generated from the unmodified reference by
tests/golden/make_l3.py).  Bands: tests/l3_util.py."""
import json
import os

import pytest

import l3_util

pytestmark = pytest.mark.gpu

REF = l3_util.load()
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("name", sorted(REF["configs"]))
def test_device_philox_matches_reference_distribution(name, dtype):
    import stochopy_b200 as sb

    cfg = REF["configs"][name]
    if dtype == "float32" and name in ("c4_cmaes_rosenbrock", "c5_vdcma_ackley", "de_rand1bin_random_sphere", "na_sphere"):
        pytest.skip("fp32 cannot resolve this config's 1e-8 ftol / 1e-6 final values: fp64 only")
    fun = getattr(sb.factory, cfg["fun"])
    b = [[-REF["bound"], REF["bound"]]] * cfg["N"]
    got = []
    for s in range(REF["seeds"]):
        r = sb.optimize.minimize(fun, b, method=cfg["method"], options=dict(cfg["options"], seed=1000 + s, dtype=dtype))
        got.append([int(r.status), int(r.nit), int(r.nfev), float(r.fun)])
    stats = l3_util.compare(name, cfg["runs"], got)
    try:
        os.makedirs(OUT, exist_ok=True)
        with open(os.path.join(OUT, "l3_stats.jsonl"), "a") as f:
            f.write(json.dumps(dict(stats, dtype=dtype), default=str) + "\n")
    except OSError:
        pass
