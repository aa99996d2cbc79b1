"""Stage clocks of vd_update_kernel's single-CTA phase (profiling hook sp_debug_vd_clocks), in-run (no profiler)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.getcwd())
import torch  # noqa: E402

import stochopy_b200 as sb  # noqa: E402
from stochopy_b200 import _lib as L  # noqa: E402

off = dict(xtol=-1.0, ftol=-1.0e300)
for dt in ("float32", "float64"):
    sb.optimize.minimize(sb.factory.ackley, [[-5.12, 5.12]] * 1024, method="vdcma",
                         options=dict(maxiter=40, popsize=16384, seed=0, dtype=dt, **off))
    torch.cuda.synchronize()
    out = (C.c_longlong * 16)()
    lib = L.load()
    lib.sp_debug_vd_clocks.argtypes = [C.POINTER(C.c_longlong), C.c_int]
    lib.sp_debug_vd_clocks(out, 0 if dt == "float32" else 1)
    v = list(out)
    names = ["phase1", "loads+pre", "L1 reduce", "p/q vectors", "vq reduce", "ria/via (+reduce)", "svnn (+reduce)",
             "ngv/ngd (+reduce)", "apply + L6 pre", "L6 reduce", "tail stores"]
    mhz = 1965.0
    print(dt, "total cycles", v[11] - v[0], "= %.2f us" % ((v[11] - v[0]) / mhz))
    for i, n in enumerate(names):
        print("  %-20s %8d cycles  %6.2f us" % (n, v[i + 1] - v[i], (v[i + 1] - v[i]) / mhz))
    print("  timeline of the last generation (globaltimer): sample + rank %.2f us, wsum %.2f us, update %.2f us"
          % ((v[13] - v[12]) / 1e3, (v[14] - v[13]) / 1e3, (v[15] - v[14]) / 1e3))
