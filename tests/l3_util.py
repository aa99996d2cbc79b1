"""Level-3 statistical gate: two samples of optimiser outcomes -- the real
reference's (tests/golden/l3_reference.json, numpy MT19937, recorded by make_l3.py) and a
counter-based Philox run of the same options -- must look like draws from one distribution.

Bands (stated here, asserted by `compare`), n = m = 64 seeds:
  status   total-variation distance between the two status histograms <= 0.25
           (binomial sd of a frequency difference at n=64 is <= 0.09)
  fun      two-sample Kolmogorov-Smirnov D <= 0.34  (alpha ~ 0.001 at n=m=64: 1.95 sqrt(2/64))
           and the Philox median inside the reference's [10 %, 90 %] range
  nit      KS D <= 0.34 and |median_philox - median_ref| <= max(3, 0.25 median_ref)
"""
import json
import os

import numpy as np

G = os.path.join(os.path.dirname(__file__), "golden")
KS_MAX = 0.34
TV_MAX = 0.25


def load():
    with open(os.path.join(G, "l3_reference.json")) as f:
        return json.load(f)


def ks(a, b):
    a, b = np.sort(np.asarray(a, dtype=np.float64)), np.sort(np.asarray(b, dtype=np.float64))
    grid = np.concatenate([a, b])
    fa = np.searchsorted(a, grid, side="right") / a.size
    fb = np.searchsorted(b, grid, side="right") / b.size
    return float(np.abs(fa - fb).max())


def compare(name, ref_runs, got_runs):
    """ref_runs / got_runs: rows [status, nit, nfev, fun].  Returns a dict of the statistics;
    raises AssertionError with all of them when a band is broken."""
    r, g = np.asarray(ref_runs, dtype=np.float64), np.asarray(got_runs, dtype=np.float64)
    codes = sorted(set(r[:, 0].astype(int)) | set(g[:, 0].astype(int)))
    tv = 0.5 * sum(abs((r[:, 0] == c).mean() - (g[:, 0] == c).mean()) for c in codes)
    d_fun, d_nit = ks(r[:, 3], g[:, 3]), ks(r[:, 1], g[:, 1])
    lo, hi = np.percentile(r[:, 3], [10.0, 90.0])
    med_g, med_r = float(np.median(g[:, 3])), float(np.median(r[:, 3]))
    nit_g, nit_r = float(np.median(g[:, 1])), float(np.median(r[:, 1]))
    stats = dict(config=name, tv_status=float(tv), ks_fun=d_fun, ks_nit=d_nit, fun_median=(med_g, med_r),
                 ref_p10_p90=(float(lo), float(hi)), nit_median=(nit_g, nit_r),
                 status_ref={int(c): int((r[:, 0] == c).sum()) for c in codes},
                 status_got={int(c): int((g[:, 0] == c).sum()) for c in codes})
    ok = (tv <= TV_MAX and d_fun <= KS_MAX and d_nit <= KS_MAX and lo - 1e-12 <= med_g <= hi + 1e-12
          and abs(nit_g - nit_r) <= max(3.0, 0.25 * nit_r))
    assert ok, stats
    # nfev bookkeeping is deterministic given nit: identical relation in both samples
    ratio_r, ratio_g = r[:, 2] / r[:, 1], g[:, 2] / g[:, 1]
    assert np.all(ratio_r == ratio_r[0]) and np.all(ratio_g == ratio_r[0]), (ratio_r[0], ratio_g[:4])
    return stats
