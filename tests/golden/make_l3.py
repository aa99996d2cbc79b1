"""Generate l3_reference.json: the REAL reference's outcome distribution over many seeds.

SURVEY.md 8c, level 3: the default device mode (rng="philox": counter-based draws, O(k)
donors, SFU Box-Muller, Feistel LHS permutations, Jacobi eigenvectors) cannot follow the
reference's MT19937 trajectories, so it is gated statistically -- same status/nit
distribution and final fun within a band over >= 32 seeds on scaled-down BASELINE configs.
This script runs the unmodified reference (numpy MT19937) for SEEDS seeds per config and
records (status, nit, nfev, fun) of every run; tests/test_gpu_l3.py runs the device path with
the same options and compares the two samples.

Run in the dev container only (needs /root/reference):

    python tests/golden/make_l3.py
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
import stochopy  # noqa: E402
from stochopy import factory  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SEEDS = 64
BOUND = 5.12

# scaled-down BASELINE configs (BASELINE.json configs[1..4] + the headline); method defaults
# of SURVEY.md 8d; deferred updating = the synchronous algorithm the device implements
CONFIGS = {
    "c2_de_rastrigin": dict(method="de", fun="rastrigin", N=10, options=dict(
        maxiter=300, popsize=48, strategy="best1bin", mutation=0.5, recombination=0.9, updating="deferred")),
    "headline_de_rosenbrock": dict(method="de", fun="rosenbrock", N=8, options=dict(
        maxiter=400, popsize=40, strategy="best1bin", mutation=0.5, recombination=0.9, updating="deferred")),
    "de_rand1bin_random_sphere": dict(method="de", fun="sphere", N=12, options=dict(
        maxiter=200, popsize=48, strategy="rand1bin", mutation=0.5, recombination=0.9, updating="deferred",
        constraints="Random")),
    "c3_pso_styblinski": dict(method="pso", fun="styblinski_tang", N=8, options=dict(
        maxiter=200, popsize=32, inertia=0.7298, cognitivity=1.49618, sociability=1.49618, updating="deferred")),
    "c3_cpso_styblinski": dict(method="cpso", fun="styblinski_tang", N=8, options=dict(
        maxiter=200, popsize=32, inertia=0.7298, cognitivity=1.49618, sociability=1.49618, competitivity=1.0,
        updating="deferred", constraints="Shrink")),
    "c4_cmaes_rosenbrock": dict(method="cmaes", fun="rosenbrock", N=10, options=dict(
        maxiter=400, popsize=24, sigma=0.1, muperc=0.5)),
    "c5_vdcma_ackley": dict(method="vdcma", fun="ackley", N=24, options=dict(
        maxiter=300, popsize=24, sigma=0.1, muperc=0.5)),
    "na_sphere": dict(method="na", fun="sphere", N=4, options=dict(maxiter=40, popsize=16, nrperc=0.5)),
}


def main():
    out = {"seeds": SEEDS, "bound": BOUND, "configs": {}}
    for name, cfg in CONFIGS.items():
        fun = getattr(factory, cfg["fun"])
        runs = []
        for seed in range(SEEDS):
            r = stochopy.optimize.minimize(fun, [[-BOUND, BOUND]] * cfg["N"], method=cfg["method"],
                                           options=dict(cfg["options"], seed=seed))
            runs.append([int(r.status), int(r.nit), int(r.nfev), float(r.fun)])
        a = np.array(runs)
        print(f"{name:28s} status {dict(zip(*np.unique(a[:, 0].astype(int), return_counts=True)))} "
              f"nit median {np.median(a[:, 1]):.0f} fun median {np.median(a[:, 3]):.3e} "
              f"IQR [{np.percentile(a[:, 3], 25):.3e}, {np.percentile(a[:, 3], 75):.3e}]")
        out["configs"][name] = dict(cfg, runs=runs)
    with open(os.path.join(HERE, "l3_reference.json"), "w") as f:
        json.dump(out, f)
    print("wrote l3_reference.json")


if __name__ == "__main__":
    main()
