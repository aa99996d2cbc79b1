#!/usr/bin/env python
"""fp64 GEMM ceiling of this GPU ("fp64 peak is not in MEASURED_PEAKS.json --
measure a DGEMM probe on the box").  cuBLAS DGEMM through torch.matmul, square sizes, best of
several runs with CUDA events; also the C4-shaped products (4096x256 @ 256x256 and
256x2048 @ 2048x256) so the CMA-ES kernels have a library number beside them.
Usage: python profiles/dgemm_probe.py [out.json]"""
import json
import sys

import torch


def best_ms(fn, reps=10):
    for _ in range(3):
        fn()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    dev = torch.device("cuda", 0)
    out = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "square": {}, "c4_shapes": {}}
    for n in (2048, 4096, 8192):
        A = torch.randn(n, n, dtype=torch.float64, device=dev)
        B = torch.randn(n, n, dtype=torch.float64, device=dev)
        ms = best_ms(lambda: torch.matmul(A, B), reps=5 if n == 8192 else 10)
        out["square"][str(n)] = {"ms": ms, "tflops": 2.0 * n**3 / ms / 1e9}
    # C4: sampling GEMM (P x N)(N x N) and rank-mu covariance (N x mu)(mu x N), fp64
    P, N, mu = 4096, 256, 2048
    Z = torch.randn(P, N, dtype=torch.float64, device=dev)
    Bm = torch.randn(N, N, dtype=torch.float64, device=dev)
    Y = torch.randn(mu, N, dtype=torch.float64, device=dev)
    ms = best_ms(lambda: torch.matmul(Z, Bm), reps=50)
    out["c4_shapes"]["sample_4096x256x256"] = {"us": ms * 1e3, "tflops": 2.0 * P * N * N / ms / 1e9}
    ms = best_ms(lambda: torch.matmul(Y.t(), Y), reps=50)
    out["c4_shapes"]["cov_256x2048x256"] = {"us": ms * 1e3, "tflops": 2.0 * mu * N * N / ms / 1e9}
    w = torch.empty(N, dtype=torch.float64, device=dev)
    C = Y.t() @ Y / mu
    ms = best_ms(lambda: torch.linalg.eigh(C), reps=5)
    out["c4_shapes"]["cusolver_eigh_256"] = {"us": ms * 1e3}
    out["fp64_peak_tflops"] = max(v["tflops"] for v in out["square"].values())
    s = json.dumps(out, indent=1)
    print(s)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(s + "\n")


if __name__ == "__main__":
    main()
