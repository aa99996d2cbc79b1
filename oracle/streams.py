"""Oracle: the two random-draw providers (test infrastructure).

``MTStream`` consumes numpy's legacy MT19937 generator in the reference's exact
order (SURVEY.md 8c "RNG draw order per method"); ``PhiloxStream`` reproduces
the device's counter-based draws (oracle/philox.py).  Both expose the same
methods, one per place where the reference draws.
"""
import numpy as np

from . import philox as px


class MTStream:
    """np.random.seed(seed) + the reference's draw sequence."""

    kind = "mt"

    def __init__(self, seed):
        self.rs = np.random.RandomState(seed)

    # _common.py:109-120 -- uniform(size=(P,N)) then N permutations of P
    def lhs(self, P, N):
        jitter = self.rs.uniform(size=(P, N))
        perms = np.stack([self.rs.permutation(P) for _ in range(N)], axis=1)
        return jitter, perms

    # _de.py:250 then :306,311 then :340 then de/_constraints.py:24
    def de(self, it, P, N, k, lower, upper, repair):
        r1 = self.rs.rand(P, N)
        donors = np.empty((P - 1, P), dtype=np.int64)
        for i in range(P):
            donors[:, i] = self.rs.permutation(np.delete(np.arange(P), i))
        irand = self.rs.randint(N, size=P)
        rep = self.rs.uniform(lower, upper, (P, N)) if repair else None
        return r1, donors[:k], irand, rep

    # asynchronous DE draws one individual at a time (_de.py:373-388)
    def de_rand(self, P, N):
        return self.rs.rand(P, N)

    def de_one(self, i, P, N, k, lower, upper, repair):
        donors = self.rs.permutation(np.delete(np.arange(P), i))[:k]
        irand = self.rs.randint(N)
        return donors, irand

    def repair_one(self, lower, upper, N):
        return self.rs.uniform(lower, upper, (N,))

    # _cpso.py:262-263
    def pso(self, it, P, N):
        return self.rs.rand(P, N), self.rs.rand(P, N)

    # _cpso.py:422 -- rows are given in reset order
    def pso_restart(self, it, rows, N, lower, upper):
        return self.rs.uniform(lower, upper, (len(rows), N))

    # _cmaes.py:180 / _vdcma.py:181
    def es_mean0(self, N):
        return self.rs.uniform(-1.0, 1.0, N)

    # _vdcma.py:208
    def vd_v0(self, N):
        return self.rs.normal(0.0, 1.0, N)

    # _cmaes.py:234 (P calls of randn(N) == one randn(P,N)) / _vdcma.py:239
    def es_z(self, it, P, N):
        return self.rs.randn(P, N)

    # _vdcma.py:246
    def vd_inject(self, it, N):
        return self.rs.randn(N)

    # _na.py:299
    def na_uniform(self, it, i, j, low, high):
        return self.rs.uniform(low, high)


class PhiloxStream:
    """The device's draws: counter = (block, row, generation, purpose)."""

    kind = "philox"

    def __init__(self, seed, dtype=np.float64):
        self.seed = int(seed)
        self.dtype = np.dtype(dtype)

    def lhs(self, P, N):
        jitter = px.uniform(np.arange(P), N, 0, px.LHS_JITTER, self.seed, self.dtype)
        perms = np.stack([px.lhs_permutation(P, j, self.seed) for j in range(N)], axis=1)
        return jitter, perms

    def de(self, it, P, N, k, lower, upper, repair):
        r1 = px.de_cross_uniform(P, N, it, self.seed, self.dtype)
        irand, donors = px.de_indices(P, N, k, it, self.seed)
        rep = None
        if repair:
            u = px.uniform(np.arange(P), N, it, px.DE_REPAIR, self.seed, self.dtype)
            lo = np.asarray(lower, dtype=self.dtype)
            hi = np.asarray(upper, dtype=self.dtype)
            rep = lo + (hi - lo) * u
        return r1, donors, irand, rep

    def pso(self, it, P, N):
        return px.pso_uniforms(np.arange(P), N, it, self.seed, self.dtype)

    def pso_restart(self, it, rows, N, lower, upper):
        u = px.uniform(np.asarray(rows), N, it, px.PSO_RESTART, self.seed, self.dtype)
        lo = np.asarray(lower, dtype=self.dtype)
        hi = np.asarray(upper, dtype=self.dtype)
        return lo + (hi - lo) * u

    def es_mean0(self, N):
        u = px.uniform([0], N, 0, px.ES_MEAN0, self.seed, self.dtype)[0]
        return (self.dtype.type(2.0) * u - self.dtype.type(1.0)).astype(self.dtype)

    def vd_v0(self, N):
        return px.normal([0], N, 0, px.VD_V0, self.seed, self.dtype)[0]

    def es_z(self, it, P, N):
        return px.normal(np.arange(P), N, it, px.ES_Z, self.seed, self.dtype)

    def vd_inject(self, it, N):
        return px.normal([0], N, it, px.VD_INJECT, self.seed, self.dtype)[0]

    def na_uniform(self, it, i, j, low, high):
        u = px.uniform([i], j + 1, it, px.NA_WALK, self.seed, self.dtype)[0][j]
        return low + (high - low) * u
