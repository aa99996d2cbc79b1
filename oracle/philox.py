"""Oracle: numpy model of the device random generator (test infrastructure).

The CUDA kernels draw from Philox4x32-10 (Salmon et al., SC'11) with
  key     = (seed & 0xffffffff, seed >> 32)
  counter = (block, row, generation, purpose)
(the DE crossover decisions: 16-bit pieces of Philox2x32-10, see de_cross_uniform)
so that a draw depends only on *what it is for*, never on the launch shape or
on how many GPUs share the work.  This file reproduces those draws bit for bit
(integers, uniforms) or to libm accuracy (Box-Muller normals) so the oracle's
step functions can be fed exactly what the device used.
The spec is mirrored in stochopy_b200/csrc/philox.cuh.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)

# purpose tags (counter word 3)
LHS_JITTER = 1
DE_CROSS = 2
DE_INDEX = 3
DE_REPAIR = 4
PSO_R1 = 5
PSO_R2 = 6
PSO_RESTART = 7
ES_Z = 8
VD_INJECT = 9
ES_MEAN0 = 10
VD_V0 = 11
NA_WALK = 12


# rounds per purpose: 10, except the ES normal draws (csrc/philox.cuh kEsZRounds)
ROUNDS_BY_PURPOSE = {8: 7}


def philox4x32(c0, c1, c2, c3, seed, rounds=None):
    """Vectorised Philox4x32-R (R = 10, or ROUNDS_BY_PURPOSE[c3] when c3 is a scalar purpose tag).
    Counter words broadcast; returns 4 uint64 arrays (<2^32)."""
    if rounds is None:
        rounds = ROUNDS_BY_PURPOSE.get(int(c3), 10) if np.ndim(c3) == 0 else 10
    c0, c1, c2, c3 = np.broadcast_arrays(
        *(np.asarray(c, dtype=np.uint64) & MASK32 for c in (c0, c1, c2, c3))
    )
    c0, c1, c2, c3 = (c.copy() for c in (c0, c1, c2, c3))
    k0 = int(seed) & 0xFFFFFFFF
    k1 = (int(seed) >> 32) & 0xFFFFFFFF
    for _ in range(rounds):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c0, c1, c2, c3 = (
            hi1 ^ c1 ^ np.uint64(k0),
            lo1,
            hi0 ^ c3 ^ np.uint64(k1),
            lo0,
        )
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def _blocks(ncols, per_block):
    return (ncols + per_block - 1) // per_block


def uniform(rows, ncols, gen, purpose, seed, dtype):
    """U[0,1) matrix (len(rows), ncols).  fp32: 24-bit, 4 columns per block;
    fp64: 53-bit, 2 columns per block."""
    rows = np.asarray(rows, dtype=np.uint64)[:, None]
    if np.dtype(dtype) == np.float32:
        nb = _blocks(ncols, 4)
        o = philox4x32(np.arange(nb)[None, :], rows, gen, purpose, seed)
        w = np.stack(o, axis=-1).reshape(len(rows), nb * 4)[:, :ncols]
        return ((w >> np.uint64(8)).astype(np.float32) * np.float32(2.0**-24)).astype(np.float32)
    nb = _blocks(ncols, 2)
    o = philox4x32(np.arange(nb)[None, :], rows, gen, purpose, seed)
    a = (o[0] << np.uint64(21)) | (o[1] >> np.uint64(11))
    b = (o[2] << np.uint64(21)) | (o[3] >> np.uint64(11))
    w = np.stack((a, b), axis=-1).reshape(len(rows), nb * 2)[:, :ncols]
    return w.astype(np.float64) * 2.0**-53


def normal(rows, ncols, gen, purpose, seed, dtype):
    """N(0,1) matrix by Box-Muller on the same blocks as ``uniform``.
    fp32 block -> (r0 cos, r0 sin, r1 cos, r1 sin); fp64 block -> (r cos, r sin).
    The fp32 variant is evaluated in float64 and rounded once (the device uses the SFU
    approximations; the stated tolerance of the comparison covers both)."""
    rows = np.asarray(rows, dtype=np.uint64)[:, None]
    if np.dtype(dtype) == np.float32:
        # fp32 definition of csrc/philox.cuh::normal_block: 23-bit fractions k 2^-23 from the LOW bits of
        # each word; radius from u = 1 - k 2^-23 in (0, 1], angle 2 pi t with t = k' 2^-23 - 0.5
        nb = _blocks(ncols, 4)
        o = philox4x32(np.arange(nb)[None, :], rows, gen, purpose, seed)
        f = np.float32
        frac = [((w & np.uint64(0x7FFFFF)).astype(np.float64) * 2.0**-23) for w in o]
        out = np.empty((len(rows), nb, 4), dtype=f)
        for h in (0, 1):
            r = np.sqrt(-2.0 * np.log1p(-frac[2 * h]))
            t = frac[2 * h + 1] - 0.5
            out[:, :, 2 * h] = (r * np.cos(2.0 * np.pi * t)).astype(f)
            out[:, :, 2 * h + 1] = (r * np.sin(2.0 * np.pi * t)).astype(f)
        return out.reshape(len(rows), nb * 4)[:, :ncols]
    nb = _blocks(ncols, 2)
    o = philox4x32(np.arange(nb)[None, :], rows, gen, purpose, seed)
    a = ((o[0] << np.uint64(21)) | (o[1] >> np.uint64(11))).astype(np.float64)
    b = ((o[2] << np.uint64(21)) | (o[3] >> np.uint64(11))).astype(np.float64)
    u1 = (a + 1.0) * 2.0**-53
    u2 = b * 2.0**-53
    r = np.sqrt(-2.0 * np.log(u1))
    out = np.empty((len(rows), nb, 2))
    out[:, :, 0] = r * np.cos(np.pi * (2.0 * u2))
    out[:, :, 1] = r * np.sin(np.pi * (2.0 * u2))
    return out.reshape(len(rows), nb * 2)[:, :ncols]


M2 = np.uint64(0xD256D193)


def philox2x32(c0, c1, key, rounds=10):
    """Vectorised Philox2x32-R (Salmon et al., SC'11): counter (c0, c1), 32-bit key bumped by W0 per round."""
    c0, c1 = np.broadcast_arrays(*(np.asarray(c, dtype=np.uint64) & MASK32 for c in (c0, c1)))
    c0, c1 = c0.copy(), c1.copy()
    k = int(key) & 0xFFFFFFFF
    for _ in range(rounds):
        p = M2 * c0
        c0, c1 = (p >> np.uint64(32)) ^ np.uint64(k) ^ c1, p & MASK32
        k = (k + W0) & 0xFFFFFFFF
    return c0, c1


def de_cross_uniform(P, N, gen, seed, dtype):
    """The DE crossover uniforms of the device (csrc/philox.cuh "DE crossover stream"): 16-bit pieces.
    Stream id (key, o0, o1) = words 0..2 of Philox4x32-10(counter (0, 0, gen, DE_CROSS), seed); per row and
    group of 4 columns (x, y) = Philox2x32-10(counter (row + o0, group ^ o1), key); column j takes piece j & 3 of
    (x >> 16, x & 0xffff, y >> 16, y & 0xffff); u = piece * 2^-16 (exact in fp32 and fp64)."""
    sk = philox4x32(0, 0, gen, DE_CROSS, seed, rounds=10)
    key, o0, o1 = (int(np.asarray(w).reshape(-1)[0]) for w in sk[:3])
    ng = _blocks(N, 4)
    rows = (np.arange(P, dtype=np.uint64)[:, None] + np.uint64(o0)) & MASK32
    groups = np.arange(ng, dtype=np.uint64)[None, :] ^ np.uint64(o1)
    x, y = philox2x32(rows, groups, key)
    lo = np.uint64(0xFFFF)
    pieces = np.stack((x >> np.uint64(16), x & lo, y >> np.uint64(16), y & lo), axis=-1).reshape(P, ng * 4)[:, :N]
    return (pieces.astype(np.float64) * 2.0**-16).astype(dtype)


def pso_uniforms(rows, N, gen, seed, dtype):
    """PSO velocity coefficients of the device (csrc/philox.cuh pso_r12): one Philox4x32-10 call per row and
    group of 4 columns, counter (group, row, gen, PSO_R1); column 4g+p: r1 = (w_p >> 16) 2^-16, r2 = (w_p & 0xffff) 2^-16."""
    rows = np.asarray(rows, dtype=np.uint64)[:, None]
    ng = _blocks(N, 4)
    o = philox4x32(np.arange(ng)[None, :], rows, gen, PSO_R1, seed, rounds=10)
    w = np.stack(o, axis=-1).reshape(len(rows), ng * 4)[:, :N]
    r1 = ((w >> np.uint64(16)).astype(np.float64) * 2.0**-16).astype(dtype)
    r2 = ((w & np.uint64(0xFFFF)).astype(np.float64) * 2.0**-16).astype(dtype)
    return r1, r2


def _mulhi(word, n):
    return (word * np.uint64(n)) >> np.uint64(32)


def de_indices(P, N, k, gen, seed):
    """Per individual: the forced crossover column and k distinct donors != i.
    Block 0 -> (irand, d0, d1, d2); block 1 -> (d3, d4, -, -).
    Donor t is drawn from [0, P-1-t) and shifted past the already excluded
    indices in ascending order (uniform over the remaining rows)."""
    rows = np.arange(P, dtype=np.uint64)
    a = philox4x32(0, rows, gen, DE_INDEX, seed)
    b = philox4x32(1, rows, gen, DE_INDEX, seed)
    irand = _mulhi(a[0], N).astype(np.int64)
    words = [a[1], a[2], a[3], b[0], b[1]]
    donors = np.empty((k, P), dtype=np.int64)
    excl = [np.arange(P, dtype=np.int64)]
    for t in range(k):
        r = _mulhi(words[t], P - 1 - t).astype(np.int64)
        srt = np.sort(np.stack(excl, axis=0), axis=0)
        for e in srt:
            r = r + (r >= e)
        donors[t] = r
        excl.append(r)
    return irand, donors


def _mix32(h):
    h = h & MASK32
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x85EBCA6B)) & MASK32
    h ^= h >> np.uint64(13)
    h = (h * np.uint64(0xC2B2AE35)) & MASK32
    h ^= h >> np.uint64(16)
    return h


def lhs_permutation(P, col, seed):
    """Keyed bijection of [0,P): 4-round Feistel on 2*half bits + cycle walking."""
    bits = max(2, int(P - 1).bit_length())
    half = (bits + 1) // 2
    mask = np.uint64((1 << half) - 1)
    k0 = np.uint64(int(seed) & 0xFFFFFFFF)
    k1 = np.uint64((int(seed) >> 32) & 0xFFFFFFFF)
    colk = np.uint64((int(col) * 0x9E3779B1) & 0xFFFFFFFF)

    def feistel(v):
        L = v >> np.uint64(half)
        R = v & mask
        for rnd in range(4):
            f = _mix32(R ^ colk ^ (k0 if rnd % 2 == 0 else k1) ^ np.uint64((rnd + 1) * 0x7F4A7C15 & 0xFFFFFFFF))
            L, R = R, (L ^ f) & mask
        return (L << np.uint64(half)) | R

    v = feistel(np.arange(P, dtype=np.uint64))
    bad = v >= np.uint64(P)
    while bad.any():
        v[bad] = feistel(v[bad])
        bad = v >= np.uint64(P)
    return v.astype(np.int64)
