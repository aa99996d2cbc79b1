"""The oracle's model of the device generators against the published known-answer vectors of
Random123 (Salmon, Moraes, Dror, Shaw, SC'11; kat_vectors of the Random123 distribution) and
basic properties of the derived streams.  CPU only."""
import numpy as np
import pytest

from oracle import philox as px


@pytest.mark.parametrize("ctr,key,out", [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
])
def test_philox4x32_10_known_answers(ctr, key, out):
    seed = key[0] | (key[1] << 32)
    got = tuple(int(np.asarray(w).reshape(-1)[0]) for w in px.philox4x32(*ctr, seed, rounds=10))
    assert got == out


@pytest.mark.parametrize("ctr,key,out", [
    ((0, 0), 0, (0xFF1DAE59, 0x6CD10DF2)),
    ((0xFFFFFFFF, 0xFFFFFFFF), 0xFFFFFFFF, (0x2C3F628B, 0xAB4FD7AD)),
    ((0x243F6A88, 0x85A308D3), 0x13198A2E, (0xDD7CE038, 0xF62A4C12)),
])
def test_philox2x32_10_known_answers(ctr, key, out):
    got = tuple(int(np.asarray(w).reshape(-1)[0]) for w in px.philox2x32(*ctr, key))
    assert got == out


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_de_cross_uniform_pieces(dtype):
    """16-bit pieces: values are multiples of 2^-16 in [0, 1), fp32 and fp64 streams are the same numbers,
    different generations / seeds give different streams, and P(u <= CR) is CR to sampling accuracy."""
    u = px.de_cross_uniform(512, 130, 5, 1234, dtype)
    assert u.shape == (512, 130) and u.dtype == np.dtype(dtype)
    k = u.astype(np.float64) * 65536.0
    assert np.array_equal(k, np.round(k)) and k.min() >= 0 and k.max() <= 65535
    assert np.array_equal(u.astype(np.float64), px.de_cross_uniform(512, 130, 5, 1234, np.float64))
    assert not np.array_equal(u, px.de_cross_uniform(512, 130, 6, 1234, dtype))
    assert not np.array_equal(u, px.de_cross_uniform(512, 130, 5, 1235, dtype))
    for cr in (0.1, 0.5, 0.9):
        assert abs(float(np.mean(u <= dtype(cr))) - cr) < 4.0 * np.sqrt(cr * (1 - cr) / u.size) + 2.0**-16
    # no visible structure along rows or columns
    assert abs(np.corrcoef(u[:, :-1].ravel(), u[:, 1:].ravel())[0, 1]) < 0.02
    assert abs(np.corrcoef(u[:-1].ravel(), u[1:].ravel())[0, 1]) < 0.02


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_pso_uniform_pieces(dtype):
    """r1 / r2 of the PSO velocity update: 16-bit pieces of one Philox4x32-10 call per 4 columns -- multiples of
    2^-16 in [0, 1), the same numbers in fp32 and fp64, r1 and r2 uncorrelated, mean 1/2."""
    rows = np.arange(300, 812)
    r1, r2 = px.pso_uniforms(rows, 70, 9, 4321, dtype)
    for r in (r1, r2):
        assert r.shape == (512, 70) and r.dtype == np.dtype(dtype)
        k = r.astype(np.float64) * 65536.0
        assert np.array_equal(k, np.round(k)) and k.min() >= 0 and k.max() <= 65535
        assert abs(float(r.mean()) - 0.5) < 4.0 / np.sqrt(12.0 * r.size) + 2.0**-16
    a1, a2 = px.pso_uniforms(rows, 70, 9, 4321, np.float64)
    assert np.array_equal(r1.astype(np.float64), a1) and np.array_equal(r2.astype(np.float64), a2)
    assert abs(np.corrcoef(r1.ravel(), r2.ravel())[0, 1]) < 0.02
    b1, _ = px.pso_uniforms(rows, 70, 10, 4321, dtype)
    assert not np.array_equal(r1, b1)
    # a draw depends on (row, column, generation), not on which rows are asked for (sharded swarm)
    c1, c2 = px.pso_uniforms(rows[100:200], 70, 9, 4321, dtype)
    assert np.array_equal(c1, r1[100:200]) and np.array_equal(c2, r2[100:200])
