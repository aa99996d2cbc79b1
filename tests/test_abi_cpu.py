"""CPU-side checks of the boundary: the C-ABI library loads and exports every
symbol include/stochopy_b200.h declares; ctypes mirrors match; the front-ends
raise the reference's exception types before touching a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import stochopy_b200
from stochopy_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "stochopy_b200.h")).read()


def declared():
    return sorted(set(re.findall(r"^(?:int|int64_t|const char\*)\s+(sp_\w+)\(", HEADER, flags=re.M)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    names = declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(L.SIGNATURES) == names  # ctypes table and header agree
    assert lib.sp_abi_version() == int(re.search(r"#define SP_ABI_VERSION (\d+)", HEADER).group(1))
    assert lib.sp_scratch_bytes() > 0 and lib.sp_launch_count() >= 0


def test_struct_mirrors():
    assert C.sizeof(L.Ctrl) == 64
    assert L.SP_RUNNING == int(re.search(r"#define SP_RUNNING \((-?\d+)\)", HEADER).group(1))
    for name, val in L.OBJECTIVES.items():
        assert re.search(rf"SP_OBJ_{name.upper()} = {val}\b", HEADER), name
    for name, val in L.DE_STRATEGIES.items():
        assert re.search(rf"SP_DE_{name.upper()} = {val}\b", HEADER), name


def test_bad_arguments_fail_in_the_library_without_a_device():
    lib = L.load()
    assert lib.sp_eval(99, 0, None, 1, 1, 4, None, None, None, None) < 0
    assert b"sp_eval" in lib.sp_last_error()
    st = L.DeState()
    assert lib.sp_de_generation(C.byref(st), 2, None) < 0


@pytest.mark.parametrize("method", ["de", "pso", "cpso", "cmaes", "vdcma", "na"])
def test_front_end_validation_matches_reference(method):
    m = stochopy_b200.optimize.minimize
    b = [[-1.0, 1.0]] * 2
    with pytest.raises(TypeError):
        m(None, b, method=method)
    with pytest.raises(ValueError):
        m(lambda x: 0.0, [-1.0, 1.0], method=method)
    with pytest.raises(ValueError):
        m(lambda x: 0.0, b, method=method, callback=3)
    if method in ("de", "pso", "cpso", "na"):
        with pytest.raises(ValueError):
            m(lambda x: 0.0, b, method=method, options={"popsize": 1})
        with pytest.raises(ValueError):
            m(lambda x: 0.0, b, x0=np.zeros(2), method=method)
    if method in ("de", "pso", "cpso"):
        with pytest.raises(ValueError):
            m(lambda x: 0.0, b, method=method, options={"updating": "sometimes"})
        with pytest.raises(KeyError):
            m(lambda x: 0.0, b, method=method, options={"constraints": "Nope"})
    if method in ("cmaes", "vdcma"):
        with pytest.raises(ValueError):
            m(lambda x: 0.0, b, method=method, options={"sigma": 0.0})
        with pytest.raises(ValueError):
            m(lambda x: 0.0, b, method=method, options={"muperc": 0.0})
    if method == "de":
        with pytest.raises(KeyError):
            m(lambda x: 0.0, b, method="de", options={"strategy": "rand3bin"})
        with pytest.raises(ValueError):
            m(lambda x: 0.0, b, method="de", options={"mutation": 2.5})


def test_unknown_method_is_a_keyerror():
    with pytest.raises(KeyError):
        stochopy_b200.optimize.minimize(lambda x: 0.0, [[-1, 1]], method="simplex")


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(L.EngineError):
        stochopy_b200.optimize.minimize(stochopy_b200.factory.sphere, [[-1, 1]] * 2, method="de")


def test_result_type():
    r = stochopy_b200.optimize.OptimizeResult(x=np.ones(2), fun=1.0, xall=np.zeros((2, 2, 2)))
    assert r.fun == 1.0 and "xall" not in repr(r) and "fun" in dir(r)
    with pytest.raises(AttributeError):
        r.nothing


def test_jit_objective_compiles_without_a_device():
    """8f-3: NVRTC turns the user's CUDA source into an sm_100a cubin here (no GPU needed);
    compile errors come back with the compiler log."""
    ok = stochopy_b200.jit_objective(
        "__device__ real objective(const real* x, int n) { real s = 0; for (int i = 0; i < n; ++i) s += x[i] * x[i]; return s; }")
    assert ok.check("float64") > 1000 and ok.check("float32") > 1000
    bad = stochopy_b200.jit_objective("__device__ real objective(const real* x, int n) { return undefined_name; }")
    with pytest.raises(L.EngineError, match="undefined"):
        bad.check()
    with pytest.raises(ValueError):
        stochopy_b200.jit_objective("int f();")


def test_sizes_of_state_structs_match_the_header_layout():
    # pointers / int64 / double are 8-byte aligned: the ctypes mirrors must not be packed differently
    assert C.sizeof(L.PsoState) % 8 == 0 and C.sizeof(L.DeState) % 8 == 0
    assert L.PsoState.peers.offset == L.PsoState.mailbox.offset + 8
