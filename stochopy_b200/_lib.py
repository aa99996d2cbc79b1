"""ctypes binding of libstochopy_b200.so (the C ABI in include/stochopy_b200.h).

The shared library is built in-tree by ``__graft_entry__.build()`` (or
``make -C stochopy_b200/csrc``).  There is no CPU fallback: if the library or a
CUDA device is missing, loading / calling fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libstochopy_b200.so")

SP_RUNNING = -1000
SP_STATUS_PEER_TIMEOUT = -900
SP_STATUS_RESTART_PENDING = -901
SP_STATUS_INTERNAL = -902
PEER_HANDLE_BYTES = 64
SP_F32, SP_F64 = 0, 1
OBJECTIVES = {
    "ackley": 0,
    "griewank": 1,
    "quartic": 2,
    "rastrigin": 3,
    "rosenbrock": 4,
    "sphere": 5,
    "styblinski_tang": 6,
}
SP_OBJ_HOST = 100
DE_STRATEGIES = {"rand1bin": 0, "rand2bin": 1, "best1bin": 2, "best2bin": 3}
DE_DONORS = {"rand1bin": 3, "rand2bin": 5, "best1bin": 2, "best2bin": 4}
CONS_NONE, CONS_RANDOM, CONS_SHRINK, CONS_PENALIZE = 0, 1, 2, 3
# Philox purpose tags (csrc/philox.cuh)
PURPOSE_ES_MEAN0, PURPOSE_VD_V0 = 10, 11

vp = C.c_void_p


class Ctrl(C.Structure):
    """sp_ctrl (64 bytes, device resident; this is its host mirror)."""

    _fields_ = [
        ("status", C.c_int32),
        ("nit", C.c_int32),
        ("done_blocks", C.c_uint32),
        ("flag", C.c_int32),
        ("gbest_row", C.c_int64),
        ("gfit", C.c_double),
        ("dist", C.c_double),
        ("aux", C.c_double * 3),
    ]


class DeState(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("objective", C.c_int32), ("strategy", C.c_int32), ("constraint", C.c_int32),
        ("P", C.c_int64), ("N", C.c_int32), ("maxiter", C.c_int32), ("ld", C.c_int64),
        ("F", C.c_double), ("CR", C.c_double), ("xtol", C.c_double), ("ftol", C.c_double),
        ("seed", C.c_uint64),
        ("X", vp * 2), ("pbestfit", vp), ("pfit", vp), ("gbest", vp), ("lower", vp), ("upper", vp),
        ("ctrl", vp), ("scratch", vp),
        ("r1", vp), ("donors", vp), ("irand", vp), ("repair", vp),
    ]


class PsoState(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("objective", C.c_int32), ("constraint", C.c_int32), ("pad_", C.c_int32),
        ("P", C.c_int64), ("N", C.c_int32), ("maxiter", C.c_int32), ("ld", C.c_int64),
        ("w", C.c_double), ("c1", C.c_double), ("c2", C.c_double), ("xtol", C.c_double), ("ftol", C.c_double),
        ("gamma", C.c_double), ("delta", C.c_double),
        ("seed", C.c_uint64),
        ("X", vp), ("V", vp), ("pbest", vp), ("pbestfit", vp), ("pfit", vp), ("gbest", vp),
        ("lower", vp), ("upper", vp), ("ctrl", vp), ("scratch", vp),
        ("r1", vp), ("r2", vp),
        ("row0", C.c_int64), ("P_total", C.c_int64), ("xch", vp), ("shard", C.c_int32), ("pad2_", C.c_int32),
        ("world", C.c_int32), ("rank", C.c_int32), ("mailbox", vp), ("peers", vp), ("chain_rows", vp),
    ]


class EsCtrl(C.Structure):
    """sp_es_ctrl: control block of the evolution-strategy methods (cmaes, vdcma)."""

    _fields_ = [
        ("base", Ctrl),
        ("sigma", C.c_double), ("sigma_gen", C.c_double), ("ps_norm", C.c_double), ("vd_ps", C.c_double),
        ("nfev", C.c_int64), ("eigeneval", C.c_int64),
        ("hsig", C.c_int32), ("do_eig", C.c_int32), ("inject", C.c_int32), ("validfitval", C.c_int32),
        ("iniphase", C.c_int32), ("hist_len", C.c_int32), ("sweeps", C.c_int32), ("pad_", C.c_int32),
        ("aux", C.c_double * 16),
    ]


class CmaState(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("objective", C.c_int32), ("constraint", C.c_int32), ("N", C.c_int32),
        ("P", C.c_int64), ("ld", C.c_int64),
        ("mu", C.c_int32), ("maxiter", C.c_int32), ("ilim", C.c_int32), ("hist_cap", C.c_int32),
        ("cc", C.c_double), ("cs", C.c_double), ("c1", C.c_double), ("cmu", C.c_double), ("damps", C.c_double),
        ("chind", C.c_double), ("mueff", C.c_double), ("xtol", C.c_double), ("ftol", C.c_double),
        ("insigma", C.c_double),
        ("seed", C.c_uint64),
        ("xmean", vp), ("xold", vp), ("pc", vp), ("ps", vp), ("C", vp), ("B", vp), ("D", vp),
        ("invsqrtC", vp), ("arx", vp), ("arfit", vp), ("Z", vp), ("weights", vp), ("xscale", vp), ("xshift", vp),
        ("besthist", vp), ("work", vp), ("rank", vp), ("bnd_weights", vp), ("dfithist", vp), ("ctrl", vp),
        ("scratch", vp),
        ("host_z", C.c_int32), ("host_eigh", C.c_int32),
    ]


class VdState(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("objective", C.c_int32), ("constraint", C.c_int32), ("N", C.c_int32),
        ("P", C.c_int64), ("ld", C.c_int64),
        ("mu", C.c_int32), ("maxiter", C.c_int32), ("ilim", C.c_int32), ("hist_cap", C.c_int32),
        ("cc", C.c_double), ("c1", C.c_double), ("cmu", C.c_double), ("mueff", C.c_double), ("wsum", C.c_double),
        ("xtol", C.c_double), ("ftol", C.c_double), ("insigma", C.c_double),
        ("seed", C.c_uint64),
        ("xmean", vp), ("xold", vp), ("dx", vp), ("pc", vp), ("dvec", vp), ("vvec", vp), ("vn", vp), ("diagC", vp),
        ("dy", vp), ("ginj", vp), ("arx", vp), ("ary", vp), ("yvn", vp), ("arfit", vp), ("weights", vp),
        ("xscale", vp), ("xshift", vp), ("besthist", vp), ("work", vp), ("rank", vp), ("bnd_weights", vp),
        ("dfithist", vp), ("ctrl", vp), ("scratch", vp),
        ("host_z", C.c_int32), ("lean", C.c_int32),
    ]


_i, _i64, _d, _u64 = C.c_int, C.c_int64, C.c_double, C.c_uint64

# name -> (restype, argtypes); mirrors include/stochopy_b200.h one to one
SIGNATURES = {
    "sp_abi_version": (_i, []),
    "sp_last_error": (C.c_char_p, []),
    "sp_device_info": (_i, [_i, C.POINTER(_i), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i)]),
    "sp_launch_count": (_i64, []),
    "sp_scratch_bytes": (_i64, []),
    "sp_eval": (_i, [_i, _i, vp, _i64, _i, _i64, vp, vp, vp, vp]),
    "sp_lhs_init": (_i, [_i, vp, _i64, _i, _i64, vp, vp, _u64, vp, vp, vp]),
    "sp_lhs_init_shard": (_i, [_i, vp, _i64, _i, _i64, vp, vp, _u64, _i64, _i64, vp]),
    "sp_gbest_reduce": (_i, [_i, vp, _i, _i, _i64, vp, vp, _i, _i, _d, _d, vp]),
    "sp_select_sync": (_i, [_i, _i, _i, _d, _d, vp, vp, vp, vp, _i64, _i, _i64, _i, vp, vp, vp, vp]),
    "sp_best_init": (_i, [_i, vp, vp, _i64, _i, _i64, vp, vp, vp, vp]),
    "sp_de_generation": (_i, [C.POINTER(DeState), _i, vp]),
    "sp_de_generation_chained": (_i, [C.POINTER(DeState), _i, _i, vp]),
    "sp_de_chainable": (_i, [C.POINTER(DeState)]),
    "sp_de_propose": (_i, [C.POINTER(DeState), _i, vp]),
    "sp_de_run": (_i, [C.POINTER(DeState), _i, _i, vp]),
    "sp_pso_generation": (_i, [C.POINTER(PsoState), _i, vp]),
    "sp_pso_propose": (_i, [C.POINTER(PsoState), _i, vp]),
    "sp_cpso_restart_plan": (_i, [C.POINTER(PsoState), _i, vp, vp]),
    "sp_cpso_restart_apply": (_i, [C.POINTER(PsoState), _i, vp, vp, vp]),
    "sp_cpso_restart": (_i, [C.POINTER(PsoState), _i, vp, vp]),
    "sp_cpso_radius": (_i, [C.POINTER(PsoState), _i, vp]),
    "sp_cpso_decide": (_i, [C.POINTER(PsoState), _i, vp]),
    "sp_pso_chain_scalars": (_i64, [_i64]),
    "sp_pso_run_lazy": (_i, [C.POINTER(PsoState), _i, _i, vp]),
    "sp_cpso_restart_resume": (_i, [C.POINTER(PsoState), _i, vp, vp]),
    "sp_pso_run": (_i, [C.POINTER(PsoState), _i, _i, vp, vp]),
    "sp_peer_bytes": (_i64, [_i, _i, _i64, _i64]),
    "sp_peer_alloc": (_i, [_i64, C.POINTER(vp), vp]),
    "sp_peer_open": (_i, [vp, C.POINTER(vp)]),
    "sp_peer_close": (_i, [vp]),
    "sp_peer_free": (_i, [vp]),
    "sp_pso_run_sharded": (_i, [C.POINTER(PsoState), _i, _i, vp, vp]),
    "sp_jit_check": (_i, [C.c_char_p, _i, C.POINTER(_i64)]),
    "sp_jit_compile": (_i, [C.c_char_p, _i, C.POINTER(vp)]),
    "sp_jit_eval": (_i, [vp, _i, vp, _i64, _i, _i64, vp, vp, vp, vp]),
    "sp_jit_free": (_i, [vp]),
    "sp_random_fill": (_i, [_i, vp, _i64, _i, _i64, _i, _i, _u64, _i, vp]),
    "sp_fitness_rank": (_i, [_i, vp, _i64, vp, vp]),
    "sp_sym_eigh": (_i, [_i, vp, _i, vp, vp, vp, _i, vp, vp]),
    "sp_sym_eigh_work_scalars": (_i64, [_i]),
    "sp_cma_work_scalars": (_i64, [_i, _i64]),
    "sp_cma_generation": (_i, [C.POINTER(CmaState), _i, vp]),
    "sp_cma_sample": (_i, [C.POINTER(CmaState), _i, vp]),
    "sp_cma_update": (_i, [C.POINTER(CmaState), _i, vp]),
    "sp_cma_finish_generation": (_i, [C.POINTER(CmaState), _i, vp]),
    "sp_cma_run": (_i, [C.POINTER(CmaState), _i, _i, vp]),
    "sp_na_append": (_i, [_i, vp, _i64, _i64, vp, _i64, _i, _i64, vp]),
    "sp_na_cells": (_i, [vp, _i64, C.c_int32, vp, vp]),
    "sp_na_resample": (_i, [_i, vp, _i64, _i64, vp, C.c_int32, vp, _i64, _i, _i64, vp, vp, _u64, _i, vp, vp]),
    "sp_vd_work_scalars": (_i64, [_i, _i64]),
    "sp_vd_refresh": (_i, [C.POINTER(VdState), vp]),
    "sp_vd_sample": (_i, [C.POINTER(VdState), _i, _i, vp]),
    "sp_vd_update": (_i, [C.POINTER(VdState), _i, vp]),
    "sp_vd_generation": (_i, [C.POINTER(VdState), _i, vp]),
    "sp_vd_run": (_i, [C.POINTER(VdState), _i, _i, vp]),
}

_lib = None


class EngineError(RuntimeError):
    """A call into libstochopy_b200.so failed."""


def load():
    """Load the shared library once; raises ImportError if it was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C stochopy_b200/csrc` (there is no CPU fallback)"
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        msg = load().sp_last_error().decode("utf-8", "replace")
        raise EngineError(f"libstochopy_b200 error {rc}: {msg}")


# NVTX: one range per optimiser run (optimize/_helpers.py) always; with SP_NVTX=1 also one range per C-ABI
# call (a chunk of generations / a chain stage), named after the entry point -- off by default because the
# per-generation host paths would pay ~2 us per call for it.
_NVTX_CALLS = os.environ.get("SP_NVTX", "") not in ("", "0")


def nvtx_range(name):
    """Context manager: an NVTX range (visible in Nsight Systems / ncu --nvtx) around a block of host code."""
    import contextlib

    import torch

    @contextlib.contextmanager
    def _range():
        torch.cuda.nvtx.range_push(name)
        try:
            yield
        finally:
            torch.cuda.nvtx.range_pop()

    return _range()


def call(name, *args):
    if _NVTX_CALLS:
        with nvtx_range(name):
            check(getattr(load(), name)(*args))
        return
    check(getattr(load(), name)(*args))


def launch_count():
    return int(load().sp_launch_count())
