"""ORACLE (test infrastructure, not product code) -- ctypes front end of oracle/gca_oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  ``build()`` compiles the C restatement with gcc into
oracle/_build/libgca_oracle_<cpu-tag>.so (git-ignored, travels to the GPU box with the snapshot;
rebuilt there on first use if the box's CPU differs).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import prng
from .alexandridis import EnvConstants, VEG_PROBS, DEN_PROBS, F32

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "gca_oracle.c")
MAXR = 10


def _cpu_tag() -> str:
    """-march=native output is only valid on the CPU family that built it: tag the file with a
    hash of the CPU flags so the GPU box rebuilds (gcc is in the image) if its CPU differs."""
    import hashlib
    flags = ""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    flags = line
                    break
    except OSError:
        pass
    return hashlib.sha1(flags.encode()).hexdigest()[:8]


_OUT = os.path.join(_HERE, "_build", f"libgca_oracle_{_cpu_tag()}.so")


def build(force: bool = False) -> str:
    if not force and os.path.exists(_OUT) and os.path.getmtime(_OUT) >= os.path.getmtime(_SRC):
        return _OUT
    os.makedirs(os.path.dirname(_OUT), exist_ok=True)
    cmd = ["gcc", "-O3", "-march=native", "-fno-fast-math", "-ffp-contract=off", "-fopenmp", "-shared",
           "-fPIC", "-o", _OUT, _SRC, "-lm"]
    subprocess.check_call(cmd)
    return _OUT


class _Params(C.Structure):
    _fields_ = [
        ("H", C.c_int32), ("W", C.c_int32), ("R", C.c_int32), ("K", C.c_int32), ("rng_mode", C.c_int32),
        ("age_lo", C.c_int32), ("age_span", C.c_int32), ("age_mult", C.c_int32),
        ("day_length", C.c_int32),
        ("p_tree", C.c_float), ("p_wind_change", C.c_float), ("t_any", C.c_float),
        ("t_move", C.c_float * 9), ("t_shoot", C.c_float * 2),
        ("onep_veg", C.c_float * 6), ("onep_den", C.c_float * 6),
        ("winds", C.c_float * 72),
        ("dousing_weights", C.c_float * 25),
        ("burn_kernel", C.c_float * ((2 * MAXR + 1) ** 2)),
    ]


class _Inject(C.Structure):
    _fields_ = [("u_burn", C.c_void_p), ("u_grow", C.c_void_p), ("age_new", C.c_void_p),
                ("u_wind", C.c_void_p), ("wind_step", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_env_step.restype = C.c_int
        _lib.oracle_max_threads.restype = C.c_int
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class COracle:
    """Steps reference-layout state dicts (as produced by oracle.init_state) in place."""

    def __init__(self, E: EnvConstants, winds, K: int = 1, mode: int = prng.LEGACY):
        self.E = E
        P = _Params()
        P.H, P.W, P.R, P.K, P.rng_mode = E.nrows, E.ncols, E.ca.radius, K, mode
        lo, span, mult = prng.randint_params(E.ca.fire_age_min, E.ca.fire_age_max)
        P.age_lo, P.age_span, P.age_mult = lo, span, mult
        P.day_length = E.day_length
        P.p_tree, P.p_wind_change, P.t_any = float(E.p_tree), float(E.p_wind_change), float(E.t_any_f32)
        P.t_move[:] = [float(x) for x in E.movement_timings]
        P.t_shoot[:] = [float(x) for x in E.shooting_timings]
        P.onep_veg[:] = [float(x) for x in (F32(1) + VEG_PROBS).astype(F32)]
        P.onep_den[:] = [float(x) for x in (F32(1) + DEN_PROBS).astype(F32)]
        w = np.asarray(winds, dtype=F32)[:, 0].reshape(-1)
        P.winds[:] = [float(x) for x in w]
        P.dousing_weights[:] = [float(x) for x in E.ca.dousing_weights.reshape(-1)]
        bk = E.ca.burn_kernel.reshape(-1)
        for i, x in enumerate(bk):
            P.burn_kernel[i] = float(x)
        self.P = P
        self.K = K

    def step(self, state, action, inject=None, want_p=False, nthreads=0):
        """state: dict(per_env_context, position, time) of C-contiguous NumPy arrays with the
        reference dtypes; modified IN PLACE.  Returns (reward, terminated, counts[, p])."""
        ctx = state["per_env_context"]
        N, H, W = ctx["true_grid"].shape
        for k in ("true_grid", "fire_age", "dousing_count", "vegetation", "density", "pslope",
                  "wind_index", "key", "is_night", "time_step"):
            assert ctx[k].flags["C_CONTIGUOUS"], k
        assert ctx["true_grid"].dtype == np.float32 and ctx["fire_age"].dtype == np.float32
        assert ctx["dousing_count"].dtype == np.int32 and ctx["vegetation"].dtype == np.int32
        assert ctx["key"].dtype == np.uint32 and ctx["wind_index"].dtype == np.int32
        action = np.ascontiguousarray(action, dtype=np.int32)
        reward = np.zeros(N, dtype=np.float32)
        term = np.zeros(N, dtype=np.uint8)
        counts = np.zeros((N, 2), dtype=np.int32)
        p = np.zeros((N, H, W, 9), dtype=np.float32) if want_p else None
        inj = None
        keep = []
        if inject is not None:
            inj = _Inject()
            for name, dt in (("u_burn", np.float32), ("u_grow", np.float32), ("age_new", np.int32),
                             ("u_wind", np.float32), ("wind_step", np.int32)):
                a = inject.get(name)
                if a is not None:
                    a = np.ascontiguousarray(a, dtype=dt)
                    keep.append(a)
                    setattr(inj, name, a.ctypes.data)
        rc = lib().oracle_env_step(
            C.byref(self.P), C.c_int32(N), _ptr(ctx["true_grid"]), _ptr(ctx["fire_age"]),
            _ptr(ctx["dousing_count"]), _ptr(ctx["vegetation"]), _ptr(ctx["density"]), _ptr(ctx["pslope"]),
            _ptr(ctx["wind_index"]), _ptr(ctx["key"]), _ptr(ctx["is_night"]), _ptr(ctx["time_step"]),
            _ptr(state["position"]), _ptr(state["time"]), _ptr(action), _ptr(reward), _ptr(term),
            _ptr(counts), C.byref(inj) if inj is not None else None, _ptr(p), C.c_int32(nthreads))
        if rc != 0:
            raise RuntimeError(f"oracle_env_step failed: {rc}")
        if want_p:
            return reward, term.astype(bool), counts, p
        return reward, term.astype(bool), counts

    @staticmethod
    def max_threads() -> int:
        return int(lib().oracle_max_threads())
