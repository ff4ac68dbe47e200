"""ctypes binding of libgca.so (the C ABI declared in include/gca.h).

PyTorch tensors are used only as zero-copy device buffers: every call passes raw
``data_ptr()`` addresses, sizes and the current CUDA stream handle.  There is NO CPU fallback:
if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
# GCA_LIB_PATH selects another build of the same sources (A/B measurements of kernel variants)
LIB_PATH = os.environ.get("GCA_LIB_PATH") or os.path.join(_HERE, "libgca.so")
SOURCES = ("gca_step64.cu", "gca_bb.cu", "gca_hidden.cu", "gca_tiled.cu", "gca_windy.cu", "gca_aux.cu", "gca_abi.cu")
HEADERS = ("gca_common.cuh", os.path.join("..", "..", "include", "gca.h"))

GCA_MAX_R = 10
GCA_MAX_K = 8
RNG_LEGACY, RNG_PARTITIONABLE = 0, 1
FLAG_AUTO_RESET, FLAG_NO_HIDDEN, FLAG_CA_ONLY, FLAG_NO_TMA, FLAG_WORK_CYCLES, FLAG_HOST_COPY = 1, 2, 4, 8, 16, 96
FLAG_HOST_COPY_IN, FLAG_HOST_COPY_OUT = 32, 64
FLAG_HOST_ASYNC = 128
FLAG_HOST_MAPPED = 256
FLAG_RENDER = 512
FLAG_GENERIC_TILES = 1024
GCA_VERSION = 105  # include/gca.h: the ctypes structures below mirror that version of the header


class GcaError(RuntimeError):
    pass


class GcaParams(C.Structure):
    _fields_ = [
        ("H", C.c_int32), ("W", C.c_int32), ("R", C.c_int32), ("K", C.c_int32), ("rng_mode", C.c_int32),
        ("age_lo", C.c_int32), ("age_span", C.c_uint32), ("age_mult", C.c_uint32),
        ("day_length", C.c_int32),
        ("p_tree", C.c_float), ("p_wind_change", C.c_float), ("t_any", C.c_float),
        ("t_move", C.c_float * 9), ("t_shoot", C.c_float * 2),
        ("onep_veg", C.c_float * 8), ("onep_den", C.c_float * 8),
        ("winds", C.c_float * 72),
        ("dous_border", C.c_float), ("dous_inner", C.c_float),
        ("ring_w", C.c_float * (GCA_MAX_R + 1)),
    ]


class GcaState(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("reserved", C.c_int32),
        ("cell", C.c_void_p), ("death", C.c_void_p), ("hidden", C.c_void_p), ("doused", C.c_void_p),
        ("pslope", C.c_void_p), ("row_min", C.c_void_p), ("tick", C.c_void_p), ("key", C.c_void_p),
        ("wind_index", C.c_void_p), ("position", C.c_void_p), ("time", C.c_void_p),
        ("time_step", C.c_void_p), ("is_night", C.c_void_p),
        ("steps_elapsed", C.c_void_p), ("reward_accumulated", C.c_void_p),
        ("scratch_cell", C.c_void_p), ("scratch_u32", C.c_void_p),
        ("work", C.c_void_p), ("order", C.c_void_p), ("bb", C.c_void_p),
    ]


class GcaStepOut(C.Structure):
    _fields_ = [("reward", C.c_void_p), ("step_reward", C.c_void_p), ("terminated", C.c_void_p),
                ("counts", C.c_void_p), ("obs_night", C.c_void_p), ("stats", C.c_void_p),
                ("host_reward", C.c_void_p), ("host_terminated", C.c_void_p),
                ("host_done", C.c_void_p), ("done_counter", C.c_void_p), ("done_token", C.c_uint32),
                ("rgb_u8", C.c_uint32), ("rgb", C.c_void_p)]


_EPISODE_FIELDS = ("episode_returns", "episode_lengths", "returned_episode_returns", "returned_episode_lengths",
                   "amount_finished", "recent_returns", "recent_lengths", "recent_idx", "current_day_correct",
                   "current_night_correct", "current_day_steps", "current_night_steps", "recent_day_correct",
                   "recent_night_correct", "recent_day_steps", "recent_night_steps")


class GcaEpisodeStats(C.Structure):
    _fields_ = [(f, C.c_void_p) for f in _EPISODE_FIELDS]


class GcaInject(C.Structure):
    _fields_ = [("u_burn", C.c_void_p), ("u_grow", C.c_void_p), ("age_new", C.c_void_p),
                ("u_wind", C.c_void_p), ("wind_step", C.c_void_p)]


def source_hash() -> str:
    """16 hex digits over the CUDA sources and headers next to this file: the build id compiled into libgca.so."""
    import hashlib
    h = hashlib.sha256()
    for d in [os.path.join(_CSRC, s) for s in SOURCES] + [os.path.normpath(os.path.join(_CSRC, x)) for x in HEADERS]:
        if os.path.exists(d):
            with open(d, "rb") as f:
                h.update(f.read())
    return h.hexdigest()[:16]


def _built_id(path: str):
    """Build id of a libgca.so, read without binding anything else (None: not loadable / no such symbol)."""
    try:
        lib = C.CDLL(path)
        lib.gca_build_id.restype = C.c_char_p
        return lib.gca_build_id().decode()
    except (OSError, AttributeError):
        return None


def _needs_build() -> bool:
    """The library is missing or was built from other sources than the ones next to it (content hash, not mtimes: a
    copied tree keeps no useful timestamps)."""
    return not os.path.exists(LIB_PATH) or _built_id(LIB_PATH) != source_hash()


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a into gym_cellular_automata_b200/libgca.so (in-tree)."""
    if not force and not _needs_build():
        return LIB_PATH
    srcs = [os.path.join(_CSRC, s) for s in SOURCES if os.path.exists(os.path.join(_CSRC, s))]
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           f'-DGCA_BUILD_ID="{source_hash()}"', "-Xcompiler", "-fPIC", "-shared", "-o", LIB_PATH] + srcs
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise GcaError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB_PATH


_lib = None


def load():
    """Load libgca.so; raises GcaError when it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH) and os.environ.get("GCA_LIB_PATH"):
        raise GcaError(f"{LIB_PATH} is missing.  There is no CPU fallback.")
    if not os.environ.get("GCA_LIB_PATH") and _needs_build():
        # a binary built from other sources would be bound with the wrong structure layouts: rebuild it (nvcc is part
        # of the image: ~30 s), refuse it when that is not possible
        try:
            import fcntl
            with open(LIB_PATH + ".lock", "w") as lk:   # ranks of one node start together: one of them builds
                fcntl.flock(lk, fcntl.LOCK_EX)
                if _needs_build():
                    sys.stderr.write(f"[gca] {LIB_PATH} was not built from the sources next to it: rebuilding\n")
                    build_library(force=True)
        except Exception as exc:
            raise GcaError(f"{LIB_PATH} was not built from the sources next to it (csrc/*.cu, include/gca.h) and could not "
                           f"be rebuilt ({exc}); run `python -c 'import __graft_entry__ as g; g.build()'`") from exc
    lib = C.CDLL(LIB_PATH)
    lib.gca_version.restype = C.c_int
    if lib.gca_version() != GCA_VERSION and not os.environ.get("GCA_SKIP_VERSION_CHECK"):  # (A/B runs against older builds)
        raise GcaError(f"{LIB_PATH} reports gca_version {lib.gca_version()}, this binding was written for {GCA_VERSION}: "
                       "rebuild the library (structure layouts may differ)")
    lib.gca_last_error.restype = C.c_char_p
    for name in EXPORTS:
        if name not in ("gca_version", "gca_last_error", "gca_build_id") and hasattr(lib, name):
            getattr(lib, name).restype = C.c_int
    lib.gca_params_init.argtypes = [C.POINTER(GcaParams), C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double,
                                    C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int32,
                                    C.c_void_p]
    lib.gca_env_step.argtypes = [C.POINTER(GcaParams), C.POINTER(GcaState), C.c_void_p, C.POINTER(GcaStepOut),
                                 C.POINTER(GcaInject), C.POINTER(GcaState), C.c_void_p, C.c_uint32, C.c_void_p]
    lib.gca_env_step_host.argtypes = [C.POINTER(GcaParams), C.POINTER(GcaState), C.c_void_p, C.c_void_p,
                                      C.POINTER(GcaStepOut), C.POINTER(GcaState), C.c_void_p, C.c_uint32,
                                      C.c_void_p, C.c_void_p, C.c_void_p]
    if hasattr(lib, "gca_host_wait"):
        lib.gca_host_wait.argtypes = [C.c_void_p, C.c_uint32, C.c_double, C.c_void_p]
    lib.gca_alexandridis_step.argtypes = [C.POINTER(GcaParams), C.POINTER(GcaState), C.POINTER(GcaStepOut),
                                          C.POINTER(GcaInject), C.c_uint32, C.c_void_p]
    lib.gca_move_modify.argtypes = [C.POINTER(GcaParams), C.POINTER(GcaState), C.c_void_p, C.c_void_p]
    lib.gca_reward_done.argtypes = [C.POINTER(GcaParams), C.POINTER(GcaState), C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p]
    lib.gca_conditional_reset.argtypes = [C.POINTER(GcaParams), C.POINTER(GcaState), C.POINTER(GcaState),
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.gca_render_rgb.argtypes = [C.POINTER(GcaParams), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.gca_render_rgb_actions.argtypes = lib.gca_render_rgb.argtypes
    lib.gca_pack_state.argtypes = [C.POINTER(GcaParams), C.POINTER(GcaState)] + [C.c_void_p] * 8
    lib.gca_unpack_state.argtypes = [C.POINTER(GcaParams), C.POINTER(GcaState)] + [C.c_void_p] * 4
    lib.gca_balance_order.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.gca_generate_hidden.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_uint64, C.c_int32] + [C.c_void_p] * 7
    lib.gca_episode_stats_update.argtypes = [C.c_int32, C.POINTER(GcaEpisodeStats)] + [C.c_void_p] * 6
    lib.gca_windy_env_step.argtypes = [C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 7 + [C.c_int32, C.c_double,
                                      C.c_double, C.c_double] + [C.c_void_p] * 5
    lib.gca_windy_pack.argtypes = [C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 4
    lib.gca_windy_unpack.argtypes = [C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 4
    lib.gca_threefry_bits.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]
    lib.gca_threefry_split.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    _lib = lib
    return lib


# every symbol include/gca.h declares (tests check the library exports all of them)
EXPORTS = ("gca_version", "gca_last_error", "gca_build_id", "gca_params_init", "gca_env_step", "gca_env_step_host", "gca_host_wait", "gca_alexandridis_step", "gca_render_rgb_actions",
           "gca_move_modify", "gca_reward_done", "gca_conditional_reset", "gca_render_rgb", "gca_pack_state",
           "gca_unpack_state", "gca_balance_order", "gca_generate_hidden", "gca_episode_stats_update", "gca_windy_env_step", "gca_windy_pack", "gca_windy_unpack",
           "gca_threefry_bits", "gca_threefry_split")


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().gca_last_error().decode(errors="replace")
        raise GcaError(f"{what or 'libgca'} failed (status {rc}): {msg}")


def ptr(t, dtype=None, numel=None, name="tensor", allow_pinned=False):
    """Raw device address of a contiguous CUDA tensor (None -> NULL).  ``allow_pinned``: a pinned host tensor is
    accepted as well -- its pages are mapped into the device's address space, a kernel reads them over the bus."""
    if t is None:
        return None
    if not t.is_cuda and not (allow_pinned and t.is_pinned()):
        raise GcaError(f"{name}: expected a CUDA tensor (libgca has no CPU path)")
    if not t.is_contiguous():
        raise GcaError(f"{name}: tensor must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise GcaError(f"{name}: dtype {t.dtype}, expected {dtype}")
    if numel is not None and t.numel() != numel:
        raise GcaError(f"{name}: {t.numel()} elements, expected {numel}")
    return C.c_void_p(t.data_ptr())


def current_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
