"""ctypes view of include/gpat_cuda.h and the loader of libgpat_cuda.so.

The shared library is the product; this module is only a thin binding so that the
Python host mirror (driver.py), the tests and bench.py can call the same C ABI the
Fortran driver would call through ISO_C_BINDING.  There is NO fallback: if the
library is missing, or no CUDA device is present at gpat_init time, the call
fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# GPAT_LIB selects an experiment build of the same library (csrc/Makefile `variant`)
LIB_PATH = os.environ.get("GPAT_LIB") or os.path.join(_HERE, "csrc", "libgpat_cuda.so")

GPAT_OK = 0
COUNT_FLAG_INBOX = 1
RNG_PHILOX = 0
RNG_TABLE = 1


class HistSpec(C.Structure):
    _fields_ = [
        ("enabled", C.c_int32), ("npbins", C.c_int32), ("nmu", C.c_int32),
        ("rx", C.c_int32), ("ry", C.c_int32), ("rz", C.c_int32),
        ("pmin", C.c_double), ("pmax", C.c_double),
    ]


# gpat_inject_targeted modes (include/gpat_cuda.h)
INJECT_LARGE_JZ, INJECT_LARGE_ABSJ, INJECT_LARGE_DB2, INJECT_LARGE_DIVV, INJECT_LARGE_RHO = 1, 2, 3, 4, 5


class Params(C.Structure):
    """struct gpat_params (include/gpat_cuda.h)."""
    _fields_ = [
        ("ndim", C.c_int32), ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
        ("time_interp", C.c_int32), ("pbc", C.c_int32 * 3),
        ("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double),
        ("xmin", C.c_double), ("ymin", C.c_double), ("zmin", C.c_double),
        ("xmax", C.c_double), ("ymax", C.c_double), ("zmax", C.c_double),
        ("lx", C.c_double), ("ly", C.c_double), ("lz", C.c_double),
        ("b0", C.c_double), ("p0", C.c_double), ("pmin", C.c_double), ("pmax", C.c_double),
        ("gamma_turb", C.c_double), ("pindex", C.c_double),
        ("kpara0", C.c_double), ("kret", C.c_double),
        ("dt_min_rel", C.c_double), ("dt_max_rel", C.c_double),
        ("momentum_dependency", C.c_int32), ("mag_dependency", C.c_int32),
        ("acc_region_flag", C.c_int32), ("pad0_", C.c_int32),
        ("acc_region", C.c_double * 6),
        ("dpp_wave", C.c_int32), ("dpp_shear", C.c_int32), ("weak_scattering", C.c_int32),
        ("keep_rho", C.c_int32),
        ("tau0", C.c_double),
        ("drift1", C.c_double), ("drift2", C.c_double),
        ("pcharge", C.c_int32), ("check_drift_2d", C.c_int32),
        ("include_3rd_dim", C.c_int32), ("nlgc", C.c_int32),
        ("kperp_kpara", C.c_double), ("duu0", C.c_double),
        ("focused_transport", C.c_int32), ("spherical_coord", C.c_int32),
        ("nonuniform_grid", C.c_int32),
        ("deltab_flag", C.c_int32), ("correlation_flag", C.c_int32), ("acc_by_surface", C.c_int32),
        ("npp_global", C.c_int32), ("nmu_global", C.c_int32),
        ("local", HistSpec * 4),
        ("seed", C.c_uint64),
        ("rng_mode", C.c_int32),
        ("mpi_rank", C.c_int32),
        ("strict_math", C.c_int32),
        ("surface_norm1", C.c_int32), ("surface_norm2", C.c_int32),
        ("surface2_existed", C.c_int32), ("is_intersection", C.c_int32), ("pad2_", C.c_int32),
    ]

    def copy(self) -> "Params":
        q = Params()
        C.memmove(C.byref(q), C.byref(self), C.sizeof(Params))
        return q


class Counters(C.Structure):
    _fields_ = [
        ("nptl_current", C.c_int64), ("nptl_split", C.c_int64), ("nptl_escaped", C.c_int64),
        ("nptl_max", C.c_int64), ("tag_max", C.c_int64),
        ("leak", C.c_double), ("leak_negp", C.c_double),
    ]


class Timings(C.Structure):
    _fields_ = [
        ("mover_ms", C.c_float), ("push_ms", C.c_float), ("compact_ms", C.c_float),
        ("upload_ms", C.c_float), ("grad_ms", C.c_float),
        ("inject_ms", C.c_float), ("split_ms", C.c_float), ("diag_ms", C.c_float),
        ("push_steps", C.c_uint64),
        ("push_launches", C.c_uint32), ("total_launches", C.c_uint32),
    ]


# AoS particle record == particle_type (particle_module.f90:38-50), 104 bytes
PARTICLE_DTYPE = np.dtype(
    {
        "names": ["split_times", "count_flag", "origin", "nsteps_tracked", "nsteps_pushed",
                  "tag_injected", "tag_splitted", "x", "y", "z", "p", "v", "mu", "weight", "t",
                  "dt", "padding"],
        "formats": ["i1", "i1", "i4", "i4", "i4", "i4", "i4", "f8", "f8", "f8", "f8", "f8", "f8",
                    "f8", "f8", "f8", "f8"],
        "offsets": [0, 1, 4, 8, 12, 16, 20, 24, 32, 40, 48, 56, 64, 72, 80, 88, 96],
        "itemsize": 104,
    }
)


def rng_steps(ptl: np.ndarray) -> np.ndarray:
    """The per-particle 64-bit RNG step counter carried in `padding`."""
    return np.ascontiguousarray(ptl["padding"]).view(np.uint64)


_DP = C.POINTER(C.c_double)
_FP = C.POINTER(C.c_float)

# name -> (restype, argtypes); must list every symbol include/gpat_cuda.h declares
SIGNATURES = {
    "gpat_init": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int64, C.POINTER(Params)]),
    "gpat_set_params": (C.c_int, [C.c_void_p, C.POINTER(Params)]),
    "gpat_finalize": (C.c_int, [C.c_void_p]),
    "gpat_last_error": (C.c_char_p, [C.c_void_p]),
    "gpat_upload_fields": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]),
    "gpat_swap_fields": (C.c_int, [C.c_void_p]),
    "gpat_upload_turbulence": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "gpat_upload_acc_surface": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "gpat_prefetch_fields": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "gpat_inject_uniform": (C.c_int, [C.c_void_p, C.c_int64, C.c_double, C.c_int, C.c_double,
                                      C.c_double, C.c_double, _DP, C.c_double]),
    "gpat_inject_targeted": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_int, C.c_double,
                                       C.c_double, C.c_double, _DP, C.c_double, C.c_int, C.c_double,
                                       C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "gpat_inject_at_shock": (C.c_int, [C.c_void_p, C.c_int64, C.c_double, C.c_int, C.c_double, C.c_double,
                                       C.c_double]),
    "gpat_init_tracking": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int]),
    "gpat_tracked_shape": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "gpat_download_tracked": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gpat_reset_tracked": (C.c_int, [C.c_void_p]),
    "gpat_particle_mover": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int,
                                      C.c_int, C.POINTER(C.c_uint64)]),
    "gpat_split": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_int]),
    "gpat_download_particles": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64,
                                          C.POINTER(C.c_int64)]),
    "gpat_upload_particles": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "gpat_download_escaped": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "gpat_reset_escaped": (C.c_int, [C.c_void_p]),
    "gpat_get_counters": (C.c_int, [C.c_void_p, C.POINTER(Counters)]),
    "gpat_set_counters": (C.c_int, [C.c_void_p, C.POINTER(Counters)]),
    "gpat_diagnostics": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_void_p),
                                   C.c_void_p, C.c_void_p]),
    "gpat_escaped_diagnostics": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gpat_escaped_local_diagnostics": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gpat_hist_edges": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "gpat_comm_unique_id": (C.c_int, [C.c_char_p]),
    "gpat_comm_init": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.c_int]),
    "gpat_comm_destroy": (C.c_int, [C.c_void_p]),
    "gpat_get_timings": (C.c_int, [C.c_void_p, C.POINTER(Timings)]),
    "gpat_set_rng_table": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]),
    "gpat_debug_gradients": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "gpat_debug_push_n": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_int,
                                    C.POINTER(C.c_uint64)]),
    "gpat_debug_interp": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
}

_lib = None


def load_library(path: str | None = None) -> C.CDLL:
    """dlopen libgpat_cuda.so and bind every entry point; raises if it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or make -C stochastic_parker_b200/csrc). There is no CPU fallback.")
    lib = C.CDLL(p, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def ptr(a: np.ndarray | None):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"] or a.flags["F_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)
