"""stochastic_parker_b200 -- B200-native pseudo-particle SDE push for GPAT.

The product is csrc/libgpat_cuda.so (hand-written sm_100a CUDA behind the C ABI of
include/gpat_cuda.h).  The Python modules are the host-side mirror of the reference's
driver for this one path: `abi` (ctypes binding), `driver` (call sequence of
stochastic-mhd.f90), `config` (conf.dat grammar + named workloads), `mhd` (synthetic
frames in the reference's on-disk format).
"""
from . import outputs  # noqa: F401
from .abi import LIB_PATH, PARTICLE_DTYPE, Counters, Params, Timings, load_library  # noqa: F401
from .config import WORKLOADS, Workload, build_params  # noqa: F401
from .driver import GpatError, GpatSim, dump_restart, read_restart, run_intervals  # noqa: F401
from .multi import bootstrap_comm, rank_info, reduce_diagnostics, shard_count  # noqa: F401
