"""Generates tests/golden/ref_f90/<case>.npz by RUNNING THE REFERENCE'S OWN FORTRAN.

    python tests/golden/make_ref_f90_golden.py [case ...]        (in the build container; needs /root/reference)

No Fortran compiler exists in this image or on the B200 box (profiles/r02a_fortran_probe.log), so the unmodified
procedures of /root/reference/src/modules/*.f90 are executed by oracle/f90/f90run.py (a Fortran-90-subset
translator with Fortran's kind, promotion, literal and evaluation-order rules; see its header) through
oracle/f90/refsim.py.  The files this script writes are what pins oracle/gpat_oracle.c -- and, on the B200, the
CUDA library -- to the reference: tests/test_cpu_reference_f90.py holds the C oracle to them BIT FOR BIT, and
tests/test_gpu_parity.py::test_gpu_against_reference_golden holds the GPU build to them at north_star's 1e-12.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "f90"))

from helpers import REF_GOLDEN_CASES, golden_collect  # noqa: E402
from refsim import RefSim  # noqa: E402


def main(names):
    os.makedirs(os.path.join(HERE, "ref_f90"), exist_ok=True)
    for name in names:
        t = time.time()
        out = golden_collect(lambda P, n: RefSim(P, n), name)
        np.savez_compressed(os.path.join(HERE, "ref_f90", name + ".npz"), **out)
        print(f"{name}: {int(out['run_steps'])} interval steps, {time.time() - t:.0f} s", flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or REF_GOLDEN_CASES)
