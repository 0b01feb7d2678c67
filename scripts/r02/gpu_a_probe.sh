#!/bin/bash
# Round 2, call A: Fortran-compiler probe on the B200 box + state of the GPU tests + staged random-switch test.
mkdir -p gpurun_out
{
  echo "== fortran probe on $(hostname) $(date -u +%FT%TZ)"
  for c in gfortran flang flang-new nvfortran pgfortran ifort ifx mpif90 mpifort h5fc lfortran f2c f77 f95 g77; do
    p=$(command -v $c 2>/dev/null); echo "$c: ${p:-absent}"
  done
  echo "gcc f951: $(gcc -print-prog-name=f951)"
  ls /usr/lib/gcc/x86_64-linux-gnu/*/f951 2>&1
  find / -xdev \( -name 'f951*' -o -name 'flang*' -o -name 'nvfortran*' -o -name 'gfortran*' \) 2>/dev/null | grep -v '^/proc' | head
  ls /opt/nvidia/hpc_sdk 2>&1 | head -3
  echo "== host"; nproc; lscpu | grep -E 'Model name|Socket|Core|Thread' ; nvidia-smi -L
} > gpurun_out/r02a_fortran_probe.log 2>&1
python -m pytest tests -m gpu -q -x > gpurun_out/r02a_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest_gpu.log
PYTHONPATH=tests python -m pytest scripts/next_round/test_gpu_random_switches.py -q > gpurun_out/r02a_random_switches.log 2>&1; echo "rc=$?" >> gpurun_out/r02a_random_switches.log
tail -5 gpurun_out/r02a_pytest_gpu.log gpurun_out/r02a_random_switches.log; cat gpurun_out/r02a_fortran_probe.log
