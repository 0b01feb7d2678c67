#!/bin/bash
# Round 2, call F: neighbour-preserving lane assignment (refill PERM4) A/B on C5; sanity subset of the GPU tests.
mkdir -p gpurun_out
T=r02f
python -m pytest tests -m gpu -q -x -k "side_plane or golden or step_parity or interval_parity" > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest_gpu.log
tail -3 gpurun_out/${T}_pytest_gpu.log
run() { n=$1; e=$2; shift 2
  env $e timeout 900 python bench.py "$@" --no-cpu-baseline --no-membw > gpurun_out/${T}_$n.json 2> gpurun_out/${T}_$n.err
  python -c "
import json;d=json.load(open('gpurun_out/${T}_$n.json'));print('$n value %.4g e2e %.4g push_ms %.1f clocks %s' % (d['value'], d['e2e']['value'], d['breakdown_ms_per_step']['push_ms'], d['clocks'].get('sm_mhz')))" || tail -3 gpurun_out/${T}_$n.err
}
NP=GPAT_LIB=$PWD/stochastic_parker_b200/csrc/libgpat_cuda.noperm.so
run c5_256_16m_perm "X=1" --workload c5 --grid 256 --nptl 16000000 --steps 2 --warmup 1
run c5_256_16m_noperm "$NP" --workload c5 --grid 256 --nptl 16000000 --steps 2 --warmup 1
run c5_256_2m_perm "X=1" --workload c5 --grid 256 --nptl 2000000 --steps 3 --warmup 1
run c5_256_2m_noperm "$NP" --workload c5 --grid 256 --nptl 2000000 --steps 3 --warmup 1
run c1_perm "X=1" --steps 4 --warmup 2 --no-strong
run c1_noperm "$NP" --steps 4 --warmup 2 --no-strong
timeout 600 ncu --set full --clock-control none -k regex:push_kernel -s 2 -c 1 -o gpurun_out/${T}_prof_c5_perm python bench.py --workload c5 --grid 256 --nptl 16000000 --steps 2 --warmup 1 --no-cpu-baseline --no-membw > gpurun_out/${T}_ncu_c5.log 2>&1
run c5_full_perm "X=1" --workload c5 --nptl 125000000 --steps 2 --warmup 1
