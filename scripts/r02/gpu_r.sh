#!/bin/bash
# Round 2, call R: published factors (PUBF) vs published weights A/B on the 2-D kernels.
mkdir -p gpurun_out
T=r02r
python -m pytest tests -m gpu -q -x -k "side_plane or golden or step_parity or interval_parity or restart or sorted or table" > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest_gpu.log
tail -3 gpurun_out/${T}_pytest_gpu.log
run() { n=$1; e=$2; shift 2
  env $e timeout 900 python bench.py "$@" --no-cpu-baseline --no-membw --no-strong > gpurun_out/${T}_$n.json 2> gpurun_out/${T}_$n.err
  python -c "
import json;d=json.load(open('gpurun_out/${T}_$n.json'));print('$n value %.4g e2e %.4g push_ms %.2f clocks %s' % (d['value'], d['e2e']['value'], d['breakdown_ms_per_step']['push_ms'], d['clocks'].get('sm_mhz')))" || tail -3 gpurun_out/${T}_$n.err
}
OLD=GPAT_LIB=$PWD/stochastic_parker_b200/csrc/libgpat_cuda.pubw.so
run c1_pubf "X=1" --steps 6 --warmup 3
run c1_pubw "$OLD" --steps 6 --warmup 3
run c1_pubf_b "X=1" --steps 6 --warmup 3
run c1_pubw_b "$OLD" --steps 6 --warmup 3
run c3_pubf "X=1" --workload c3 --steps 4 --warmup 2
run c3_pubw "$OLD" --workload c3 --steps 4 --warmup 2
run c4_pubf "X=1" --workload c4 --grid 1024 --nptl 2000000 --steps 1 --warmup 1
run c4_pubw "$OLD" --workload c4 --grid 1024 --nptl 2000000 --steps 1 --warmup 1
run c2_pubf "X=1" --workload c2 --nptl 4000000 --steps 1 --warmup 1
run c2_pubw "$OLD" --workload c2 --nptl 4000000 --steps 1 --warmup 1
