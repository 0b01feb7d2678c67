#!/bin/bash
# Round 2, call D: full GPU suite; L3D (3-D record split at the 128-byte line) against L3B at 256^3 and at full size;
# ncu of the L3D kernel; C1 with the warp-aggregated local histograms + fresh ncu capture and launch list.
mkdir -p gpurun_out
T=r02d
python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest_gpu.log
tail -4 gpurun_out/${T}_pytest_gpu.log
run() { # name, env, args...
  n=$1; e=$2; shift 2
  env $e timeout 900 python bench.py "$@" --no-cpu-baseline --no-membw > gpurun_out/${T}_$n.json 2> gpurun_out/${T}_$n.err
  python -c "
import json;d=json.load(open('gpurun_out/${T}_$n.json'));print('$n value %.4g e2e %.4g frac %.3f push_ms %.1f diag_ms %.2f layout %s clocks %s' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['breakdown_ms_per_step']['push_ms'], d['breakdown_ms_per_step']['diag_ms'], d['config']['field_layout'], d['clocks'].get('sm_mhz')))" || tail -3 gpurun_out/${T}_$n.err
}
run c5_256_l3d "X=1" --workload c5 --grid 256 --nptl 2000000 --steps 3 --warmup 1
run c5_256_l3b "GPAT_NO_L3D=1" --workload c5 --grid 256 --nptl 2000000 --steps 3 --warmup 1
run c5_256_l3d_8m "X=1" --workload c5 --grid 256 --nptl 16000000 --steps 2 --warmup 1
run c5_256_l3b_8m "GPAT_NO_L3D=1" --workload c5 --grid 256 --nptl 16000000 --steps 2 --warmup 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:push_kernel -s 2 -c 1 -o gpurun_out/${T}_prof_c5_l3d python bench.py --workload c5 --grid 256 --nptl 2000000 --steps 2 --warmup 1 --no-cpu-baseline --no-membw > gpurun_out/${T}_ncu_c5.log 2>&1
run c5_full_l3d "X=1" --workload c5 --nptl 125000000 --steps 2 --warmup 1
run c4_1024_2m "X=1" --workload c4 --grid 1024 --nptl 2000000 --steps 1 --warmup 1
timeout 600 python bench.py > gpurun_out/${T}_bench_c1.json 2> gpurun_out/${T}_bench_c1.err
python -c "
import json;d=json.load(open('gpurun_out/${T}_bench_c1.json'));print('c1 value %.4g e2e %.4g frac %.3f diag_ms %.2f' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['breakdown_ms_per_step']['diag_ms']), d['roofline'].get('l2'))"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:push_kernel -s 1 -c 1 -o gpurun_out/${T}_prof_c1 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-membw --no-strong > gpurun_out/${T}_ncu_c1.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_c1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-membw --no-strong > gpurun_out/${T}_ncu_l_c1.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:diag_kernel -s 1 -c 1 -o gpurun_out/${T}_prof_diag python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-membw --no-strong > gpurun_out/${T}_ncu_diag.log 2>&1
