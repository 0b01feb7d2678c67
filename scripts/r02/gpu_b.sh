#!/bin/bash
# Round 2, call B: full GPU test suite, sanitizer over smoke(), default bench (with the live L2 / HBM read
# bandwidths), ncu captures of the C4 and C5 steady state, full-size C2 and C5 (single-GPU share).
mkdir -p gpurun_out
T=r02b
python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest_gpu.log
tail -3 gpurun_out/${T}_pytest_gpu.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/${T}_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/${T}_sanitizer_racecheck.log
tail -2 gpurun_out/${T}_sanitizer_memcheck.log gpurun_out/${T}_sanitizer_racecheck.log
timeout 600 python bench.py > gpurun_out/${T}_bench_c1.json 2> gpurun_out/${T}_bench_c1.err
python -c "
import json;d=json.load(open('gpurun_out/${T}_bench_c1.json'));print('c1 value %.4g e2e %.4g frac %.3f' % (d['value'], d['e2e']['value'], d['roofline']['frac']), d['roofline'].get('l2'), d.get('strong_scaling'))"
# ncu: steady-state C4 (L2-resident grid) and sorted C5
timeout 600 ncu --set full --clock-control none --import-source on -k regex:push_kernel -s 1 -c 1 -o gpurun_out/${T}_prof_c4 python scripts/c4_probe.py c4 1024 400000 2 > gpurun_out/${T}_ncu_c4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:push_kernel -s 2 -c 1 -o gpurun_out/${T}_prof_c5 python bench.py --workload c5 --grid 256 --nptl 2000000 --steps 2 --warmup 1 --no-cpu-baseline --no-membw > gpurun_out/${T}_ncu_c5.log 2>&1
# full size: C5 single-GPU share (512^3, 1.25e8 particles), C2 (1e8 particles)
timeout 900 python bench.py --workload c5 --nptl 125000000 --steps 2 --warmup 1 --no-cpu-baseline --no-membw > gpurun_out/${T}_full_c5.json 2> gpurun_out/${T}_full_c5.err
timeout 900 python bench.py --workload c2 --nptl 100000000 --steps 1 --warmup 1 --no-cpu-baseline --no-membw > gpurun_out/${T}_full_c2.json 2> gpurun_out/${T}_full_c2.err
for wl in c5 c2; do python -c "
import json;d=json.load(open('gpurun_out/${T}_full_$wl.json'));print('$wl FULL value %.4g e2e %.4g frac %.3f push_ms %.1f nptl %d' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['breakdown_ms_per_step']['push_ms'], d['config']['particles_per_gpu']))"; tail -2 gpurun_out/${T}_full_$wl.err; done
nvidia-smi --query-gpu=memory.total,memory.used --format=csv
free -g | head -2
