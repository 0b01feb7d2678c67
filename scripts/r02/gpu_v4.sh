#!/bin/bash
# Round 2, call V4: final state of the round -- full GPU suite, smoke, default bench (both arms), C4 and C3 lines.
mkdir -p gpurun_out
T=r02v4
python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest_gpu.log
tail -3 gpurun_out/${T}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log
timeout 600 python bench.py --impl reference > gpurun_out/${T}_ref_c1.json 2> gpurun_out/${T}_ref_c1.err; cut -c1-160 gpurun_out/${T}_ref_c1.json
timeout 600 python bench.py > gpurun_out/${T}_bench_c1.json 2> gpurun_out/${T}_bench_c1.err
python -c "
import json;d=json.load(open('gpurun_out/${T}_bench_c1.json'));print('c1 value %.4g e2e %.4g frac %.3f cpu %.4g/%d cores launches %d' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['gpu_launches']), d['roofline']['l2'])" || tail -3 gpurun_out/${T}_bench_c1.err
timeout 900 python bench.py --workload c4 --grid 1024 --nptl 2000000 --steps 1 --warmup 1 --no-cpu-baseline --no-membw > gpurun_out/${T}_c4_1024.json 2>/dev/null
timeout 900 python bench.py --workload c3 --steps 4 --warmup 3 --no-cpu-baseline --no-membw > gpurun_out/${T}_c3.json 2>/dev/null
for n in c4_1024 c3; do python -c "
import json;d=json.load(open('gpurun_out/${T}_$n.json'));print('$n value %.4g e2e %.4g' % (d['value'], d['e2e']['value']))"; done
