#!/bin/bash
# Round 2, call L: validation of the last edits (3-D weight-scale folding, bench traffic key): GPU suite, C5 probe, C1 bench.
mkdir -p gpurun_out
T=r02l
python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest_gpu.log
tail -3 gpurun_out/${T}_pytest_gpu.log
python scripts/r02/c5_probe.py 256 16000000 "default:" > gpurun_out/${T}_c5_probe_256.log 2>&1; tail -1 gpurun_out/${T}_c5_probe_256.log
python scripts/r02/c5_probe.py 512 125000000 "default:" > gpurun_out/${T}_c5_probe_512.log 2>&1; tail -1 gpurun_out/${T}_c5_probe_512.log
timeout 600 python bench.py > gpurun_out/${T}_bench_c1.json 2> gpurun_out/${T}_bench_c1.err
python -c "
import json;d=json.load(open('gpurun_out/${T}_bench_c1.json'));print('c1 value %.4g e2e %.4g frac %.3f traffic %s' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic']))" || tail -3 gpurun_out/${T}_bench_c1.err
timeout 600 python bench.py --workload c3 --steps 3 --warmup 2 --no-cpu-baseline --no-membw > gpurun_out/${T}_bench_c3.json 2> gpurun_out/${T}_bench_c3.err
python -c "
import json;d=json.load(open('gpurun_out/${T}_bench_c3.json'));print('c3 value %.4g e2e %.4g frac %.3f traffic %s' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic']))" || tail -3 gpurun_out/${T}_bench_c3.err
