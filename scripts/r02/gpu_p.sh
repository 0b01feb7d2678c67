#!/bin/bash
# Round 2, call P: ncu of the rolled 3-D kernel (256^3, 1.6e7 particles) + the full-size single-GPU C5 bench line.
mkdir -p gpurun_out
T=r02p
timeout 600 ncu --set full --clock-control none --import-source on -k regex:push_kernel -s 1 -c 1 -o gpurun_out/${T}_prof_c5_rolled python scripts/r02/c5_probe.py 256 16000000 "rolled:" > gpurun_out/${T}_ncu_c5.log 2>&1
tail -2 gpurun_out/${T}_ncu_c5.log
timeout 900 python bench.py --workload c5 --nptl 125000000 --steps 2 --warmup 1 --no-cpu-baseline --no-membw > gpurun_out/${T}_full_c5.json 2> gpurun_out/${T}_full_c5.err
python -c "
import json;d=json.load(open('gpurun_out/${T}_full_c5.json'));print('c5 FULL value %.4g e2e %.4g frac %.3f clocks %s' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks']))" || tail -3 gpurun_out/${T}_full_c5.err
