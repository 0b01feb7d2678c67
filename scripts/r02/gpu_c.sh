#!/bin/bash
# Round 2, call C: the side-plane layout of config C4 (L2D): parity, throughput against the extended record,
# ncu of the new kernel, full-size C4 (4096^2, 1e7 particles).
mkdir -p gpurun_out
T=r02c
python -m pytest tests -m gpu -q -x -k "side_plane or c4 or golden or step_parity or interp or restart" > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest_gpu.log
tail -3 gpurun_out/${T}_pytest_gpu.log
{
echo "== L2D (side plane), 1024^2 L2-resident"; python scripts/c4_probe.py c4 1024 400000 2
echo "== L2E (GPAT_NO_L2D=1), 1024^2";          GPAT_NO_L2D=1 python scripts/c4_probe.py c4 1024 400000 2
echo "== L2D, 4096^2";                          python scripts/c4_probe.py c4 4096 400000 2
echo "== L2E, 4096^2";                          GPAT_NO_L2D=1 python scripts/c4_probe.py c4 4096 400000 2
} > gpurun_out/${T}_c4_probe.log 2>&1
grep -E "^==|interval 2" gpurun_out/${T}_c4_probe.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:push_kernel -s 1 -c 1 -o gpurun_out/${T}_prof_c4_l2d python scripts/c4_probe.py c4 1024 400000 2 > gpurun_out/${T}_ncu_c4.log 2>&1
timeout 1500 python bench.py --workload c4 --nptl 10000000 --steps 1 --warmup 1 --no-cpu-baseline --no-membw > gpurun_out/${T}_full_c4.json 2> gpurun_out/${T}_full_c4.err
python -c "
import json;d=json.load(open('gpurun_out/${T}_full_c4.json'));print('c4 FULL value %.4g e2e %.4g frac %.3f push_ms %.1f nptl %d layout %s' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['breakdown_ms_per_step']['push_ms'], d['config']['particles_per_gpu'], d['config']['field_layout']))"; tail -2 gpurun_out/${T}_full_c4.err
