#!/bin/bash
# Round 2, call K: final build -- full GPU suite, smoke, default bench, C5 at full size on one GPU, and DRAM bytes per step
# of the push kernel on the full-size GRIDS of C2 / C4 / C5 / C3 (ncu, DRAM + L2 metrics only: one pass, no replay of the
# minutes-long kernels; populations reduced where a launch would take minutes).
mkdir -p gpurun_out
T=r02k
python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest_gpu.log
tail -3 gpurun_out/${T}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log
timeout 600 python bench.py > gpurun_out/${T}_bench_c1.json 2> gpurun_out/${T}_bench_c1.err
python -c "
import json;d=json.load(open('gpurun_out/${T}_bench_c1.json'));print('c1 value %.4g e2e %.4g frac %.3f cpu %.4g/%d cores' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['cpu_baseline']['cores']), 'extra' in d, d['roofline'].get('pipes',{}).get('l1_data_pipe_pct'))" || tail -3 gpurun_out/${T}_bench_c1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_ref_c1.json 2> gpurun_out/${T}_ref_c1.err; cut -c1-200 gpurun_out/${T}_ref_c1.json
timeout 900 python bench.py --workload c5 --nptl 125000000 --steps 2 --warmup 1 --no-cpu-baseline --no-membw > gpurun_out/${T}_full_c5.json 2> gpurun_out/${T}_full_c5.err
python -c "
import json;d=json.load(open('gpurun_out/${T}_full_c5.json'));print('c5 FULL value %.4g e2e %.4g frac %.3f clocks %s' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks']))" || tail -3 gpurun_out/${T}_full_c5.err
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_bytes.sum
for spec in "c2 4000000" "c4 1000000" "c5 125000000" "c3 1000000"; do
  set -- $spec
  timeout 900 ncu --metrics $M --clock-control none -k regex:push_kernel -s 1 -c 1 --csv --log-file gpurun_out/${T}_dram_$1.csv python bench.py --workload $1 --nptl $2 --steps 1 --warmup 1 --no-cpu-baseline --no-membw > gpurun_out/${T}_dram_$1.json 2> gpurun_out/${T}_dram_$1.err
  grep -E "push_kernel" gpurun_out/${T}_dram_$1.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr -d '"' | paste -sd' ' | cut -c1-400
  python -c "
import json;d=json.load(open('gpurun_out/${T}_dram_$1.json'));print('$1 steps per launch under ncu: %d' % round(d['value']*d['ms_per_step']*1e-3))"
done
