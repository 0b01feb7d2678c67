mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench3_c1.json 2> gpurun_out/bench3_c1.err
python -c "
import json;d=json.load(open('gpurun_out/bench3_c1.json'));print('c1 value %.4g e2e %.4g' % (d['value'], d['e2e']['value']), d['breakdown_ms_per_step'])"
timeout 300 python bench.py --workload c5 --grid 256 --nptl 2000000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench3_c5.json 2> gpurun_out/bench3_c5.err
python -c "
import json;d=json.load(open('gpurun_out/bench3_c5.json'));print('c5 value %.4g' % d['value'])"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:push_kernel -s 1 -c 1 -o gpurun_out/prof_c4 python bench.py --workload c4 --grid 2048 --nptl 150000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c4.log 2>&1
tail -2 gpurun_out/ncu_c4.log | cut -c1-300
