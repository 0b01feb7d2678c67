mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -1
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/benchF_ref.json 2> gpurun_out/benchF_ref.err
cut -c1-200 gpurun_out/benchF_ref.json
timeout 400 python bench.py > gpurun_out/benchF_c1.json 2> gpurun_out/benchF_c1.err
python -c "
import json;d=json.load(open('gpurun_out/benchF_c1.json'));print('c1 value %.4g e2e %.4g frac %.3f launches %d cpu %.4g on %d cores clocks %s' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'], d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['clocks']))"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:push_kernel -s 1 -c 1 -o gpurun_out/prof_c1h python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c1h.log 2>&1
tail -1 gpurun_out/ncu_c1h.log | cut -c1-100
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c1h.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_l_c1h.log 2>&1
timeout 200 python bench.py --strict 1 --no-cpu-baseline --steps 2 --warmup 1 --nptl 200000 > gpurun_out/benchF_strict.json 2>/dev/null
python -c "
import json;d=json.load(open('gpurun_out/benchF_strict.json'));print('strict build c1 value %.4g' % d['value'])"
