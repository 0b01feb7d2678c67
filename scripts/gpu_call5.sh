mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -1
for g in 1 0; do
GPAT_PUSH_GENERIC=$g timeout 300 python bench.py --no-cpu-baseline --steps 4 --warmup 3 > gpurun_out/bench5_c1_g$g.json 2> gpurun_out/bench5_c1_g$g.err
python -c "
import json;d=json.load(open('gpurun_out/bench5_c1_g$g.json'));print('c1 generic=$g value %.4g e2e %.4g push_ms %.2f' % (d['value'], d['e2e']['value'], d['breakdown_ms_per_step']['push_ms']))"
done
timeout 300 python bench.py --workload c3 --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/bench5_c3.json 2> gpurun_out/bench5_c3.err
python -c "
import json;d=json.load(open('gpurun_out/bench5_c3.json'));print('c3 value %.4g' % d['value'])"
timeout 300 python bench.py --workload c5 --grid 256 --nptl 2000000 --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/bench5_c5.json 2> gpurun_out/bench5_c5.err
python -c "
import json;d=json.load(open('gpurun_out/bench5_c5.json'));print('c5 value %.4g' % d['value'])"
