mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -1
grep -E "^FAILED|^E  " gpurun_out/pytest_gpu.log | head -6
timeout 300 python bench.py --workload c5 --grid 256 --nptl 2000000 --no-cpu-baseline --steps 6 --warmup 2 > gpurun_out/bench9_c5.json 2> gpurun_out/bench9_c5.err
python -c "
import json;d=json.load(open('gpurun_out/bench9_c5.json'));print('c5 value %.4g push_ms %.2f' % (d['value'], d['breakdown_ms_per_step']['push_ms']))"
