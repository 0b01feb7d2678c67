mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -1
timeout 400 python bench.py > gpurun_out/bench7_c1.json 2> gpurun_out/bench7_c1.err
python -c "
import json;d=json.load(open('gpurun_out/bench7_c1.json'));print('c1 value %.4g e2e %.4g frac %.3f cpu %.4g on %d cores' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['cpu_baseline']['cores']))"
for spec in "c2 --nptl 2000000" "c3" "c4 --nptl 1000000" "c5 --grid 256 --nptl 2000000"; do
  set -- $spec; wl=$1; shift
  timeout 600 python bench.py --workload $wl "$@" --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench7_$wl.json 2> gpurun_out/bench7_$wl.err
  python -c "
import json;d=json.load(open('gpurun_out/bench7_$wl.json'));print('$wl value %.4g e2e %.4g frac %.3f push_ms %.1f' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['breakdown_ms_per_step']['push_ms']))"
done
