# bench lines of the other named configs (parity-test cases; kept for profiles/, not the headline)
mkdir -p gpurun_out
for spec in "c2 --nptl 4000000" "c3" "c4 --nptl 2000000" "c5 --grid 256 --nptl 2000000"; do
  set -- $spec; wl=$1; shift
  timeout 600 python bench.py --workload $wl "$@" --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$wl.json'))
    r=d['roofline']; b=d['breakdown_ms_per_step']
    print('$wl', '%.4g steps/s' % d['value'], 'e2e %.4g' % d['e2e']['value'], 'algo GB/s %.0f frac %.3f' % (r['achieved'], r['frac']), 'push %.1f ms upload %.1f diag %.1f' % (b['push_ms'], b['upload_ms'], b['diag_ms']), d['config']['field_layout'])
except Exception as e:
    print('$wl failed', e, open('gpurun_out/bench_$wl.err').read()[-500:])
PY
done
