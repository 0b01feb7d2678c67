# parity tests + prefetch/occupancy experiments on the HBM-resident configs
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
LOG=gpurun_out/exp2.log; : > $LOG
run() { # name lib prefetch workload args...
  name=$1; lib=$2; pf=$3; shift 3
  GPAT_LIB=$PWD/stochastic_parker_b200/csrc/$lib GPAT_PUSH_PREFETCH=$pf timeout 300 python bench.py --workload "$@" --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/exp2_$name.json 2>> $LOG
  python - <<PY >> $LOG
import json
try:
    d=json.load(open('gpurun_out/exp2_$name.json'))
    print('$name', '%.4g steps/s' % d['value'], 'frac %.3f' % d['roofline']['frac'], 'push_ms %.2f' % d['breakdown_ms_per_step']['push_ms'])
except Exception as e:
    print('$name failed', e)
PY
}
for lib in libgpat_cuda.so libgpat_cuda.mb4.so; do
  for pf in 0 1 2; do
    run c5_${lib#libgpat_cuda.}_pf$pf $lib $pf c5 --grid 256 --nptl 2000000
    run c4_${lib#libgpat_cuda.}_pf$pf $lib $pf c4 --nptl 60000
  done
done
run c1_pf0 libgpat_cuda.so 0 c1
run c1_pf1 libgpat_cuda.so 1 c1
run c2_pf0 libgpat_cuda.so 0 c2 --nptl 1000000
run c2_pf1 libgpat_cuda.so 1 c2 --nptl 1000000
cat $LOG
