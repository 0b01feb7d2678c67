# cell-sorted particles on C5 (256^3) at the bench density and at the density of the full-size config
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "c5 or 3d or full_size" 2>&1 | tail -2
run() { # name nptl env...
  name=$1; nptl=$2; shift 2
  env "$@" timeout 300 python bench.py --workload c5 --grid 256 --nptl $nptl --no-cpu-baseline --steps 3 --warmup 2 > gpurun_out/sort_c5_$name.json 2>> gpurun_out/sort_c5.err
  python -c "
import json;d=json.load(open('gpurun_out/sort_c5_$name.json'));b=d['breakdown_ms_per_step'];print('c5 $name value %.4g push_ms %.2f mover_ms %.2f e2e %.4g' % (d['value'], b['push_ms'], b['mover_ms'], d['e2e']['value']))"
}
for nptl in 2000000 16000000; do
  run n${nptl}_nosort $nptl GPAT_PUSH_SORT=0
  run n${nptl}_sort_cap2 $nptl GPAT_PUSH_SORT=1
  run n${nptl}_sort_cap3 $nptl GPAT_PUSH_SORT=1 GPAT_PUSH_MAXCTAS=3
  run n${nptl}_sort_cap4 $nptl GPAT_PUSH_SORT=1 GPAT_PUSH_MAXCTAS=4
done
