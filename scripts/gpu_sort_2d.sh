mkdir -p gpurun_out
run() { # name workload-args... -- env
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 400 python bench.py "$@" --no-cpu-baseline > gpurun_out/sort2d_$name.json 2>> gpurun_out/sort2d.err
  python -c "
import json;d=json.load(open('gpurun_out/sort2d_$name.json'));b=d['breakdown_ms_per_step'];print('$name value %.4g push_ms %.2f mover_ms %.2f e2e %.4g' % (d['value'], b['push_ms'], b['mover_ms'], d['e2e']['value']))"
}
run c5_default X=1 -- --workload c5 --grid 256 --nptl 2000000 --steps 4 --warmup 2
run c1_sort GPAT_PUSH_SORT=1 -- --steps 4 --warmup 3
run c1_nosort GPAT_PUSH_SORT=0 -- --steps 4 --warmup 3
run c2_sort GPAT_PUSH_SORT=1 -- --workload c2 --nptl 2000000 --steps 2 --warmup 1
run c4_sort GPAT_PUSH_SORT=1 -- --workload c4 --nptl 300000 --steps 1 --warmup 1
