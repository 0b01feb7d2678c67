mkdir -p gpurun_out
run() { # name env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --workload c5 --grid 256 --nptl 2000000 --no-cpu-baseline --steps 4 --warmup 2 > gpurun_out/occ_c5c_$name.json 2>> gpurun_out/occ_c5c.err
  python -c "
import json;d=json.load(open('gpurun_out/occ_c5c_$name.json'));print('c5 $name value %.4g push_ms %.2f' % (d['value'], d['breakdown_ms_per_step']['push_ms']))"
}
run cap2 GPAT_PUSH_MAXCTAS=2
run cap1 GPAT_PUSH_MAXCTAS=1
run cap3 GPAT_PUSH_MAXCTAS=3
run pad60k GPAT_PUSH_SMEM_PAD=60000
run pad75k GPAT_PUSH_SMEM_PAD=75000
