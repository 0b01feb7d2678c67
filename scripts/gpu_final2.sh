mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/benchJ_c1.json 2> gpurun_out/benchJ_c1.err
python -c "
import json;d=json.load(open('gpurun_out/benchJ_c1.json'));print('c1 value %.4g e2e %.4g frac %.3f launches %d cpu %.4g on %d cores clocks %s' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'], d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['clocks']))"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:push_kernel -s 1 -c 1 -o gpurun_out/prof_c1j python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c1j.log 2>&1
tail -1 gpurun_out/ncu_c1j.log | cut -c1-100
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c1j.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_l_c1j.log 2>&1
for spec in "c2 --nptl 2000000" "c3" "c4 --nptl 300000 --steps 1" "c5 --grid 256 --nptl 2000000"; do
  set -- $spec; wl=$1; shift
  timeout 600 python bench.py --workload $wl --steps 2 --warmup 1 "$@" --no-cpu-baseline > gpurun_out/benchJ_$wl.json 2> gpurun_out/benchJ_$wl.err
  python -c "
import json;d=json.load(open('gpurun_out/benchJ_$wl.json'));print('$wl value %.4g e2e %.4g frac %.3f push_ms %.1f' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['breakdown_ms_per_step']['push_ms']))"
done
