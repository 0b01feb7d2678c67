for cap in 2 3; do
GPAT_LIB=$PWD/stochastic_parker_b200/csrc/libgpat_cuda.${GPAT_VARIANT:-mb3}.so GPAT_PUSH_MAXCTAS=$cap timeout 300 python bench.py --workload c5 --grid 256 --nptl 2000000 --no-cpu-baseline --steps 4 --warmup 2 > gpurun_out/occ_c5d_$cap.json 2>> gpurun_out/occ_c5d.err
python -c "
import json;d=json.load(open('gpurun_out/occ_c5d_$cap.json'));print('c5 ${GPAT_VARIANT:-mb3} cap $cap value %.4g push_ms %.2f' % (d['value'], d['breakdown_ms_per_step']['push_ms']))"
done
