#!/bin/bash
# Round 2, call J (8 GPUs): C1 weak + strong scaling with the in-bench all-reduce value check, the reference arm under
# torchrun (thread count), and BASELINE config[4] as stated: 512^3 field per GPU, 1e9 particles over 8 GPUs.
mkdir -p gpurun_out
T=r02j
{ nvidia-smi -L; free -g | head -2; nproc; } > gpurun_out/${T}_box.log 2>&1; cat gpurun_out/${T}_box.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/${T}_bench_c1_n8.json 2> gpurun_out/${T}_bench_c1_n8.err
python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench_c1_n8.json').read().strip().splitlines()[-1]);print('N=8 value %.4g e2e %.4g' % (d['value'], d['e2e']['value']), d.get('allreduce_check'), d.get('strong_scaling'))" || tail -5 gpurun_out/${T}_bench_c1_n8.err
$TR --master-port 29522 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/${T}_ref_c1_n8.json 2> gpurun_out/${T}_ref_c1_n8.err
python -c "
import json;d=json.loads(open('gpurun_out/${T}_ref_c1_n8.json').read().strip().splitlines()[-1]);print('reference arm at N=8: value %.4g cores %d' % (d['value'], d['cpu_baseline']['cores']))" || tail -5 gpurun_out/${T}_ref_c1_n8.err
MEM=$(free -g | awk '/Mem:/{print $2}')
if [ "$MEM" -ge 400 ]; then
  timeout 1200 $TR --master-port 29523 bench.py --workload c5 --nptl 125000000 --gpus 8 --steps 1 --warmup 1 --no-membw --no-cpu-baseline > gpurun_out/${T}_full_c5_n8.json 2> gpurun_out/${T}_full_c5_n8.err
  python -c "
import json;d=json.loads(open('gpurun_out/${T}_full_c5_n8.json').read().strip().splitlines()[-1]);print('C5 x8 value %.4g e2e %.4g ms/step %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step']), d.get('allreduce_check'))" || tail -5 gpurun_out/${T}_full_c5_n8.err
else
  echo "host memory $MEM GB: C5 x8 (13 GB of pinned frames per rank) skipped"
fi
