#!/bin/bash
# Round 2, call J9 (8 GPUs), final build: bench.py C1 at N = 8 under torchrun (weak line + strong-scaling pass + all-reduce check)
mkdir -p gpurun_out
T=r02j9
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/${T}_bench_c1_n8.json 2> gpurun_out/${T}_bench_c1_n8.err
python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench_c1_n8.json').read().strip().splitlines()[-1]);print('N=8 value %.4g e2e %.4g' % (d['value'], d['e2e']['value']), d.get('allreduce_check'), {k: d['strong_scaling'][k] for k in ('value', 'e2e', 'ms_per_step')})" || tail -5 gpurun_out/${T}_bench_c1_n8.err
