#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "rc=$?"
tail -c 600 gpurun_out/bench_2gpu.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_2gpu.json') if l.startswith('{')][-1])
print('2gpu value %.0f e2e %.0f ms/step %.3f n_gpus %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['n_gpus']))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 6 --warmup 1 | tail -c 700
