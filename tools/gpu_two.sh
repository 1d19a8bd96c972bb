#!/bin/bash
# 2-GPU box: the NCCL forms of the two-rank tests and the torchrun bench line (all workloads incl. the data-parallel training step)
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_train_dp.py tests/test_gpu_sharded_encode.py -x -q 2>&1 | tail -3
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
tail -c 400 gpurun_out/r2_bench_2gpu.err
python - <<'PY'
import json
try:
    line = [l for l in open('gpurun_out/r2_bench_2gpu.json').read().splitlines() if l.startswith('{')][0]
    d = json.loads(line)
    print('N', d['n_gpus'], 'value', d['value'], 'e2e', d['e2e']['value'], 'sustained', d.get('value_sustained'))
    for k, v in d.get('workloads', {}).items():
        print(k, v['value'], v.get('e2e', {}).get('value'), v.get('value_sustained'), v.get('ms_per_iteration'), v.get('allreduce'))
except Exception as e:
    print('bench parse failed', e)
PY
