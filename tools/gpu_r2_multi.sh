#!/bin/bash
# gpurun --gpus 2 -- 'bash tools/gpu_r2_multi.sh': the driver's default bench command at N = 2 (all workloads incl. the
# data-parallel graphed training iteration) + the 2-process GPU tests
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m pytest tests/test_gpu_train_dp.py tests/test_gpu_sharded_encode.py -m gpu -q 2>&1 | tail -n 3
S=$(date +%s)
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; echo "bench rc=$? wall $(( $(date +%s) - S )) s"
tail -c 600 gpurun_out/r2_bench_${N}gpu.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_${N}gpu.json') if l.startswith('{')][-1])
print('N', d['n_gpus'], 'value %.0f e2e %.0f' % (d['value'], d['e2e']['value']))
for k, v in d['workloads'].items():
    print(k, {kk: (round(vv, 1) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ('value', 'ms_per_step', 'ms_per_iteration', 'ms_per_iteration_eager', 'batch_per_gpu', 'loss_finite')}, v.get('allreduce'))
PY
