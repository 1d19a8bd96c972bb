#!/bin/bash
# gpurun --gpus N -- 'bash tools/gpu_multi2.sh'   : configs[3] (BMCNet EventZoom, sequences sharded over N GPUs),
# the NCCL form of the split-recording encoder test and its bench.
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 300 python -m pytest tests/test_gpu_sharded_encode.py -m gpu -q 2>&1 | tail -n 2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --workload bmcnet_eventzoom --steps 50 --warmup 5 > gpurun_out/bench_ez_${N}gpu.json 2> gpurun_out/bench_ez_${N}gpu.err; echo "ez rc=$?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 tools/bench_enc_sharded.py > gpurun_out/shardenc_${N}.json 2> gpurun_out/shardenc_${N}.err; echo "enc rc=$?"
grep "^{" gpurun_out/shardenc_${N}.json
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_ez_${N}gpu.json') if l.startswith('{')][-1])
print('EZ %d gpu: B=%d value %.0f e2e %.0f ms/step %.3f tflops %.0f clocks %s' % (d['n_gpus'], d['config']['batch_per_gpu'], d['value'], d['e2e']['value'], d['ms_per_step'], d['model_tflops'], d['clocks']['sm_mhz']))
PY
