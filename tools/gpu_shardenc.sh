#!/bin/bash
# gpurun --gpus 2 -- 'bash tools/gpu_shardenc.sh'
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/shardenc_gpus.txt
timeout 600 python -m pytest tests/test_gpu_sharded_encode.py tests/test_gpu_encoders.py -m gpu -q > gpurun_out/shardenc_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/shardenc_pytest.log
timeout 300 python tools/bench_enc_sharded.py > gpurun_out/shardenc_1.json 2> gpurun_out/shardenc_1.err; echo "n1 rc=$?"
N=$(nvidia-smi -L | wc -l)
if [ "$N" -ge 2 ]; then
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/bench_enc_sharded.py > gpurun_out/shardenc_n.json 2> gpurun_out/shardenc_n.err; echo "nN rc=$?"
fi
cat gpurun_out/shardenc_1.json gpurun_out/shardenc_n.json 2>/dev/null
