#!/bin/bash
# Round-2 (second half) evidence: the split-bins encoder kernels (role_kernel), the encoder sweep, the graphed training
# iteration (kernel histogram of one replay, timings), ncu captures of role_kernel.  Summaries under gpurun_out/prof/.
mkdir -p gpurun_out/prof
P=gpurun_out/prof
NCU="ncu --clock-control none"
timeout 1000 python tools/enc_sweep.py > $P/enc_sweep.md 2> $P/enc_sweep.err; echo "sweep rc=$?"
timeout 300 python tools/split_bins_rate.py 4e8 > $P/split_bins_rate.txt 2>&1; echo "rate rc=$?"; cat $P/split_bins_rate.txt
timeout 600 $NCU --set full --import-source on -k regex:role_kernel -s 1 -c 1 -f -o $P/role_vox python tools/split_bins_prof.py voxel > /dev/null 2>&1; echo "rc=$?"
timeout 600 $NCU --set full --import-source on -k regex:role_kernel -s 1 -c 1 -f -o $P/role_cnt python tools/split_bins_prof.py counts > /dev/null 2>&1; echo "rc=$?"
for f in role_vox role_cnt; do python tools/ncu_summary.py $P/$f.ncu-rep > $P/ncu_full_$f.txt 2>&1; done
timeout 600 python tools/train_time.py 2 8 16 > $P/train_time.txt 2>&1; echo "train rc=$?"; grep "^{" $P/train_time.txt
timeout 600 python tools/train_profile.py 2 > $P/train_profile_b2.txt 2>&1
timeout 600 python tools/train_profile.py 8 > $P/train_profile_b8.txt 2>&1
head -12 $P/train_profile_b8.txt
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_encoders.py -q -x -k "cluster_bins_against_oracle or skewed" > $P/memcheck_roles.log 2>&1; echo "memcheck rc=$?"; tail -3 $P/memcheck_roles.log
