#!/bin/bash
for b in 19 38 57; do
for wl in plain_nfs bmcnet_nfs; do
timeout 600 python bench.py --workload $wl --steps 40 --warmup 5 --cpu-steps 2 --batch $b > gpurun_out/bb.json 2> gpurun_out/bb.err || tail -3 gpurun_out/bb.err
python - <<PY
import json
d=json.load(open('gpurun_out/bb.json'))
print('$wl B=$b value %.0f e2e %.0f ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
PY
done; done
