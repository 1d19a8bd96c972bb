#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -20
python - <<'PY'
import os, torch
p = torch.cuda.get_device_properties(0)
print('pci', getattr(p,'pci_domain_id',None), getattr(p,'pci_bus_id',None), getattr(p,'pci_device_id',None))
print('affinity', len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:8], '... cpu_count', os.cpu_count())
import glob
for n in sorted(glob.glob('/sys/devices/system/node/node*')):
    print(n, open(n + '/cpulist').read().strip())
try:
    bus = '%04x:%02x:%02x.0' % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
    print(bus, 'numa_node', open('/sys/bus/pci/devices/%s/numa_node' % bus).read().strip())
except Exception as e: print('numa lookup failed', e)
PY
for i in 1 2 3 4; do timeout 300 python bench.py --steps 50 --warmup 5 --cpu-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); print('run $i value %.0f e2e %.0f' % (d['value'], d['e2e']['value']))"; done
