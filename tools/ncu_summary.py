"""Key metrics of one ncu --set full capture (raw page) as text: python tools/ncu_summary.py file.ncu-rep"""
import csv
import re
import subprocess
import sys

PAT = re.compile(r'^(gpu__time_duration\.sum|sm__cycles_active\.avg|sm__cycles_elapsed\.avg|'
                 r'sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)|'
                 r'sm__mem_tensor_cycles_active\.avg\.pct_of_peak_sustained_active|'
                 r'dram__bytes_(read|write)\.sum|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|'
                 r'lts__throughput\.avg\.pct_of_peak_sustained_elapsed|lts__t_sector_hit_rate\.pct|'
                 r'l1tex__m_xbar2l1tex_read_bytes\.sum|launch__registers_per_thread|launch__grid_size|launch__block_size|'
                 r'launch__shared_mem_per_block_(dynamic|static)|smsp__issue_active\.avg\.pct_of_peak_sustained_active|'
                 r'smsp__inst_executed\.sum|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum)$')
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
head, unit = rows[0], rows[1]
for r in rows[2:]:
    name = dict(zip(head, r)).get('Kernel Name', '?')
    print('kernel:', name)
    for h, u, v in zip(head, unit, r):
        if PAT.match(h):
            print('  %-72s %-10s %s' % (h, u, v))
