"""Instruction histogram per kernel of libbmc_b200.so from `cuobjdump -sass` (no GPU needed):
    python tools/sass_histogram.py > profiles/r02_sass_histogram.md
Lists, per kernel, the counts of the mnemonics that prove the Blackwell paths (UTCHMMA = tcgen05.mma, LDTM / STTM =
tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce, UTCBAR = tcgen05.commit, SYNCS = mbarrier,
ATOMS / ATOMG / RED = atomics) plus the five most frequent other mnemonics."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'bmcnet_esr_b200', 'libbmc_b200.so')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
KEY = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTMAREDG', 'UTMAPF', 'UTCBAR', 'SYNCS', 'STSM', 'LDSM',
       'ATOMS', 'ATOMG', 'RED', 'REDG', 'HMMA', 'FFMA', 'LDG', 'STG', 'LDS', 'STS']
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r'\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and kern:
        hist[kern][m.group(1).split('.')[0]] += 1


def demangle(n):
    try:
        d = subprocess.run(['c++filt', n], capture_output=True, text=True).stdout.strip()
    except Exception:
        d = n
    d = re.sub(r'bmc::\(anonymous namespace\)::', '', d)
    d = re.sub(r'\(.*', '', d)
    return d.replace('void ', '')


print('# SASS instruction histogram per kernel (`cuobjdump -sass %s`, sm_100a)' % os.path.relpath(lib, ROOT))
print()
print('UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG / UTMAREDG = TMA tensor load / store / reduce, UTCBAR = tcgen05.commit,')
print('SYNCS = mbarrier ops, STSM = stmatrix, ATOMS / ATOMG / RED = shared / global atomics.  Counts are static instructions.')
print()
print('| kernel | instr | ' + ' | '.join(KEY) + ' | other top-5 |')
print('|---|---|' + '---|' * len(KEY) + '---|')
for k, h in hist.items():
    tot = sum(h.values())
    if tot == 0:
        continue
    rest = [(n, c) for n, c in h.most_common() if n not in KEY][:5]
    print('| `%s` | %d | ' % (demangle(k)[:70], tot) + ' | '.join(str(h.get(x, 0)) if h.get(x, 0) else '' for x in KEY)
          + ' | ' + ', '.join('%s %d' % r for r in rest) + ' |')
tot = collections.Counter()
for h in hist.values():
    tot.update(h)
print()
print('library totals: ' + ', '.join('%s %d' % (x, tot[x]) for x in KEY if tot[x]))
