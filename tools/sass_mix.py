"""Opcode mix (dynamic warp instructions + stall samples) of one kernel from an ncu report's source page.
    python tools/sass_mix.py gpurun_out/prof_X.ncu-rep <kernel regex>"""
import collections
import csv
import re
import subprocess
import sys

out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '-k', 'regex:' + sys.argv[2]], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
start = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[start]
idx = {n: i for i, n in enumerate(hdr)}
ops, samples = collections.Counter(), collections.Counter()
static = 0
for r in rows[start + 1:]:
    if len(r) < len(hdr) or r[0] == 'Address' or r[0] == 'Kernel Name':
        continue
    ins = r[idx['Source']].strip()
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', ins)
    op = m.group(2).split('.')[0] if m else ins[:10]
    ops[op] += float(r[idx['Instructions Executed']])
    samples[op] += float(r[idx['# Samples']])
    static += 1
tot = sum(ops.values())
print('dynamic warp instructions %.3e, static %d' % (tot, static))
for op, v in ops.most_common(28):
    print('%-12s %6.2f%%   stall samples %5.1f%%' % (op, 100 * v / tot, 100 * samples[op] / max(1, sum(samples.values()))))
