"""Instructions executed per CUDA source line, per kernel, from an .ncu-rep captured with --import-source on (-lineinfo build):
    python tools/ncu_lines.py <file.ncu-rep> [min_percent]"""
import collections
import csv
import io
import subprocess
import sys

out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
per_kernel = collections.OrderedDict()
path = func = None
hdr = None
for row in csv.reader(io.StringIO(out)):
    if not row:
        continue
    if row[0] == 'File Path':
        path = row[1].split('/')[-1]
    elif row[0] == 'Function Name':
        func = row[1].split('(')[0].replace('void nrp::', '')
    elif row[0] == 'Line No':
        hdr = row
        i_exec = hdr.index('Instructions Executed')
        i_samp = hdr.index('# Samples')
    elif hdr and row[0] not in ('', '-'):
        try:
            n = int(row[i_exec])
        except ValueError:
            continue
        d = per_kernel.setdefault(func, {})
        key = (path, int(row[0]), row[1].strip()[:110])
        cur = d.get(key, (0, 0))
        d[key] = (cur[0] + n, cur[1] + (int(row[i_samp]) if row[i_samp].isdigit() else 0))
seen = set()
for func, d in per_kernel.items():
    total = sum(v[0] for v in d.values())
    samples = sum(v[1] for v in d.values())
    sig = (func, total)
    if sig in seen:
        continue
    seen.add(sig)
    print('== %s: %d warp instructions, %d samples' % (func, total, samples))
    for (path, line, text), (n, smp) in sorted(d.items(), key=lambda kv: -kv[1][0]):
        if 100.0 * n / total < min_pct:
            break
        print('  %5.1f%% instr %5.1f%% stall  %s:%d  %s' % (100.0 * n / total, 100.0 * smp / max(samples, 1), path, line, text))
