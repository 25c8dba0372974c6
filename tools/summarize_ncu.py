"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.

    python tools/summarize_ncu.py <tag>        # reads gpurun_out/launches_<tag>.csv and gpurun_out/prof_<tag>.ncu-rep
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def launch_summary(tag):
    path = os.path.join(ROOT, 'gpurun_out', 'launches_%s.csv' % tag)
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    hdr, data = rows[hi], rows[hi + 1:]
    idx = {n: i for i, n in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in data:
        if len(r) < len(hdr):
            continue
        name = r[idx['Kernel Name']].split('(')[0].replace('void ', '')
        metric, unit = r[idx['Metric Name']], r[idx['Metric Unit']]
        value = float(r[idx['Metric Value']].replace(',', ''))
        d = agg.setdefault(name, {'n': 0, 'us': 0.0, 'rd': 0.0, 'wr': 0.0})
        if metric == 'gpu__time_duration.sum':
            d['n'] += 1
            d['us'] += {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(unit, 1e-3) * value
        elif metric == 'dram__bytes_read.sum':
            d['rd'] += value * UNIT.get(unit, 1)
        elif metric == 'dram__bytes_write.sum':
            d['wr'] += value * UNIT.get(unit, 1)
    total = sum(d['us'] for d in agg.values())
    lines = ['# ncu launch list (%s): gpu__time_duration.sum per kernel, cold cache, serialised -- compare SHARES' % tag,
             '# total %.1f us over %d launches' % (total, sum(d['n'] for d in agg.values())),
             '%-52s %5s %10s %7s %10s %10s' % ('kernel', 'n', 'us', 'share', 'dram rd MB', 'dram wr MB')]
    for name, d in sorted(agg.items(), key=lambda kv: -kv[1]['us']):
        lines.append('%-52s %5d %10.1f %6.1f%% %10.1f %10.1f' % (name[:52], d['n'], d['us'], 100 * d['us'] / total, d['rd'] / 1e6, d['wr'] / 1e6))
    return '\n'.join(lines) + '\n', agg


WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__grid_size', 'launch__block_size',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio']


def full_summary(tag):
    rep = os.path.join(ROOT, 'gpurun_out', 'prof_%s.ncu-rep' % tag)
    if not os.path.exists(rep):
        return None
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {n: i for i, n in enumerate(hdr)}
    lines = ['# ncu --set full --clock-control none (%s): selected metrics per captured launch' % tag]
    for r in data:
        lines.append('\n== ' + r[idx['Kernel Name']][:100])
        for w in WANT:
            if w in idx:
                lines.append('   %-88s %14s %s' % (w, r[idx[w]], units[idx[w]]))
    return '\n'.join(lines) + '\n'


def main():
    tag = sys.argv[1]
    text, agg = launch_summary(tag)
    os.makedirs(os.path.join(ROOT, 'profiles'), exist_ok=True)
    open(os.path.join(ROOT, 'profiles', 'ncu_launches_%s.txt' % tag), 'w').write(text)
    full = full_summary(tag)
    if full:
        open(os.path.join(ROOT, 'profiles', 'ncu_full_%s.txt' % tag), 'w').write(full)
    # per-launch DRAM traffic of the bulk kernels for bench.py's roofline.traffic
    traffic = {}
    for name, d in agg.items():
        short = name.split('::')[-1].split('<')[0]
        if short.startswith('k_filter_warp'):
            short = 'k_filter_warp'                  # every version of the warp-private filter under one name (bench.py looks it up)
        if short in ('k_hist_from_ascii', 'k_split_uniform', 'k_split_from_ascii', 'k_split_from_records', 'k_filter', 'k_filter_bloom', 'k_filter_warp', 'k_flag_windows') and d['n']:
            traffic[short] = int((d['rd'] + d['wr']) / d['n'])
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    blob = json.load(open(tpath)) if os.path.exists(tpath) else {}
    blob[sys.argv[2] if len(sys.argv) > 2 else 'c2'] = traffic
    blob['_source'] = 'dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu launch list ' + tag
    json.dump(blob, open(tpath, 'w'), indent=1)
    print(text)


if __name__ == '__main__':
    main()
