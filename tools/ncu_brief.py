"""Print the headline metrics and top stall reasons of every kernel in an .ncu-rep:  python tools/ncu_brief.py <file.ncu-rep>"""
import csv
import io
import subprocess
import sys

out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_shared_atom.sum', 'lts__t_sector_hit_rate.pct']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('==', d['Kernel Name'][:70])
    for w in want:
        if w in d:
            print('   %-70s %s' % (w, d[w]))
    st = sorted(((float(d[h].replace(',', '')) if d[h] else 0, h) for h in hdr
                 if 'warps_issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h), reverse=True)[:7]
    print('   stalls per issue: ' + ', '.join('%s %.2f' % (h.split('stalled_')[1].split('_per')[0], v) for v, h in st))
