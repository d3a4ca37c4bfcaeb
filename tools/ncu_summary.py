"""Dev tool: selected metrics of an `ncu --set full` report, one line per captured launch.
usage: ncu -i X.ncu-rep --page raw --csv | python tools/ncu_summary.py [out.txt]"""
import csv
import sys

WANT = ['Kernel Name', 'Grid Size', 'gpu__time_duration.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio']
rows = list(csv.reader(sys.stdin))
h = rows[0]
idx = [(w, h.index(w)) for w in WANT if w in h]
lines = ['ncu --set full --clock-control none (raw page, selected metrics; one line per captured launch; units: ' +
         ', '.join(f'{w.split("__")[-1][:28]}[{rows[1][i]}]' for w, i in idx if rows[1][i]) + ')']
for r in rows[2:]:
    lines.append(' | '.join('%s=%s' % (w.replace('smsp__average_warps_issue_stalled_', 'stall_').replace('_per_issue_active.ratio', '')
                                       .replace('.avg.pct_of_peak_sustained_elapsed', '%').replace('.avg.pct_of_peak_sustained_active', '%act'),
                                       r[i][:70]) for w, i in idx))
txt = '\n'.join(lines) + '\n'
if len(sys.argv) > 1:
    open(sys.argv[1], 'w').write(txt)
sys.stdout.write(txt)
