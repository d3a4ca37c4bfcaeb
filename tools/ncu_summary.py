"""Dev tool: compact text summary of an .ncu-rep (ncu --set full) for profiles/.
usage: python tools/ncu_summary.py report.ncu-rep > profiles/xxx.txt"""
import csv, subprocess, sys, io

KEYS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
    'lts__t_requests_srcunit_tex_op_red.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
    'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_subpipe_umma_cycles_active.avg.pct_of_peak_sustained_active' ,
    'sm__inst_executed_pipe_uniform.sum',
]


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    h, units, rows = r[0], r[1], r[2:]
    ix = {n: i for i, n in enumerate(h)}
    print(f'# {path}: ncu --set full --clock-control none; one column per captured launch')
    print('kernel'.ljust(72), [row[ix['Kernel Name']].split('(')[0][-40:] for row in rows])
    for k in KEYS:
        if k in ix:
            print((k + ' [' + units[ix[k]] + ']').ljust(72), [row[ix[k]] for row in rows])
    print('# warp stall reasons (smsp__average_warps_issue_stalled_*_per_issue_active, > 0.3)')
    for i, n in enumerate(h):
        if 'average_warps_issue_stalled' in n and n.endswith('per_issue_active.ratio'):
            vals = [row[i] for row in rows]
            try:
                if max(float(v) for v in vals) > 0.3:
                    print(n.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '').ljust(72), vals)
            except ValueError:
                pass
    for k in h:
        if 'tensor' in k and 'pct' in k and k not in KEYS:
            vals = [row[ix[k]] for row in rows]
            try:
                if max(float(v.replace(',', '')) for v in vals) > 0:
                    print((k + ' [' + units[ix[k]] + ']').ljust(72), vals)
            except ValueError:
                pass


if __name__ == '__main__':
    main(sys.argv[1])
