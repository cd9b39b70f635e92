"""Summarises .ncu-rep captures (ncu --set full) into small text files for profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [...] --out profiles/r01"""
import csv, io, subprocess, sys, os

KEYS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
    'SM_C.TriageCompute.smsp__pipe_tensor_subpipe_dmma_cycles_active.avg',
    'sm__cycles_elapsed.max', 'smsp__cycles_active.avg',
    'TPC.TriageCompute.sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
    'SM_A.TriageCompute.sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
]

def summarise(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
        out.append('kernel: %s' % d.get('Kernel Name', '?'))
        for k in KEYS:
            if k in d:
                out.append('  %-100s %s %s' % (k, d[k], u.get(k, '')))
    return '\n'.join(out)

if __name__ == '__main__':
    args = sys.argv[1:]
    outp = None
    if '--out' in args:
        i = args.index('--out'); outp = args[i + 1]; args = args[:i] + args[i + 2:]
    for p in args:
        txt = '# ncu --set full --clock-control none --import-source on ; source: %s\n%s\n' % (os.path.basename(p), summarise(p))
        if outp:
            name = os.path.splitext(os.path.basename(p))[0].replace('prof_', '')
            with open('%s_%s.txt' % (outp, name), 'w') as f:
                f.write(txt)
        print(txt)
