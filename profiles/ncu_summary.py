"""Print the headline metrics of every kernel in an .ncu-rep (raw page) -- used to write profiles/*.md."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg', 'sm__cycles_elapsed.avg',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_lsu.sum',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor']


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print('==', d.get('Kernel Name', '?')[:110])
        for k in KEYS:
            if k in d:
                print(f'   {k} = {d[k]} {units[hdr.index(k)]}')
        for k in hdr:
            if k.startswith('smsp__average_warp') and 'issue_stalled' in k and k.endswith('.ratio'):
                try:
                    v = float(d[k])
                except ValueError:
                    continue
                if v > 0.5:
                    print(f'   stall {k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")} = {v:.2f}')


if __name__ == '__main__':
    main(sys.argv[1])
