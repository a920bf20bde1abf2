"""Summarise an `ncu --set full` report (.ncu-rep) into the text blocks kept under profiles/: one block per captured launch with the
metrics the roofline discussion uses.   python profiles/ncu_full_summary.py gpurun_out/x.ncu-rep "header line" > profiles/x.txt"""
import csv
import io
import subprocess
import sys

KEEP = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'sm__cycles_active.avg', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'launch__cluster_size', 'launch__waves_per_multiprocessor']


def main():
    rep, header = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else '')
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    names, units = rows[0], rows[1]
    if header:
        print('# ' + header)
    print('# (one B200, after the same command exited 0 without ncu; cold caches, kernels serialised)')
    for r in rows[2:]:
        print('----')
        for k in KEEP:
            if k in names:
                i = names.index(k)
                print(f'{k} [{units[i]}] = {r[i]}')


if __name__ == '__main__':
    main()
