"""tools/ncu_extract.py <report.ncu-rep> -- selected raw metrics of the first kernel of an `ncu --set full` report as CSV
(metric,unit,value); the committed form of a capture under profiles/."""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sectors.sum', 'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sector_hit_rate.pct', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_uniform.sum',
        'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__cycles_active.avg', 'sm__inst_executed_pipe_tmem',
        'lts__t_sector_hit_rate.pct']


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    name = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
    print('# kernel: %s' % name)
    print('metric,unit,value')
    for h, u, v in zip(hdr, units, vals):
        if any(h == w or h.startswith(w) for w in WANT) and v != '':
            print('%s,%s,%s' % (h, u, v))


if __name__ == '__main__':
    main(sys.argv[1])
