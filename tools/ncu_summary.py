"""Condense an .ncu-rep (read here with `ncu -i`) into a small JSON of the metrics the design
discussion uses: duration, DRAM bytes, pipe/L1/L2 utilisation, occupancy, registers, stall mix."""
import csv
import json
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_static', 'launch__shared_mem_per_block_dynamic',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__cycles_elapsed.max', 'smsp__inst_executed.sum', 'sm__sass_thread_inst_executed_op_ffma_pred_on.sum']


def main(path, out=None):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        rec = {'kernel': vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''}
        for i, h in enumerate(hdr):
            if h in KEYS or 'issue_stalled' in h and h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio'):
                try:
                    rec[h + (' [%s]' % units[i] if units[i] else '')] = float(vals[i].replace(',', ''))
                except ValueError:
                    rec[h] = vals[i]
        res.append(rec)
    s = json.dumps(res, indent=1)
    if out:
        open(out, 'w').write(s)
    print(s)


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
