"""Extracts the metrics quoted in DESIGN.md / bench.py from ncu reports: python tools/summarize_ncu.py <rep> <out.json>"""
import csv, json, subprocess, sys, io
KEEP = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'sm__inst_executed.avg.per_cycle_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__waves_per_multiprocessor', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__shared_mem_per_block_dynamic',
        'launch__shared_mem_per_block_static', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem']
def main(rep, out):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    res = []
    for r in data:
        d = {k: (r[hdr.index(k)] + ' ' + units[hdr.index(k)]).strip() for k in KEEP if k in hdr}
        for k in hdr:
            if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('per_issue_active.ratio'):
                v = float(r[hdr.index(k)] or 0)
                if v >= 0.3:
                    d[k.replace('smsp__average_warps_issue_stalled_', 'stall_').replace('_per_issue_active.ratio', '')] = round(v, 2)
        res.append(d)
    json.dump(res, open(out, 'w'), indent=1)
    return res
if __name__ == '__main__':
    for d in main(sys.argv[1], sys.argv[2]):
        print({k: d[k] for k in ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed') if k in d})
