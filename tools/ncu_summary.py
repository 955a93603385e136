"""Summarise ncu outputs brought back in gpurun_out/ (launch list CSV and .ncu-rep).

    python tools/ncu_summary.py launches gpurun_out/launches.csv
    python tools/ncu_summary.py report   gpurun_out/prof.ncu-rep
"""
import collections
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__inst_executed.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_lsu.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.max']


def launches(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        # launch lists taken with several metrics (duration + DRAM bytes) have one row per metric
        # and launch: only the durations belong in this table (tools/launch_table.py prints all)
        if row.get('Metric Name', 'gpu__time_duration.sum') != 'gpu__time_duration.sum':
            continue
        name = row['Kernel Name'].split('(')[0].replace('void ', '').replace('<unnamed>::', '')
        agg.setdefault(name, []).append(float(row['Metric Value'].replace(',', '')))
    tot = sum(sum(v) for v in agg.values())
    print("%-44s %5s %12s %8s" % ("kernel", "n", "mean us", "share"))
    for k, v in agg.items():
        print("%-44s %5d %12.1f %7.1f%%" % (k[:44], len(v), sum(v) / len(v) / 1000, 100 * sum(v) / tot))


def report(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('---- ' + r[hdr.index('Kernel Name')][:90])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print('  %-70s %s %s' % (k, r[i], units[i]))
        st = [h for h in hdr if 'issue_stalled' in h and h.endswith('per_issue_active.ratio')]
        vals = sorted(((float(r[hdr.index(h)].replace(',', '') or 0), h) for h in st), reverse=True)
        print('  stalls (warps per issue): ' + ', '.join(
            '%s %.2f' % (h.split('issue_stalled_')[1].split('_per_')[0], v) for v, h in vals[:6]))


if __name__ == "__main__":
    {'launches': launches, 'report': report}[sys.argv[1]](sys.argv[2])
