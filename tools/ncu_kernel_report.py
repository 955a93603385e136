"""Summary of one kernel's `ncu --set full` capture: headline metrics, stall reasons, pipe use,
executed-instruction histogram per CTA warp (needs --import-source on).
    python tools/ncu_kernel_report.py gpurun_out/x.ncu-rep [launch_index] [warps_per_cta]"""
import collections
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
wpc = int(sys.argv[3]) if len(sys.argv) > 3 else 16
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, r = rows[0], rows[1], rows[2 + which]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size',
        'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.sum']
for k in keys:
    if k in hdr:
        print("%-66s %s %s" % (k, r[hdr.index(k)][:90], units[hdr.index(k)]))
for i, h in enumerate(hdr):
    if h.startswith('sm__pipe_') and h.endswith('cycles_active.avg.pct_of_peak_sustained_active') or \
            (h.startswith('sm__inst_executed_pipe_') and h.endswith('.avg.pct_of_peak_sustained_active')):
        try:
            if float(r[i]) > 5:
                print("   pipe %-60s %.1f" % (h.split('.')[0], float(r[i])))
        except ValueError:
            pass
tens = []
for i, h in enumerate(hdr):
    if ('tensor' in h or 'tmem' in h) and not h.startswith('device__'):
        try:
            if float(r[i]) != 0.0 and ('.avg' in h or '.sum' in h) and 'peak_sustained_elapsed.per_second' not in h \
                    and not h.endswith('peak_sustained') and '.max' not in h and '.min' not in h:
                tens.append("   %-88s %s %s" % (h, r[i], units[i]))
        except ValueError:
            pass
if tens:
    print("tensor / TMEM counters (non-zero):")
    print("\n".join(tens[:24]))
items = []
for i, h in enumerate(hdr):
    if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and 'not_issued' not in h:
        try:
            items.append((float(r[i]), h))
        except ValueError:
            pass
print("stalls per issue:", ", ".join("%s %.2f" % (h.replace('smsp__average_warps_issue_stalled_', '').replace(
    '_per_issue_active.ratio', ''), v) for v, h in sorted(items, reverse=True)[:9]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr_i = [i for i, x in enumerate(rows) if x and x[0] == 'Address']
if hdr_i:
    lo = hdr_i[which] + 1
    hi = hdr_i[which + 1] - 1 if which + 1 < len(hdr_i) else len(rows)
    h2 = rows[hdr_i[which]]
    ia, isrc = h2.index('Instructions Executed'), h2.index('Source')
    ops, total = collections.Counter(), 0
    for x in rows[lo:hi]:
        try:
            n = int(x[ia])
        except (ValueError, IndexError):
            continue
        m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', x[isrc].strip())
        ops[m.group(2).split('.')[0] if m else '?'] += n
        total += n
    N = int(float(r[hdr.index('launch__grid_size')])) * wpc
    print("warp instructions %d = %.0f per CTA warp" % (total, total / N))
    print(", ".join("%s %.0f" % (op, n / N) for op, n in ops.most_common(32)))
