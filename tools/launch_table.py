"""One step of an ncu launch list (tools/gpu_launchlist.sh) as a table: kernel, time, DRAM bytes;
sums for the step and for the fused filter + threshold + detect stage.
    python tools/launch_table.py profiles/r02_launches_c2.csv"""
import csv
import sys


def table(path):
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith('==')]
    hi = [i for i, r in enumerate(rows) if r[0] == 'ID'][0]
    hdr = rows[hi]
    ix = {n: i for i, n in enumerate(hdr)}
    cur, order = {}, []
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        k = r[ix['ID']]
        if k not in cur:
            cur[k] = {'name': r[ix['Kernel Name']].replace('void ', '').replace('<unnamed>::', '')[:56]}
            order.append(k)
        v = float(r[ix['Metric Value']].replace(',', ''))
        f = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1, 'ms': 1e3}.get(r[ix['Metric Unit']], 1)
        cur[k][r[ix['Metric Name']]] = v * f
    names = [cur[k]['name'] for k in order]
    first = next(i for i, n in enumerate(names) if n.startswith('filt_summary') or n.startswith('filt_span'))
    last = next(i for i in range(first + 1, len(names)) if names[i].startswith('pca_project')
                or (names[i].startswith('filt_summary') or (names[i].startswith('filt_span') and i > first + 8)))
    if not names[last].startswith('pca_project'):
        last -= 1
    out, tot, sb, st, in_stage = [], 0.0, 0.0, 0.0, True
    for k in order[first:last + 1]:
        d = cur[k]
        t = d.get('gpu__time_duration.sum', 0)
        rd, wr = d.get('dram__bytes_read.sum', 0), d.get('dram__bytes_write.sum', 0)
        if d['name'].startswith('align'):
            in_stage = False
        if in_stage:
            sb += rd + wr
            st += t
        tot += t
        out.append("%-58s %9.1f us  read %8.1f MB  written %8.1f MB" % (d['name'], t, rd / 1e6, wr / 1e6))
    out.append("kernel time of the step %.1f us; filter + threshold + detect stage %.1f us, %.3f GB of DRAM traffic"
               % (tot, st, sb / 1e9))
    return "\n".join(out)


if __name__ == "__main__":
    print(table(sys.argv[1]))
