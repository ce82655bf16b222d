"""gpurun_out/ (written by profiles/ncu_capture.sh) -> profiles/<name>/: per-kernel metric summaries of the
`ncu --set full` captures (metric,unit,value rows), the launch list and per-kernel totals of the launch-list pass.

    python profiles/summarize_ncu.py r02_ncu_v1
"""
import csv
import os
import re
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, 'gpurun_out')
KEEP = re.compile(r'^(Kernel Name|dram__|lts__t_(bytes|sector_hit|sectors_srcunit_tex_op_read)|l1tex__m_xbar2l1tex_read_bytes|'
                  r'gpu__time_duration|gpc__cycles_elapsed\.avg\.per_second|sm__cycles_(active|elapsed)\.avg$|'
                  r'sm__throughput|sm__warps_active|sm__inst_executed_pipe_(fp64|tensor|alu|fma|lsu|uniform)|'
                  r'sm__pipe_(fp64|tensor|fmaheavy|alu)|sm__ops_path_tensor|sm__mem_tensor_cycles_active\.avg|'
                  r'smsp__inst_executed\.sum|smsp__issue_active|smsp__average_warps?_issue_stalled|'
                  r'smsp__sass_thread_inst_executed_op_(dfma|dadd|dmul)|launch__|sm__maximum_warps)')


def summary(kernel, out_dir):
    path = os.path.join(SRC, 'prof_%s_raw.csv' % kernel)
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    with open(os.path.join(out_dir, kernel + '_summary.csv'), 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['metric', 'unit', 'value'])
        for h, u, v in zip(hdr, units, vals):
            if KEEP.match(h) and not re.search(r'\.(max|min)($|\.)|\.sum\.(pct|per|peak)|peak_sustained$|per_cycle', h):
                w.writerow([h, u, v])


def launches(out_dir):
    path = os.path.join(SRC, 'launches.csv')
    if not os.path.exists(path):
        return
    lines = [l for l in open(path) if not l.startswith('==')]
    rows = list(csv.DictReader(lines))
    tot = OrderedDict()
    with open(os.path.join(out_dir, 'launches.csv'), 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['id', 'kernel', 'grid', 'block', 'duration'])
        for r in rows:
            if r.get('Metric Name') != 'gpu__time_duration.sum':
                continue
            name = re.sub(r'\(.*', '', r['Kernel Name'])[:72]
            try:
                val = float(r['Metric Value'].replace(',', ''))
            except ValueError:
                continue
            unit = r.get('Metric Unit', '')
            ms = val * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(unit, 1e-6)
            w.writerow([r['ID'], name, r['Grid Size'], r['Block Size'], '%.4f ms' % ms])
            t = tot.setdefault(name, [0, 0.0])
            t[0] += 1
            t[1] += ms
    total = sum(t[1] for t in tot.values()) or 1.0
    with open(os.path.join(out_dir, 'launch_totals.csv'), 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['kernel', 'launches', 'total_ms', 'share'])
        for name, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            w.writerow([name, n, '%.4f' % ms, '%.4f' % (ms / total)])


def main():
    out_dir = os.path.join(ROOT, 'profiles', sys.argv[1])
    os.makedirs(out_dir, exist_ok=True)
    for k in sys.argv[2:] or ['k_var_ozaki', 'k_predict', 'k_sobol_static']:
        summary(k, out_dir)
    launches(out_dir)


if __name__ == '__main__':
    main()
