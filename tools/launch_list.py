"""Reduce an `ncu --metrics gpu__time_duration.sum --csv` log of a bench.py run to the launches of
its LAST full train step (steps are delimited by a kernel launched once per step, default
specnorm_fwd_kernel), plus
a per-kernel summary with each kernel's share of the step.
    python tools/launch_list.py gpurun_out/x/launches_raw.csv[.gz] [delimiter kernel] [its launches per step] > profiles/rNN_launches_bench.csv"""
import csv
import sys

rows = []
import gzip
with (gzip.open(sys.argv[1], 'rt', newline='') if sys.argv[1].endswith('.gz') else open(sys.argv[1], newline='')) as f:
    lines = [l for l in f if not l.startswith('==')]
rd = csv.reader(lines)
head = next(rd)
ik, iv, iu = head.index('Kernel Name'), head.index('Metric Value'), head.index('Metric Unit')
for r in rd:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(',', ''))
    u = r[iu]
    us = v / 1e3 if u in ('ns', 'nsecond') else v if u in ('us', 'usecond') else v * 1e3 if u in ('ms', 'msecond') else v / 1e3
    rows.append((r[ik], us))
DELIM = sys.argv[2] if len(sys.argv) > 2 else 'specnorm_fwd_kernel'      # launched EVERY times per train step
EVERY = int(sys.argv[3]) if len(sys.argv) > 3 else 1
marks = [i for i, (k, _) in enumerate(rows) if DELIM in k]
if len(marks) >= EVERY + 1:
    step = rows[marks[-1 - EVERY]:marks[-1]]
else:
    step = rows
total = sum(us for _, us in step)
agg = {}
for k, us in step:
    name = k.split('(')[0][:110]
    a = agg.setdefault(name, [0.0, 0])
    a[0] += us
    a[1] += 1
w = csv.writer(sys.stdout)
w.writerow(['# one train step (one period of %s): %d launches, %.1f us serialised (ncu, cold cache, --clock-control none)' % (DELIM, len(step), total)])
w.writerow(['kernel', 'launches', 'us_total', 'share_of_step'])
for name, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    w.writerow([name, n, '%.1f' % us, '%.4f' % (us / total)])
w.writerow([])
w.writerow(['# launch list', 'us'])
for k, us in step:
    w.writerow([k.split('(')[0][:110], '%.2f' % us])
