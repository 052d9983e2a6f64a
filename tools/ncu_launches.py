"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the LAST `frac` of the
launches (e.g. the second of two identical steps), optionally listing every launch whose name matches a pattern."""
import collections
import csv
import re
import sys

path = sys.argv[1]
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
pat = sys.argv[3] if len(sys.argv) > 3 else None
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
per = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[ix['Metric Name']] != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', r[ix['Kernel Name']])
    name = re.sub(r'void |<unnamed>::|at::native::', '', name)[:72]
    v = float(r[ix['Metric Value']].replace(',', ''))
    unit = r[ix['Metric Unit']]
    v = v / 1000.0 if unit == 'ns' else (v * 1000.0 if unit == 'ms' else v)
    per.append((name, v, r[ix['Grid Size']]))
n = len(per)
per = per[n - n // nsteps:]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for name, v, _ in per:
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
print('launches %d (of %d), total %.1f us' % (len(per), n, tot))
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print('%-74s n=%4d %9.1f us %5.1f%%' % (k, c, v, 100 * v / tot))
if pat:
    print()
    for name, v, g in per:
        if re.search(pat, name):
            print('%-50s %9.1f us  grid %s' % (name[:50], v, g))
