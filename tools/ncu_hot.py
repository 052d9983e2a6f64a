"""Top SASS instructions by warp-stall samples from `ncu -i X.ncu-rep --page source --csv --kernel-name regex:K`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
data = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        n = int(r[ix['# Samples']] or 0)
    except ValueError:
        continue
    data.append((n, r))
tot = sum(n for n, _ in data)
print('total samples', tot)
for pos, (n, r) in enumerate(data):
    r.append(pos)
for n, r in sorted(data, key=lambda t: -t[0])[:top]:
    st = sorted(((int(r[ix[s]] or 0), s) for s in stalls), reverse=True)[:3]
    print('%5.1f%% #%4d  %-60s exec=%-8s %s' % (100.0 * n / tot, r[-1], r[ix['Source']].strip()[:60], r[ix['Instructions Executed']],
                                          ' '.join('%s:%d' % (s[6:], c) for c, s in st if c)))
