"""Per-layer roofline table from a `bench.py --layers` listing (no GPU needed):

    python tools/roofline_table.py profiles/r01_bench_layers_b20.txt > profiles/r01_roofline_per_layer.md

Every launch is put against the roofline that bounds it: the measured sustained bf16 tensor peak when its arithmetic
intensity (algorithmic flop / algorithmic byte) is above the ridge, the measured HBM copy bandwidth otherwise
(MEASURED_PEAKS.json).  Times are CUDA-event pairs around single launches of a batch of 20 patches of 96^3 (fp16)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(path):
    import re
    m = re.search(r'_b(\d+)', os.path.basename(path))
    batch = int(m.group(1)) if m else 20
    pk = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    hbm, tf = pk['hbm_gbs'], pk['bf16_tflops_sustained']
    ridge = tf * 1e3 / hbm
    rows = []
    for line in open(path):
        f = line.split()
        if len(f) < 8 or f[0] == 'KIND' or f[3] != 'ms':
            continue
        rows.append((f[0], f[1], float(f[2]), float(f[4]), float(f[6])))
    total = sum(r[2] for r in rows)
    print('# every launch of one VNet forward against its roofline (batch of 96^3 patches as in the source listing, fp16)\n')
    print('Source: `%s` (`python bench.py --layers`, CUDA events per launch).  Peaks (MEASURED_PEAKS.json): HBM %.0f GB/s, bf16 tensor %.1f '
          'TFLOP/s sustained; ridge %.0f flop/B.  A launch is tensor-bound when its algorithmic flop/byte exceeds the ridge.  Launches under '
          '~20 us are dominated by the event pair around them (their GB/s is a lower bound).\n' % (os.path.relpath(path, ROOT), hbm, tf, ridge))
    print('| layer | kernel class | ms | share | TFLOP/s | GB/s | bound | fraction of roofline |')
    print('|---|---|---|---|---|---|---|---|')
    agg = {}
    for name, kind, ms, tfl, gbs in rows:
        ai = (tfl * 1e3 / gbs) if gbs > 0 else 0.0
        bound = 'tensor' if ai > ridge else 'hbm'
        frac = tfl / tf if bound == 'tensor' else gbs / hbm
        note = '' if ms >= 0.02 else ' (launch-latency floor)'
        print('| %s | %s | %.3f | %.1f %% | %.1f | %.0f | %s | %.2f%s |' % (name, kind, ms, 100 * ms / total, tfl, gbs, bound, frac, note))
        a = agg.setdefault(bound, [0.0, 0.0])
        a[0] += ms
        a[1] += ms * frac
    print('\nTime-weighted: ' + '; '.join('%s-bound launches %.2f ms (%.0f %% of the forward) at %.2f of their roofline' % (
        b, a[0], 100 * a[0] / total, a[1] / a[0]) for b, a in sorted(agg.items())) + '.  Whole forward %.2f ms per batch of %d = %.3f ms per patch.' % (total, batch, total / batch))
    big = [(n, ms) for n, k, ms, t, g in rows if ms >= 0.02]
    below = []
    for name, kind, ms, tfl, gbs in rows:
        if ms < 0.02:
            continue
        ai = (tfl * 1e3 / gbs) if gbs > 0 else 0.0
        frac = tfl / tf if ai > ridge else gbs / hbm
        if frac < 0.5:
            below.append('%s (%.2f)' % (name, frac))
    print('\nLaunches of >= 20 us below the 50 %% target: %s.' % (', '.join(below) if below else 'none'))


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'profiles', 'r01_bench_layers_b20.txt'))
