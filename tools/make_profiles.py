"""Turn the raw ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

    python tools/make_profiles.py <tag> <fwd_full.ncu-rep> <infer_launches.csv> <train_launches.csv>

* <tag>_fwd_full_summary.md : one row per tensor-core launch of ONE VNet forward (B = 20 patches of 96^3, fp16) from the
  `ncu --set full` capture: duration, DRAM bytes, tensor-pipe / L2 / DRAM utilisation.
* <tag>_traffic.json        : DRAM bytes per launch per kernel class (bench.py reads it for roofline.traffic).
* <tag>_infer_launches.csv / <tag>_train_launches.csv + *_summary.md : gpu__time_duration launch lists.
"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, 'profiles')

# tensor-core launches of one VNet forward in plan order (segmentation3d/_b200/plan.py)
VNET_ORDER = (['conv_tc_cin1', 'conv_tc_cin1', 'conv_tc_k2s2', 'conv_tc_k3', 'conv_tc_k2s2', 'conv_tc_k3', 'conv_tc_k3', 'conv_tc_k2s2'] +
              ['conv_tc_k3'] * 3 + ['conv_tc_k2s2'] + ['conv_tc_k3'] * 3 + ['conv_tc_t2s2'] + ['conv_tc_k3'] * 3 +
              ['conv_tc_t2s2'] + ['conv_tc_k3'] * 3 + ['conv_tc_t2s2'] + ['conv_tc_k3'] * 2 + ['conv_tc_t2s2', 'conv_tc_k3',
                                                                                               'conv_tc_narrow'])
VNET_NAMES = (['in_block.conv.stats', 'in_block.conv.gn_relu', 'down_32.down_conv', 'down_32.rblock.0', 'down_64.down_conv', 'down_64.rblock.0', 'down_64.rblock.1',
               'down_128.down_conv'] + ['down_128.rblock.%d' % i for i in range(3)] + ['down_256.down_conv'] +
              ['down_256.rblock.%d' % i for i in range(3)] + ['up_256.up_conv'] + ['up_256.rblock.%d' % i for i in range(3)] +
              ['up_128.up_conv'] + ['up_128.rblock.%d' % i for i in range(3)] + ['up_64.up_conv', 'up_64.rblock.0', 'up_64.rblock.1',
                                                                                'up_32.up_conv', 'up_32.rblock.0', 'out_block.conv1'])


def short(name):
    name = re.sub(r'\(.*', '', name)
    return re.sub(r'void |<unnamed>::|at::native::', '', name)[:60]


BATCH = int(os.environ.get('PROFILE_BATCH', '36'))


def full_summary(tag, rep):
    if rep.endswith('.csv'):      # `ncu -i <rep> --page raw --csv` already run on the GPU box (a 29-launch --set full report exceeds the transfer limit)
        raw = open(rep).read()
    else:
        raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    cols = [('gpu__time_duration.sum', 'us'), ('dram__bytes_read.sum', 'MB rd'), ('dram__bytes_write.sum', 'MB wr'),
            ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor %act'),
            ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 %'),
            ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM %'),
            ('l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'smem->TC %'),
            ('launch__grid_size', 'grid'), ('launch__registers_per_thread', 'regs')]

    def val(r, key):
        if key not in ix:
            return float('nan')
        v = float(r[ix[key]].replace(',', '') or 'nan')
        u = units[ix[key]]
        if key.startswith('dram__bytes'):
            v *= {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(u, 1.0)
        if key == 'gpu__time_duration.sum':
            v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0, 'nsecond': 1e-3, 'msecond': 1e3}.get(u, 1.0)
        return v

    data = rows[2:]
    lines = ['# %s: `ncu --set full` of the tensor-core launches of one VNet forward (B = %d patches of 96^3, fp16)' % (tag, BATCH), '',
             'Command: `ncu --set full --clock-control none --import-source on -k regex:"zmarch|persistent|fold|cin1" -s %d -c %d '
             'python tools/profile_forward.py %d fp16` (per-launch times are cold-cache and serialised).' % (len(VNET_ORDER), len(VNET_ORDER), BATCH), '',
             '| # | layer | class | kernel | ' + ' | '.join(c[1] for c in cols) + ' |', '|' + '---|' * (4 + len(cols))]
    traffic = collections.defaultdict(lambda: {'launches': 0, 'dram_bytes': 0.0, 'us': 0.0})
    for i, r in enumerate(data):
        kind = VNET_ORDER[i] if len(data) == len(VNET_ORDER) else '?'
        lname = VNET_NAMES[i] if len(data) == len(VNET_NAMES) else '?'
        vals = [val(r, c[0]) for c in cols]
        lines.append('| %d | %s | %s | %s | ' % (i, lname, kind, short(r[ix['Kernel Name']])) + ' | '.join('%.1f' % v for v in vals) + ' |')
        t = traffic[kind]
        t['launches'] += 1
        t['dram_bytes'] += (vals[1] + vals[2]) * 1e6
        t['us'] += vals[0]
    out = {}
    for k, t in traffic.items():
        out[k] = {'launches_per_forward': t['launches'], 'dram_bytes_per_launch': t['dram_bytes'] / t['launches'],
                  'avg_launch_us_under_ncu': t['us'] / t['launches']}
    out['_source'] = ('dram__bytes_read.sum + dram__bytes_write.sum from ncu --set full of one VNet forward, batch %d x 96^3 fp16 '
                      '(profiles/%s_fwd_full_summary.md)' % (BATCH, tag))
    out['_batch'] = BATCH
    out['_dram_bytes_per_patch_conv_launches'] = sum(t['dram_bytes'] for t in traffic.values()) / BATCH
    lines += ['', 'DRAM bytes per launch by class (written to `%s_traffic.json`):' % tag, '']
    for k, v in out.items():
        if not k.startswith('_'):
            lines.append('* %s: %.1f MB per launch over %d launches' % (k, v['dram_bytes_per_launch'] / 1e6, v['launches_per_forward']))
    open(os.path.join(PROF, tag + '_fwd_full_summary.md'), 'w').write('\n'.join(lines) + '\n')
    json.dump(out, open(os.path.join(PROF, tag + '_traffic.json'), 'w'), indent=1)


def launch_summary(tag, what, path, nsteps, note, marker=None):
    shutil.copy(path, os.path.join(PROF, '%s_%s_launches.csv' % (tag, what)))
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    per = []
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or r[ix['Metric Name']] != 'gpu__time_duration.sum':
            continue
        v = float(r[ix['Metric Value']].replace(',', ''))
        u = r[ix['Metric Unit']]
        per.append((short(r[ix['Kernel Name']]), v / 1000.0 if u == 'ns' else (v * 1000.0 if u == 'ms' else v)))
    n = len(per)
    marks = [i for i, (k, _) in enumerate(per) if k.startswith(marker)] if marker else []
    if len(marks) >= 2:      # one steady-state step = the launches between the last two occurrences of the marker kernel
        per = per[marks[-2]:marks[-1]]
    else:
        per = per[n - n // nsteps:]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, v in per:
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v for _, v in per)
    lines = ['# %s: %s launch list (gpu__time_duration.sum, --clock-control none)' % (tag, what), '', note, '',
             '%d launches, %.2f ms of kernel time (cold-cache, serialised: use the SHARES, not the absolute times).' % (len(per), tot / 1e3), '',
             '| kernel | launches | us | share |', '|---|---|---|---|']
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v / tot >= 0.002:
            lines.append('| %s | %d | %.1f | %.1f %% |' % (k, c, v, 100 * v / tot))
    open(os.path.join(PROF, '%s_%s_launches_summary.md' % (tag, what)), 'w').write('\n'.join(lines) + '\n')


if __name__ == '__main__':
    tag, rep, infer_csv, train_csv = sys.argv[1:5]
    os.makedirs(PROF, exist_ok=True)
    full_summary(tag, rep)
    launch_summary(tag, 'infer', infer_csv, 1, 'Command: `ncu --metrics gpu__time_duration.sum --clock-control none -s 1900 -c 660 --csv python bench.py '
                   '--steps 1 --warmup 3 --no-cpu-baseline` (one 512x512x400 volume = 5 forwards of 36 patches + gather / blend / finalize).')
    launch_summary(tag, 'train', train_csv, 3, 'Command: `ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/train_one_step.py '
                   'bf16 8 4`; the launches between the last two dice_terms_kernel launches = one steady-state training step '
                   '(backward + Adam of step 3, forward of step 4; B = 8 x 96^3, bf16, Dice, Adam).', marker='dice_terms_kernel')
