"""Per-kernel SASS evidence for the tensor-core / TMA path (B200_PROFILING.md: UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor
load, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, SYNCS = mbarrier) plus registers / shared memory from the cubin.

    python tools/sass_summary.py > profiles/r01_sass_summary.md        (no GPU needed: reads the built .so)
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'medical-segmentation3d-toolkit_b200', 'lib', 'libseg3d_b200.so')
KEYS = ('UTCHMMA', 'UTMALDG', 'UTMASTG', 'LDTM', 'UTCBAR', 'UTCATOMSWS', 'SYNCS', 'HMMA', 'FFMA', 'REDG', 'ATOMG', 'ATOMS', 'LDG', 'STG', 'LDS', 'STS')


def demangle(names):
    out = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True).stdout.split('\n')
    return dict(zip(names, out))


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in sass.split('\n'):
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r'\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
        if m:
            op = m.group(1)
            counts[cur]['_total'] += 1
            for k in KEYS:
                if op == k or op.startswith(k + '.') or (k in ('UTCHMMA', 'UTMALDG', 'UTMASTG', 'LDTM', 'UTCBAR', 'UTCATOMSWS', 'SYNCS') and op.startswith(k)):
                    counts[cur][k] += 1
                    break
    res = subprocess.run(['cuobjdump', '-res-usage', LIB], capture_output=True, text=True).stdout
    usage, cur = {}, None
    for line in res.split('\n'):
        m = re.match(r'\s*Function (\S+):', line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r'REG:(\d+).*?SHARED:(\d+)', line)
        if m and cur:
            usage[cur] = (int(m.group(1)), int(m.group(2)))
    names = demangle(list(counts))
    print('# r01: SASS summary of libseg3d_b200.so (sm_100a) - `python tools/sass_summary.py`\n')
    print('`UTCHMMA` = tcgen05.mma, `UTMALDG` = TMA tensor load (cp.async.bulk.tensor), `LDTM` = tcgen05.ld (TMEM -> registers), '
          '`UTCBAR` = tcgen05.commit, `UTCATOMSWS` = tcgen05.alloc/dealloc, `SYNCS` = mbarrier ops.  Counts are static '
          'instructions in the kernel body; REG / static SHARED from `cuobjdump -res-usage` (dynamic shared memory is set at launch).\n')
    print('| kernel | instr | UTCHMMA | UTMALDG | LDTM | UTCBAR | SYNCS | FFMA | REDG/ATOMG | REG | static smem |')
    print('|---|---|---|---|---|---|---|---|---|---|---|')
    rows = []
    for mangled, c in counts.items():
        nm = names.get(mangled, mangled)
        nm = re.sub(r'\(anonymous namespace\)::', '', nm)
        nm = re.sub(r'^void ', '', nm)
        nm = re.sub(r'\(.*$', '', nm)
        reg, sh = usage.get(mangled, ('', ''))
        rows.append((0 if c['UTCHMMA'] else 1, nm, c, reg, sh))
    for _, nm, c, reg, sh in sorted(rows, key=lambda r: (r[0], r[1])):
        print('| `%s` | %d | %d | %d | %d | %d | %d | %d | %d | %s | %s |' % (
            nm, c['_total'], c['UTCHMMA'], c['UTMALDG'], c['LDTM'], c['UTCBAR'], c['SYNCS'], c['FFMA'], c['REDG'] + c['ATOMG'], reg, sh))
    tc = [r for r in rows if r[0] == 0]
    print('\n%d kernels, %d of them issue tcgen05.mma; no `HMMA` (mma.sync) instruction anywhere: %s.' % (
        len(rows), len(tc), 'true' if not any(r[2]['HMMA'] for r in rows) else 'FALSE'))


if __name__ == '__main__':
    main()
