"""SASS evidence: counts of the Blackwell-native instructions per kernel of libdfol_b200.so (cuobjdump -sass), as a
markdown table -> profiles/r2_sass_summary.md.  Mnemonics: UTCHMMA = tcgen05.mma (kind::f16), LDTM = tcgen05.ld,
UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = cp.async.bulk, UTMAPF / UBLKPF = bulk L2 prefetch,
UTCBAR = tcgen05.commit, FFMA2 = packed fp32 pair FMA."""
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, 'dfol_vqa_b200', 'libdfol_b200.so')
COLS = [('UTCHMMA', 'UTCHMMA (tcgen05.mma)'), ('LDTM', 'LDTM (tcgen05.ld)'), ('UTMALDG', 'UTMALDG (TMA load)'),
        ('UTMASTG', 'UTMASTG (TMA store)'), ('UBLKCP', 'UBLKCP (bulk copy)'), ('UTMAPF|UBLKPF', 'L2 prefetch'),
        ('UTCBAR', 'UTCBAR (tcgen05.commit)'), ('FFMA2', 'FFMA2')]


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    kernels, name = {}, None
    for line in sass.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            name = m.group(1)
            kernels[name] = {c: 0 for c, _ in COLS}
            continue
        if name is None:
            continue
        for c, _ in COLS:
            if re.search(r'\b(%s)\b' % c, line):
                kernels[name][c] += 1
    demangled = subprocess.run(['c++filt'], input='\n'.join(kernels), capture_output=True, text=True).stdout.splitlines()
    out = ['# SASS evidence (cuobjdump -sass dfol_vqa_b200/libdfol_b200.so, sm_100a), round 2 final state', '',
           'Instruction counts per kernel (only kernels that use tcgen05 / TMA / bulk-async); `tools/sass_summary.py`.', '',
           '| kernel | ' + ' | '.join(t for _, t in COLS) + ' |', '|---|' + '---|' * len(COLS)]
    for mangled, nice in sorted(zip(kernels, demangled), key=lambda x: x[1]):
        k = kernels[mangled]
        if not any(k[c] for c, _ in COLS[:7]):
            continue
        short = re.sub(r'\(.*', '', nice)
        out.append('| `%s` | ' % short + ' | '.join(str(k[c]) for c, _ in COLS) + ' |')
    text = '\n'.join(out) + '\n'
    open(os.path.join(REPO, 'profiles', 'r2_sass_summary.md'), 'w').write(text)
    sys.stdout.write(text)


if __name__ == '__main__':
    main()
