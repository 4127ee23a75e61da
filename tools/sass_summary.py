"""SASS mnemonic summary of the product library (cuobjdump -sass): per kernel the counts of the
instructions that identify the hardware paths - tcgen05 (UTCHMMA / UTCBAR / LDTM), TMA tensor and
bulk copies (UTMALDG / UBLKCP), legacy tensor pipe (HMMA / LDSM), local-memory spills (STL / LDL),
atomics (ATOM / RED / ATOMS) - written as CSV.  Usage: python tools/sass_summary.py [out.csv]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'htd_b200', '_lib', 'libhtd_b200.so')
KEYS = ['UTCHMMA', 'UTCBAR', 'LDTM', 'UTMALDG', 'UBLKCP', 'SYNCS', 'HMMA', 'LDSM', 'STL', 'LDL',
        'ATOM', 'ATOMS', 'RED', 'DFMA', 'FFMA']


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'profiles', 'r02_sass_summary.csv')
    txt = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    rows, name, cnt, total = [], None, None, 0
    ins = re.compile(r'^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)')
    for line in txt.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            if name:
                rows.append((name, total, cnt))
            name, cnt, total = m.group(1), collections.Counter(), 0
            continue
        m = ins.match(line)
        if m and name:
            cnt[m.group(1)] += 1
            total += 1
    if name:
        rows.append((name, total, cnt))
    demangled = subprocess.run(['c++filt'], input='\n'.join(r[0] for r in rows), capture_output=True,
                               text=True).stdout.splitlines()
    with open(out, 'w') as f:
        f.write('kernel,instructions,' + ','.join(KEYS) + '\n')
        for (n, total, c), d in sorted(zip(rows, demangled), key=lambda x: x[1]):
            short = re.sub(r'\(.*', '', d).replace('void ', '')
            f.write('"%s",%d,%s\n' % (short, total, ','.join(str(c.get(k, 0)) for k in KEYS)))
    print(out, len(rows), 'kernels')


if __name__ == '__main__':
    main()
