"""profiles/sass_summary.txt: which Blackwell-native instructions the shipped library contains, per kernel
(`cuobjdump -sass` of xnrs_b200/csrc/libxnrs_b200.so; SASS mnemonics per B200_PROFILING.md: tcgen05.mma -> UTC*MMA,
tcgen05.ld -> LDTM, TMA -> UTMALDG / UBLKCP, mma.sync -> HMMA).  Runs without a GPU.
    python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xnrs_b200 import _lib  # noqa: E402

MNEMONICS = ['UTCHMMA', 'UTCHMMA.2CTA', 'UTCQMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMALDG.2D.GATHER4', 'UTMASTG', 'UBLKCP', 'UTCBAR',
             'SYNCS', 'LDGSTS', 'HMMA', 'FFMA2', 'FFMA', 'REDG', 'RED.E', 'ATOMG']


def main():
    lib = _lib.build()
    sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = re.sub(r'\(.*', '', cur)
            per[cur] = collections.Counter()
            continue
        m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m and cur:
            op = m.group(1)
            for mn in MNEMONICS:
                if op == mn or op.startswith(mn + '.'):
                    per[cur][mn] += 1
    total = collections.Counter()
    for c in per.values():
        total.update(c)
    print(f'# {os.path.relpath(lib)}: SASS instruction counts (cuobjdump -sass), sm_100a only')
    print('# whole library: ' + ', '.join(f'{k} {v}' for k, v in sorted(total.items(), key=lambda kv: -kv[1])))
    print()
    for k, c in per.items():
        native = {m: n for m, n in c.items() if m.startswith(('UTC', 'LDTM', 'STTM', 'UTMA', 'UBLKCP', 'SYNCS', 'LDGSTS', 'HMMA'))}
        if native:
            print(f'{k}: ' + ', '.join(f'{m} {n}' for m, n in sorted(native.items())) + (f' | FFMA2 {c["FFMA2"]}' if c['FFMA2'] else ''))
    print()
    print('# SIMT kernels (no tensor-core / TMA instructions): ' + ', '.join(k for k, c in per.items()
          if not any(m.startswith(('UTC', 'LDTM', 'UTMA', 'HMMA')) for m in c)))


if __name__ == '__main__':
    main()
