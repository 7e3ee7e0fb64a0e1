"""Opcode evidence of the Blackwell-native path: per kernel of libyunet_b200.so the counts of the
SASS mnemonics that prove tcgen05 / TMEM / TMA (UTCHMMA, LDTM, STTM, UTMALDG, UTMAPF, UBLKCP,
SYNCS = mbarrier) next to the generic ones (FFMA2, LDS, STS, LDG, STG, LD/ST generic, LDL/STL spills).

    python tools/sass_counts.py > profiles/r2_sass_counts.txt
"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'libfacedetection', 'train_b200', 'libyunet_b200.so')
KEYS = ['UTCHMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMAPF', 'UBLKCP', 'SYNCS', 'UTCBAR', 'FFMA2', 'FADD2', 'FFMA',
        'LDS', 'STS', 'LDG', 'STG', 'LD', 'ST', 'LDL', 'STL']


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r'arch = (sm_\w+)', out)))
    kernels = OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r'yunet::\(anonymous namespace\)::|\(anonymous namespace\)::|yunet::', '', cur)
            cur = re.sub(r'\(.*', '', cur).replace('void ', '')
            kernels[cur] = Counter()
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)', line)
        if m and cur is not None:
            kernels[cur][m.group(1)] += 1
    print(f'# {os.path.relpath(LIB, ROOT)}: architectures {arch}; {len(kernels)} kernels')
    print(f'{"kernel":58s} ' + ' '.join(f'{k:>7s}' for k in KEYS))
    tot = Counter()
    for name, c in kernels.items():
        tot.update(c)
        if sum(c[k] for k in KEYS[:8]) == 0 and '--all' not in sys.argv:
            continue
        print(f'{name[:58]:58s} ' + ' '.join(f'{c[k]:7d}' for k in KEYS))
    print(f'{"TOTAL (all kernels)":58s} ' + ' '.join(f'{tot[k]:7d}' for k in KEYS))


if __name__ == '__main__':
    main()
