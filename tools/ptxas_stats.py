#!/usr/bin/env python
"""Register / spill table of one CUDA source: tools/ptxas_stats.py <file.cu> [filter] [-- extra nvcc flags]"""
import re, subprocess, sys, os
src = os.path.abspath(sys.argv[1]); flt = sys.argv[2] if len(sys.argv) > 2 and sys.argv[2] != '--' else ''
extra = sys.argv[sys.argv.index('--') + 1:] if '--' in sys.argv else []
out = subprocess.run(['nvcc', '-O3', '-std=c++17', '-lineinfo', '-gencode', 'arch=compute_100a,code=sm_100a',
                      '--expt-relaxed-constexpr', '-Xptxas', '-v', '-c', src, '-o', '/tmp/_ptxas_stats.o'] + extra,
                     capture_output=True, text=True, cwd=os.path.dirname(os.path.abspath(src))).stderr
names = subprocess.run(['c++filt'], input=out, capture_output=True, text=True).stdout
cur = None
for line in names.splitlines():
    m = re.search(r"Compiling entry function '(.*)' for", line)
    if m: cur = re.sub(r'yunet::\(anonymous namespace\)::|void |yunet::', '', m.group(1)); cur = re.sub(r'\(.*\)$', '', cur); continue
    m = re.search(r'(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads', line)
    if m: sp = m.groups(); continue
    m = re.search(r'Used (\d+) registers', line)
    if m and cur and flt in cur:
        print('%-60s regs %3s  stack %4s  spill st %4s ld %4s' % (cur, m.group(1), *sp))
