"""Per-source-line hot spots of one kernel from an ncu report captured with --import-source on:

    python tools/ncu_hotlines.py gpurun_out/prof_x.ncu-rep [top_n] [kernel_index]

Aggregates the cuda,sass correlated source page by (file, line): stall samples, instructions
executed, shared-memory wavefronts, dominant stall reasons.
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source',
                          'cuda,sass'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    cur_file = None
    hdr = None
    agg = defaultdict(lambda: defaultdict(float))
    src = {}
    kernels_seen = 0
    want = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    fn_names = []
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            cur_file = r[1].split('/')[-1]
            continue
        if r[0] == 'Function Name':
            if r[1] not in fn_names:
                fn_names.append(r[1])
            continue
        if r[0] == 'Line No':
            hdr = r
            continue
        if hdr is None:
            continue
        if r[0]:
            cur_line = (cur_file, int(r[0]))
            src[cur_line] = r[1].strip()
            continue
        if len(r) < len(hdr) or r[2] == '...':
            continue
        d = dict(zip(hdr[2:], r[2:]))
        a = agg[cur_line]

        def num(k):
            try:
                return float(d.get(k, '0').replace(',', ''))
            except ValueError:
                return 0.0
        a['samples'] += num('# Samples')
        a['inst'] += num('Instructions Executed')
        a['smem_wf'] += num('L1 Wavefronts Shared')
        a['smem_ideal'] += num('L1 Wavefronts Shared Ideal')
        for k in hdr:
            if k.startswith('stall_') and 'Not Issued' not in k:
                a[k] += num(k)
    tot_s = sum(a['samples'] for a in agg.values()) or 1
    tot_i = sum(a['inst'] for a in agg.values()) or 1
    print(f'# {rep}: {int(tot_s)} samples, {int(tot_i)} warp instructions (all captured launches)')
    print(f'{"file:line":28s} {"samp%":>6s} {"inst%":>6s} {"smem_wf":>10s} {"xs":>5s}  stalls | source')
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1]['samples'])[:top]:
        st = sorted(((k[6:], v) for k, v in a.items() if k.startswith('stall_')), key=lambda kv: -kv[1])[:3]
        sts = ' '.join(f'{k}:{100 * v / max(a["samples"], 1):.0f}' for k, v in st)
        xs = a['smem_wf'] / a['smem_ideal'] if a['smem_ideal'] else 0
        print(f'{key[0][:20]}:{key[1]:<6d} {100 * a["samples"] / tot_s:6.2f} {100 * a["inst"] / tot_i:6.2f} '
              f'{int(a["smem_wf"]):10d} {xs:5.2f}  {sts} | {src.get(key, "")[:90]}')


if __name__ == '__main__':
    main()
