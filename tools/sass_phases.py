#!/usr/bin/env python
"""Per-phase breakdown of a kernel from an ncu report: the SASS of a kernel is cut at its BAR.SYNC instructions and the
executed warp instructions / warp-stall samples of every segment are summed.

    ncu -i prof.ncu-rep --page source --csv --kernel-name regex:<kernel> > k.csv
    python tools/sass_phases.py k.csv [launch index]
"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
    block = rows[starts[which]:starts[which + 1]]
    print(block[0][1])
    hdr, data = block[1], [r for r in block[2:] if len(r) > 10]
    ia, isamp, isrc = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Source')
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    tot_i = sum(int(r[ia]) for r in data)
    tot_s = sum(int(r[isamp]) for r in data)
    print('warp instructions %d, samples %d, SASS lines %d' % (tot_i, tot_s, len(data)))
    segs, cur = [], None

    def fresh():
        return dict(i=0, s=0, n=0, ops={}, st={})

    cur = fresh()
    for r in data:
        toks = r[isrc].split()
        op = toks[1] if toks[0].startswith('@') else toks[0]
        base = op.split('.')[0]
        cur['i'] += int(r[ia])
        cur['s'] += int(r[isamp])
        cur['n'] += 1
        cur['ops'][base] = cur['ops'].get(base, 0) + int(r[ia])
        for i, h in stall_cols:
            cur['st'][h] = cur['st'].get(h, 0) + int(r[i] or 0)
        if base in ('BAR', 'EXIT'):
            segs.append(cur)
            cur = fresh()
    segs.append(cur)
    for k, s in enumerate(segs):
        if not s['n']:
            continue
        top = sorted(s['ops'].items(), key=lambda kv: -kv[1])[:9]
        st = sorted(s['st'].items(), key=lambda kv: -kv[1])[:4]
        print('%2d static %5d  inst %5.1f%%  samples %5.1f%%  | %s | %s' % (
            k, s['n'], 100.0 * s['i'] / tot_i, 100.0 * s['s'] / max(tot_s, 1),
            ' '.join('%s:%.1f%%' % (a, 100.0 * b / tot_i) for a, b in top),
            ' '.join('%s:%.1f%%' % (a[6:], 100.0 * b / max(tot_s, 1)) for a, b in st)))


if __name__ == '__main__':
    main()
