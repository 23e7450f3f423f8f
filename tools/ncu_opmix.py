"""Dynamic SASS opcode mix of the kernel in an ncu report (--set full --import-source on).

    python tools/ncu_opmix.py gpurun_out/prof.ncu-rep [rows] [kernel-name-substring]     # rows: units of work to normalise by

Reads `ncu --page source --print-source sass --csv` and sums "Instructions Executed" (warp level) and the stall samples per opcode."""
import collections
import csv
import io
import re
import subprocess
import sys


def main():
    rep = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'sass', '--csv'], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    pat = sys.argv[3] if len(sys.argv) > 3 else ''
    starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']         # a report may hold several kernels
    pick = next((i for i in starts if pat in rows[i][1]), starts[0])
    rows = rows[pick:next((i for i in starts if i > pick), len(rows))]
    hdr = rows[1]
    ci, si, ni = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('# Samples')
    ops, samp = collections.Counter(), collections.Counter()
    for r in rows[2:]:
        if len(r) <= ci:
            continue
        m = re.match(r'\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)', r[si])
        if not m:
            continue
        op = m.group(1)
        if op in ('LDS', 'STS', 'LDG', 'STG') and m.group(2):
            w = re.search(r'\.(64|128)', m.group(2))
            op += '.' + (w.group(1) if w else '32')
        ops[op] += int(r[ci])
        samp[op] += int(r[ni])
    tot, ts = sum(ops.values()), sum(samp.values())
    print('total warp instructions %d (%.1f per unit), samples %d' % (tot, tot / units, ts))
    print('| opcode | warp instr per unit | share | stall-sample share |\n|---|---|---|---|')
    for op, n in ops.most_common(40):
        print('| %s | %.1f | %.1f %% | %.1f %% |' % (op, n / units, 100.0 * n / tot, 100.0 * samp[op] / max(ts, 1)))


if __name__ == '__main__':
    main()
