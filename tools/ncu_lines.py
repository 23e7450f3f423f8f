"""Where a kernel's instructions and stall samples go, by CUDA source line.

    python tools/ncu_lines.py <report.ncu-rep> <object.o built from the same sources> <kernel-name-substring> [top]

Joins the per-SASS-instruction counters of an ncu report (--set full --import-source on) with the line table of the cubin
(`nvdisasm -g`), by instruction offset inside the kernel.  Needs the object file of the build that was profiled."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, obj, pat = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith('.cubin')][0]
    dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
    # line table of the first function whose mangled name contains the pattern
    table, cur, active = {}, None, False
    for ln in dis.splitlines():
        m = re.match(r'\.text\.(\S+):', ln)
        if m:
            if active:
                break
            active = pat in m.group(1)
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);', ln)
        if m:
            table[int(m.group(1), 16)] = cur
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'sass', '--csv'], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    # a report may hold several kernels: take the first section whose kernel name matches (demangled name of `pat`'s stem)
    stem = re.sub(r'I[A-Za-z0-9_]*$', '', pat)
    starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
    pick = next((i for i in starts if stem in rows[i][1].replace(' ', '')), starts[0])
    end = next((i for i in starts if i > pick), len(rows))
    rows = rows[pick:end]
    hdr = rows[1]
    ai, ci, ni = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    base = None
    inst, samp = collections.Counter(), collections.Counter()
    for r in rows[2:]:
        if len(r) <= ci:
            continue
        addr = int(r[ai], 16)
        if base is None:
            base = addr
        key = table.get(addr - base, ('?', 0))
        inst[key] += int(r[ci])
        samp[key] += int(r[ni])
    ti, ts = sum(inst.values()), sum(samp.values())
    print('kernel *%s*: %d warp instructions, %d stall samples' % (pat, ti, ts))
    print('| file:line | instructions | stall samples |\n|---|---|---|')
    for key, n in sorted(inst.items(), key=lambda kv: -samp[kv[0]])[:top]:
        print('| %s:%d | %.1f %% | %.1f %% |' % (key[0], key[1], 100.0 * n / ti, 100.0 * samp[key] / max(ts, 1)))


if __name__ == '__main__':
    main()
