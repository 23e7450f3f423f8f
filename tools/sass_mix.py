"""SASS instruction mix of every fp32 kernel in the built objects (runs without a GPU).

    python tools/sass_mix.py > profiles/<name>.md

Counts static instructions per kernel from `cuobjdump -sass build/*.o`: FFMA2 / FMUL2 / FADD2 are the packed f32x2 forms
of sm_100a, LDC / LDCU the constant-bank (uniform) operands, LDG the global loads (composite rows, coefficient tables)."""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ['total', 'FFMA', 'FFMA2', 'FMUL', 'FADD', 'FMUL2', 'FADD2', 'SHFL', 'LDS', 'STS', 'LDG', 'STG', 'LDC', 'LDCU',
        'BAR', 'MUFU']


def main():
    out = collections.OrderedDict()
    objs = sorted(glob.glob(os.path.join(ROOT, 'build', '*_f32.o'))) + [os.path.join(ROOT, 'build', 'cm_qam_p3.o')]
    for obj in objs:
        txt = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
        cur = None
        for line in txt.splitlines():
            m = re.search(r'Function : (\S+)', line)
            if m:
                name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
                cur = name.split('(')[0].replace('void ', '')
                out[cur] = collections.Counter()
                continue
            m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
            if m and cur:
                out[cur][m.group(1)] += 1
                out[cur]['total'] += 1
    print('# Static SASS instruction mix per kernel (sm_100a, `cuobjdump -sass build/*.o`, `tools/sass_mix.py`)\n')
    print('| kernel | ' + ' | '.join(KEYS) + ' |')
    print('|---|' + '---|' * len(KEYS))
    for k, c in out.items():
        if '<float' in k:
            print('| `%s` | ' % k + ' | '.join(str(c[x]) for x in KEYS) + ' |')


if __name__ == '__main__':
    main()
