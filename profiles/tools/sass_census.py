"""SASS instruction census of libfavae_b200.so: `cuobjdump -sass` per kernel, counting the mnemonics
that prove the Blackwell-native paths (tcgen05 = UTCHMMA / UTCBAR / LDTM, TMA = UTMALDG, packed fp32x2 =
FFMA2 / FADD2 / FMUL2, cp.async = LDGSTS, bulk L2 prefetch = UBLKPF, cluster barriers = UCGABAR).
Run anywhere nvcc's tools are installed (no GPU needed):  python profiles/tools/sass_census.py > profiles/sass_census_r2.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, 'favae_b200', 'libfavae_b200.so')
KEY = ['UTCHMMA', 'UTCBAR', 'LDTM', 'UTMALDG', 'UTMAPF', 'SYNCS', 'FFMA2', 'FADD2', 'FMUL2', 'FFMA', 'FADD', 'FMUL',
       'LDGSTS', 'UBLKPF', 'UCGABAR', 'LDS', 'STS', 'LDG', 'STG', 'ATOMG', 'RED', 'MUFU', 'SHFL', 'BAR', 'LDL', 'STL']


def shorten(n):
    n = re.sub(r'\((int|bool)\)', '', n).replace('void ', '').replace('favae::', '')
    depth, cut = 0, len(n)
    for i, ch in enumerate(n):                      # drop the parameter list: first '(' outside <...>
        if ch == '<':
            depth += 1
        elif ch == '>':
            depth -= 1
        elif ch == '(' and depth == 0:
            cut = i
            break
    return n[:cut][:78]


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    kern, counts, order = None, {}, []
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            kern = m.group(1)
            counts[kern] = collections.Counter()
            order.append(kern)
            continue
        m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)', line)
        if m and kern:
            counts[kern][m.group(1)] += 1
            full = m.group(1) + m.group(2)
            if m.group(1) in ('UTCHMMA', 'UTMALDG', 'UTCBAR', 'LDTM'):
                counts[kern]['~' + full] += 1
    names = subprocess.run(['cu++filt'] + order, capture_output=True, text=True).stdout.splitlines() \
        if order else []
    names = names if len(names) == len(order) else order
    print(f'# {os.path.relpath(LIB, ROOT)}: {len(order)} kernels; counts are static SASS instructions')
    print(f'{"kernel":<78} {"total":>6} ' + ' '.join(f'{k:>7}' for k in KEY))
    for k, n in sorted(zip(order, names), key=lambda kn: kn[1]):
        c = counts[k]
        short = shorten(n)
        print(f'{short:<78} {sum(v for kk, v in c.items() if not kk.startswith("~")):6d} ' + ' '.join(f'{c.get(kk, 0):7d}' for kk in KEY))
    print('\n# tcgen05 / TMA variants seen (full mnemonics):')
    for k, n in sorted(zip(order, names), key=lambda kn: kn[1]):
        v = {kk[1:]: vv for kk, vv in counts[k].items() if kk.startswith('~')}
        if v:
            print(shorten(n), v)


if __name__ == '__main__':
    main()
