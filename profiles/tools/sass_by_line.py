"""Static SASS instruction counts of one kernel attributed to the source lines of one file, following the
inlining chain (`nvdisasm -gi`): every instruction goes to the LAST line of <file> on its chain, i.e. to the
statement of the phase driver it was inlined into.

    python profiles/tools/sass_by_line.py favae_b200/build/ffl_kernels.o <kernel-name-substring> ffl_driver.cuh [first_line]
"""
import collections
import os
import re
import subprocess
import sys
import tempfile


def main():
    obj, kernel, fname = sys.argv[1:4]
    first = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
    text = subprocess.run(['nvdisasm', '-gi', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    fp = {'FADD2', 'FFMA2', 'FMUL2', 'FFMA', 'FMUL', 'FADD', 'MUFU', 'FMNMX', 'FMNMX3'}
    mem = {'LDS', 'STS', 'LDG', 'STG', 'ST', 'LD', 'LDGSTS', 'STAS', 'UBLKCP'}
    inside, pending, cur = False, [], []
    tot, non, ops = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
    for ln in text.splitlines():
        if ln.startswith('//---') and '.text.' in ln:
            inside = kernel in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
        if m:
            pending.append((os.path.basename(m.group(1)), int(m.group(2)),
                            os.path.basename(m.group(3)) if m.group(3) else None, int(m.group(4)) if m.group(4) else 0))
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', ln)
        if not m:
            continue
        if pending:
            cur, pending = pending, []
        lines = [l for f, l, f2, l2 in cur if f == fname] + [l2 for f, l, f2, l2 in cur if f2 == fname]
        lines = [l for l in lines if l >= first]
        key = max(lines) if lines else 0
        op = m.group(1)
        tot[key] += 1
        if op not in fp and op not in mem:
            non[key] += 1
            ops[key][op] += 1
    print(f'# {kernel}: {sum(tot.values())} instructions, {sum(non.values())} neither floating point nor memory')
    print(f'# {"line":>5} {"total":>6} {"other":>6}  most frequent other opcodes')
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:40]:
        print(f'  {k:5d} {v:6d} {non[k]:6d}  ' + ' '.join(f'{o}:{c}' for o, c in ops[k].most_common(5)))


if __name__ == '__main__':
    main()
