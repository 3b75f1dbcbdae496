"""Build an experimental copy of the library with extra -D switches on ONE translation unit:

    python profiles/tools/build_variant.py <name> <source.cu> [-DFOO=1 ...]

writes profiles/tools/variants/lib_<name>.so (git-ignored, shipped to the GPU box by gpurun) from the stock
objects of favae_b200/build plus the re-compiled unit; select it with FAVAE_B200_LIB=<path>.  Prints the
ptxas resource lines of the kernels whose name contains $FAVAE_VARIANT_GREP (default: all with spills)."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from favae_b200 import _build  # noqa: E402


def main():
    name, unit, defs = sys.argv[1], sys.argv[2], sys.argv[3:]
    _build.build()                                    # stock objects
    out_dir = os.path.join(ROOT, 'profiles', 'tools', 'variants')
    os.makedirs(out_dir, exist_ok=True)
    src = os.path.join(_build.CSRC, unit)
    obj = os.path.join(out_dir, f'{name}_{unit[:-3]}.o')
    cmd = [_build._nvcc(), *_build.NVCC_FLAGS, '-Xptxas', '-v', *defs, '-c', src, '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.exit(r.stdout + r.stderr)
    pat = os.environ.get('FAVAE_VARIANT_GREP', '')
    lines = r.stderr.splitlines()
    for i, ln in enumerate(lines):
        if 'Compiling entry function' in ln and (pat and pat in ln):
            print(ln.split("'")[1][:90])
            for extra in lines[i + 1:i + 4]:
                if 'registers' in extra or 'spill' in extra:
                    print('   ', extra.strip()[:160])
    objs = [os.path.join(_build.OBJ, os.path.basename(s)[:-3] + '.o') for s in _build.sources() if os.path.basename(s) != unit]
    lib = os.path.join(out_dir, f'lib_{name}.so')
    cmd = [_build._nvcc(), '-shared', '-o', lib, obj, *objs, '-cudart', 'static', '-Xlinker', '--no-undefined',
           '-lpthread', '-ldl', '-lrt']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.exit(r.stdout + r.stderr)
    print(lib)


if __name__ == '__main__':
    main()
