"""Build libfavae_b200.so (the C-ABI library, include/favae_b200.h) in-tree with nvcc for sm_100a.

No torch headers are involved: the library links only the static CUDA runtime and obtains the
driver's tensor-map encoder through cudaGetDriverEntryPoint, so it loads (and exports its
symbols) on a machine without a GPU driver too.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, 'build')
LIB = os.environ.get('FAVAE_B200_LIB') or os.path.join(HERE, 'libfavae_b200.so')   # override: experiments only

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-lineinfo',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden', '--expt-relaxed-constexpr']


def _nvcc():
    return shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _source_digest():
    """sha256 over every source the library is built from (paths + contents)."""
    import hashlib
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, '..', 'include')):
        for f in sorted(os.listdir(root)):
            h.update(f.encode())
            with open(os.path.join(root, f), 'rb') as fh:
                h.update(fh.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


STAMP = os.path.join(HERE, 'libfavae_b200.stamp')


def needs_build():
    """True unless the library exists and was built from exactly the current sources.  A content
    digest (not mtimes) decides, so a copied tree (gpurun snapshot) never triggers a rebuild."""
    if os.environ.get('FAVAE_B200_LIB'):
        return False                       # an explicitly named library is used as it is
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    try:
        return open(STAMP).read().strip() != _source_digest()
    except OSError:
        return True


def _compile(src):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + '.o')
    cmd = [_nvcc(), *NVCC_FLAGS, '-c', src, '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'nvcc failed for {src}:\n{r.stdout}\n{r.stderr}')
    return obj


def build(force=False, verbose=True):
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(_compile, sources()))
    tmp = f'{LIB}.{os.getpid()}.tmp'           # link under a private name, then rename atomically:
    cmd = [_nvcc(), '-shared', '-o', tmp, *objs, '-cudart', 'static', '-Xlinker', '--no-undefined',
           '-lpthread', '-ldl', '-lrt']          # nobody can dlopen a half-written library
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    os.replace(tmp, LIB)
    with open(STAMP, 'w') as fh:
        fh.write(_source_digest())
    if verbose:
        print(f'[favae_b200] built {LIB}', file=sys.stderr)
    return LIB


if __name__ == '__main__':
    build(force='--force' in sys.argv)
