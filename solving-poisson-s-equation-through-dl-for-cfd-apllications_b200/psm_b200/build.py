"""Build ``libpsm_b200.so`` in-tree with nvcc for sm_100a (``python -m psm_b200.build``).

The built library lives next to this file (git-ignored, but it travels with the tree to the
GPU box).  nvcc cross-compiles without a GPU.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), 'csrc')
OUT = os.path.join(HERE, 'libpsm_b200.so')
SOURCES = ['psm_plan.cpp', 'psm_files.cpp', 'psm_partition.cpp', 'psm_kernels.cu', 'psm_gemm_tc.cu', 'psm_integrate.cu', 'psm_init.cu', 'psm_handle.cu']
HEADERS = ['psm_plan.h', 'psm_kernels.cuh', 'psm_internal.h', os.path.join('..', '..', 'include', 'psm_b200.h')]
STAMP = OUT + '.srchash'
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-lineinfo',
         '-Xcompiler', '-fPIC,-Wall,-Wno-unused-function', '-shared']


def source_hash():
    """SHA-256 over the compiler flags and the bytes of every source and header: a library that travelled with the tree is
    reused only if it was built from exactly these bytes (file times say nothing after a copy)."""
    h = hashlib.sha256(' '.join(FLAGS).encode())
    for f in SOURCES + HEADERS:
        h.update(f.encode())
        with open(os.path.join(CSRC, f), 'rb') as fh:
            h.update(fh.read())
    return h.hexdigest()


def up_to_date():
    if not (os.path.exists(OUT) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as fh:
        return fh.read().strip() == source_hash()


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    tmp = OUT + '.tmp%d' % os.getpid()                    # link next to the target, then rename: a reader never sees a half-written file
    cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-o', tmp] + SOURCES
    r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError('nvcc failed building libpsm_b200.so')
    os.replace(tmp, OUT)
    if verbose:
        sys.stderr.write(r.stderr)
    with open(STAMP, 'w') as fh:
        fh.write(source_hash() + '\n')
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
