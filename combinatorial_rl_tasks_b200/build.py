"""Builds libcrl_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc', 'crl_kernels.cu')
SRC_ENCODE = os.path.join(HERE, 'csrc', 'crl_encode.cu')     # own translation unit: see its header
LIB = os.path.join(HERE, 'libcrl_b200.so')
DEPS = [SRC, SRC_ENCODE, os.path.join(HERE, 'csrc', 'crl_core.cuh'),
        os.path.join(os.path.dirname(HERE), 'include', 'crl_b200.h')]

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-diag-suppress', '177', '-shared', '-Xcompiler', '-fPIC']


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False, defines=None, out=None):
    """defines/out: tuning variants (e.g. {'CRL_THREADS': 64} -> another .so, selected at run time
    with CRL_B200_LIB=<path>); the product is the default build."""
    out = out or LIB
    if out == LIB and not force and not stale():
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    dflags = [f'-D{k}={v}' for k, v in (defines or {}).items()]
    cmd = [nvcc] + NVCC_FLAGS + dflags + (['-Xptxas', '-v'] if verbose else []) + ['-o', out, SRC, SRC_ENCODE]
    subprocess.run(cmd, check=True)
    return out


if __name__ == '__main__':
    import sys
    if len(sys.argv) > 1 and sys.argv[1] == 'variants':       # python build.py variants 64 32
        for t in sys.argv[2:]:                                  # 64 -> CRL_THREADS=64; NAME=VAL -> -DNAME=VAL
            if '=' in t:
                k, v = t.split('=')
                print(build(force=True, defines={k: v}, out=os.path.join(HERE, f'libcrl_b200_{k.lower()}.so')))
            else:
                print(build(force=True, defines={'CRL_THREADS': int(t)}, out=os.path.join(HERE, f'libcrl_b200_t{t}.so')))
    else:
        print(build(force=True, verbose=True))
