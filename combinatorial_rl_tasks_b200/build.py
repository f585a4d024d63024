"""Builds libcrl_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc', 'crl_kernels.cu')
LIB = os.path.join(HERE, 'libcrl_b200.so')
DEPS = [SRC, os.path.join(HERE, 'csrc', 'crl_core.cuh'),
        os.path.join(os.path.dirname(HERE), 'include', 'crl_b200.h')]

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-diag-suppress', '177', '-shared', '-Xcompiler', '-fPIC']


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-o', LIB, SRC]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == '__main__':
    print(build(force=True, verbose=True))
