"""Build libposerisk_b200.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc."""
from __future__ import annotations

import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
# PRK_LIB: development override (A/B runs of kernel variants built by scripts/build_variants.py)
LIB = os.environ.get('PRK_LIB') or os.path.join(CSRC, 'libposerisk_b200.so')

NVCC_FLAGS = (['-DPRK_FUSED_DEBUG'] if os.environ.get('PRK_FUSED_DEBUG') else []) + os.environ.get('PRK_NVCC_EXTRA', '').split() + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-shared', '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden',
              '--expt-relaxed-constexpr']


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def needs_build() -> bool:
    if os.environ.get('PRK_LIB'):
        return not os.path.isfile(LIB)
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, '*.h')) + glob.glob(os.path.join(CSRC, '*.cuh')) + \
        [os.path.join(HERE, '..', 'include', 'poserisk_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-o', LIB] + sources()
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == '__main__':
    import sys
    print(build_library(force=True, verbose='-v' in sys.argv))
