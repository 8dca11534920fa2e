"""Cross-compile kernel variants of libposerisk_b200.so into build/variants/ (git-ignored, travels with gpurun).

usage: python scripts/build_variants.py name1="-DFOO=1 -DBAR" name2="" ...
Each variant is the full library built with the extra nvcc flags; scripts/fused_ab.py times them on a GPU box.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from poserisk_release_b200 import build as B  # noqa: E402

OUT = os.path.join(ROOT, 'build', 'variants')


def one(spec):
    name, _, flags = spec.partition('=')
    lib = os.path.join(OUT, f'lib_{name}.so')
    cmd = [os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')] + flags.split() + \
        [f for f in B.NVCC_FLAGS if f != '-DPRK_FUSED_DEBUG'] + ['-o', lib] + B.sources()
    r = subprocess.run(cmd, capture_output=True, text=True)
    return name, r.returncode, (r.stderr or '')[-2000:]


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    with ThreadPoolExecutor(4) as ex:
        for name, rc, err in ex.map(one, sys.argv[1:]):
            print(name, 'ok' if rc == 0 else f'FAILED\n{err}')
