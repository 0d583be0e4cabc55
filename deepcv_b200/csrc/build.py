""" Builds deepcv_b200/libdeepcv_b200.so in-tree with nvcc for sm_100a (no torch headers: the library is a plain C ABI). """
import concurrent.futures
import hashlib
import os
import subprocess
import sys
from pathlib import Path

CSRC = Path(__file__).resolve().parent
PKG = CSRC.parent
OUT = PKG / 'libdeepcv_b200.so'
OBJ_DIR = CSRC / 'build'
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC', '--threads', '2', *os.environ.get('DCV_NVCC_FLAGS', '').split()]


def _stamp(src: Path) -> str:
    h = hashlib.sha1()
    for dep in [src, *sorted(CSRC.glob('*.cuh')), PKG.parent / 'include' / 'deepcv_b200.h']:
        h.update(dep.read_bytes())
    h.update(' '.join(FLAGS).encode())
    return h.hexdigest()


def _compile(src: Path) -> Path:
    obj = OBJ_DIR / (src.stem + '.o')
    stamp_file = OBJ_DIR / (src.stem + '.stamp')
    stamp = _stamp(src)
    if obj.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
        return obj
    subprocess.run([NVCC, *FLAGS, '-c', str(src), '-o', str(obj)], check=True, cwd=str(CSRC))
    stamp_file.write_text(stamp)
    return obj


def build(verbose: bool = False) -> Path:
    OBJ_DIR.mkdir(exist_ok=True)
    sources = sorted(CSRC.glob('*.cu'))
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(sources))) as pool:
        objs = list(pool.map(_compile, sources))
    newest = max(o.stat().st_mtime for o in objs)
    if not OUT.exists() or OUT.stat().st_mtime < newest:
        subprocess.run([NVCC, '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o', str(OUT), *map(str, objs), '-cudart', 'static'], check=True)
    if verbose:
        print(f'built {OUT} from {len(sources)} sources')
    return OUT


if __name__ == '__main__':
    build(verbose=True)
    sys.exit(0)
