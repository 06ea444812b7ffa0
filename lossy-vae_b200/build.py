"""Build liblvae_b200.so (sm_100a only) with nvcc, in-tree.

    python lossy-vae_b200/build.py [--force]

Output: lossy-vae_b200/lib/liblvae_b200.so (git-ignored, travels to the GPU box with gpurun).
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / 'csrc'
LIB_DIR = ROOT / 'lib'
OBJ_DIR = LIB_DIR / 'obj'
LIB = LIB_DIR / 'liblvae_b200.so'
INCLUDE = ROOT.parent / 'include'
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
CUFLAGS = ['-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr',
           '-I', str(INCLUDE), '-I', str(CSRC)]
CXXFLAGS = ['-O3', '-std=c++17', '-Xcompiler', '-fPIC', '-I', str(INCLUDE)]


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def _compile(src: Path):
    headers = list(CSRC.glob('*.cuh')) + list(CSRC.glob('*.h')) + list(INCLUDE.glob('*.h'))
    dig = _digest([src] + headers)
    obj = OBJ_DIR / (src.name + '.o')
    stamp = OBJ_DIR / (src.name + '.sha')
    if obj.exists() and stamp.exists() and stamp.read_text() == dig:
        return obj
    flags = (ARCH + CUFLAGS) if src.suffix == '.cu' else CXXFLAGS
    cmd = [NVCC] + flags + ['-c', str(src), '-o', str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}')
    stamp.write_text(dig)
    return obj


def build(force=False, verbose=True):
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    if force:
        for f in OBJ_DIR.glob('*'):
            f.unlink()
    srcs = sorted(CSRC.glob('*.cu')) + sorted(CSRC.glob('*.cpp'))
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(_compile, srcs))
    newest = max(o.stat().st_mtime for o in objs)
    if force or not LIB.exists() or LIB.stat().st_mtime < newest:
        cmd = [NVCC] + ARCH + ['-shared', '-o', str(LIB)] + [str(o) for o in objs] + ['-lcudart']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
        if verbose:
            print(f'built {LIB}')
    return LIB


if __name__ == '__main__':
    build(force='--force' in sys.argv)
