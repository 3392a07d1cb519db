"""In-tree nvcc build of liboake_b200.so (sm_100a only).

`python -m oadp_b200.build` or `__graft_entry__.build()`.  The .so is git-ignored but travels to
the GPU box with the repo snapshot.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import hashlib
import os
import pathlib
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = pathlib.Path(__file__).resolve().parent
CSRC = ROOT / 'csrc'
INCLUDE = ROOT.parent / 'include'
LIB = ROOT / 'liboake_b200.so'
OBJ_DIR = ROOT / 'build'

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
    '-Xcompiler', '-fPIC', '-Xcompiler', '-Wall',
    '-Xptxas', '-v',
]


def _nvcc() -> str:
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found (set $NVCC)')


def sources() -> list[pathlib.Path]:
    return sorted(CSRC.glob('*.cu'))


def _digest(extra_flags: list[str]) -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob('*.cu')) + list(CSRC.glob('*.cuh')) + list(CSRC.glob('*.h')) + list(INCLUDE.glob('*.h'))):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(' '.join(NVCC_FLAGS + extra_flags).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, extra_flags: list[str] | None = None) -> pathlib.Path:
    extra_flags = list(extra_flags or [])
    if os.environ.get('OAKE_USE_BF16') == '1':
        extra_flags.append('-DOAKE_USE_BF16')
    extra_flags += os.environ.get('OAKE_NVCC_FLAGS', '').split()  # developer A/B builds (-DOAKE_ATTN_POLY=0 ...)
    stamp = OBJ_DIR / 'stamp'
    digest = _digest(extra_flags)
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB
    OBJ_DIR.mkdir(exist_ok=True)
    nvcc = _nvcc()
    logs: dict[str, str] = {}

    def compile_one(src: pathlib.Path) -> pathlib.Path:
        obj = OBJ_DIR / (src.stem + '.o')
        cmd = [nvcc, *NVCC_FLAGS, *extra_flags, '-I', str(INCLUDE), '-c', str(src), '-o', str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        logs[src.name] = r.stdout + r.stderr
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src.name}:\n{r.stdout}\n{r.stderr}')
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [nvcc, '-shared', '-cudart', 'static', '-o', str(LIB), *map(str, objs)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    (OBJ_DIR / 'ptxas.log').write_text('\n'.join(f'== {k}\n{v}' for k, v in sorted(logs.items())))
    stamp.write_text(digest)
    if verbose:
        print((OBJ_DIR / 'ptxas.log').read_text())
    return LIB


if __name__ == '__main__':
    p = build(force='--force' in sys.argv, verbose='-v' in sys.argv)
    print(p)
