"""Ahead-of-time build of libstv.so (nvcc, sm_100a) — in-tree, so the binary travels with the repo snapshot.

Used by `__graft_entry__.build()`; importable without a GPU (nvcc cross-compiles).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG/'csrc'
LIB = PKG/'libstv.so'
NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
    '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr', '-Xptxas', '-v',
]


def sources() -> list[Path]:
    return sorted(CSRC.glob('*.cu'))


def needs_build() -> bool:
    if not LIB.is_file(): return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob('*.cu')) + list(CSRC.glob('*.cuh')) + [PKG.parent/'include'/'stv.h']
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu under csrc/ into libstv.so. Objects are cached per source under csrc/build/."""
    if not force and not needs_build(): return LIB
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    objdir = CSRC/'build'
    objdir.mkdir(exist_ok=True)
    hdr_t = max(p.stat().st_mtime for p in list(CSRC.glob('*.cuh')) + [PKG.parent/'include'/'stv.h'])
    objs, procs = [], []
    for src in sources():
        obj = objdir/(src.stem + '.o')
        objs.append(obj)
        if not force and obj.is_file() and obj.stat().st_mtime > max(src.stat().st_mtime, hdr_t): continue
        cmd = [nvcc, *NVCC_FLAGS, '-c', str(src), '-o', str(obj)]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, pr in procs:
        out, _ = pr.communicate()
        log.append(f'== {src.name}\n{out}')
        if pr.returncode != 0:
            sys.stderr.write('\n'.join(log))
            raise RuntimeError(f'nvcc failed on {src}')
    (objdir/'ptxas.log').write_text('\n'.join(log))
    if verbose: print('\n'.join(log))
    cmd = [nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o', str(LIB), *map(str, objs), '-lcudart']
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
