"""Builds dfol_vqa_b200/libdfol_b200.so from csrc/*.cu with nvcc for sm_100a (in-tree, so it travels with the repo)."""

import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
INCLUDE = os.path.join(os.path.dirname(HERE), 'include')
OUT = os.path.join(HERE, 'libdfol_b200.so')
OBJ_DIR = os.path.join(HERE, 'build')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC',
              '-I' + INCLUDE, '-I' + CSRC]


def _nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found')
    return exe


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    sources = sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    headers.append(os.path.join(INCLUDE, 'dfol_b200.h'))
    jobs = []
    for src in sources:
        obj = os.path.join(OBJ_DIR, src[:-3] + '.o')
        if force or _stale(obj, [os.path.join(CSRC, src)] + headers):
            jobs.append([nvcc] + NVCC_FLAGS + ['-c', os.path.join(CSRC, src), '-o', obj])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        for cmd, r in ex.map(run, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(' '.join(cmd) + '\n' + r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError('nvcc failed on %s' % cmd[-3])
    objs = [os.path.join(OBJ_DIR, s[:-3] + '.o') for s in sources]
    if force or jobs or _stale(OUT, objs):
        cmd = [nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', OUT] + objs + ['-lcudart']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError('link failed')
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
