"""Builds lightgrad_b200/lib/liblightgrad_b200.so from csrc/*.cu with nvcc for sm_100a.

    python -m lightgrad_b200.build [--force] [-v]

Objects are cached under lightgrad_b200/csrc/build/ and rebuilt when their source or any header
changed.  nvcc cross-compiles without a GPU, so this runs in the CPU-only build container; the
resulting .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(CSRC, 'build')
LIB = os.path.join(HERE, 'lib', 'liblightgrad_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-lineinfo',
         '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra_flags=(), lib=LIB, obj_dir=OBJ):
    FLAGS_ = FLAGS + list(extra_flags)
    OBJ, LIB = obj_dir, lib
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    sources = sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    headers.append(os.path.join(HERE, '..', 'include', 'lightgrad_b200.h'))
    jobs = []
    for s in sources:
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ, s[:-3] + '.o')
        if force or _stale(obj, [src] + headers):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [NVCC] + FLAGS_ + ['-c', src, '-o', obj]
        if verbose:
            print(' '.join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r.returncode, r.stdout + r.stderr

    failed = False
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as pool:
        for src, rc, log in pool.map(compile_one, jobs):
            if rc != 0:
                failed = True
                sys.stderr.write("nvcc failed on %s:\n%s\n" % (src, log))
            elif verbose and log.strip():
                print(log)
    if failed:
        raise RuntimeError("lightgrad_b200: CUDA build failed")
    objs = [os.path.join(OBJ, s[:-3] + '.o') for s in sources]
    if force or jobs or _stale(LIB, objs):
        cmd = [NVCC, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIB] + objs + ['-ldl']
        if verbose:
            print(' '.join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
