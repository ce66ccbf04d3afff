"""Build libvmadpm.so (sm_100a) in-tree with nvcc.

``python -m vmad_b200.build`` or ``vmad_b200.build.build()``.  The shared object lands next to
this file (``vmad_b200/libvmadpm.so``): it is git-ignored but travels to the GPU box with the
repo snapshot.  nvcc cross-compiles without a GPU.  Every ``csrc/*.cu`` is one translation unit
(no device-side linking); the units are compiled in parallel and only the stale ones again.
"""
import concurrent.futures
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(CSRC, '_obj')            # git-ignored (*.o)
LIB = os.path.join(HERE, 'libvmadpm.so')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-lineinfo', '-std=c++17',
    '--fmad=true',            # position->cell arithmetic uses explicit __dmul_rn/__dsub_rn
    '-Xcompiler', '-fPIC,-O2,-Wall',
    '-Xptxas', '-v',
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def _headers():
    return glob.glob(os.path.join(CSRC, '*.cuh')) + \
        [os.path.join(HERE, '..', 'include', 'vmadpm.h'), __file__]


def _obj(src):
    return os.path.join(OBJ, os.path.basename(src)[:-3] + '.o')


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build():
    return _stale(LIB, sources() + _headers())


def _compile(nvcc, src):
    cmd = [nvcc] + NVCC_FLAGS + ['-c', src, '-o', _obj(src)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, cmd, r


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get('NVCC', 'nvcc')
    cuda_lib = os.environ.get('CUDA_LIB', '/usr/local/cuda/lib64')
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers()
    todo = [s for s in sources() if force or _stale(_obj(s), [s] + hdrs)]
    log = []
    failed = False
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, max(1, len(todo)))) as ex:
        for src, cmd, r in ex.map(lambda s: _compile(nvcc, s), todo):
            log.append(' '.join(cmd) + '\n' + r.stdout + r.stderr)
            if r.returncode != 0:
                failed = True
                sys.stderr.write(r.stdout + r.stderr)
    if not failed:
        cmd = [nvcc, '-shared', '-o', LIB] + [_obj(s) for s in sources()] + \
            ['-L' + cuda_lib, '-lcufft', '-Xlinker', '-rpath,' + cuda_lib]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append(' '.join(cmd) + '\n' + r.stdout + r.stderr)
        if r.returncode != 0:
            failed = True
            sys.stderr.write(r.stdout + r.stderr)
    with open(os.path.join(OBJ, 'build.log'), 'w') as f:     # ptxas -v output: registers, spills
        f.write('\n'.join(log))
    if failed:
        raise RuntimeError("nvcc failed building libvmadpm.so (see %s)" % os.path.join(OBJ, 'build.log'))
    if verbose:
        sys.stderr.write('\n'.join(log))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
