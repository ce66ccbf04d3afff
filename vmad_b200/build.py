"""Build libvmadpm.so (sm_100a) in-tree with nvcc.

``python -m vmad_b200.build`` or ``vmad_b200.build.build()``.  The shared object lands next to
this file (``vmad_b200/libvmadpm.so``): it is git-ignored but travels to the GPU box with the
repo snapshot.  nvcc cross-compiles without a GPU.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libvmadpm.so')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-lineinfo', '-std=c++17',
    '--fmad=true',            # position->cell arithmetic uses explicit __dmul_rn/__dsub_rn
    '-Xcompiler', '-fPIC,-O2,-Wall',
    '-Xptxas', '-v',
    '-shared',
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, '*.cuh')) + \
        [os.path.join(HERE, '..', 'include', 'vmadpm.h'), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get('NVCC', 'nvcc')
    cuda_lib = os.environ.get('CUDA_LIB', '/usr/local/cuda/lib64')
    cmd = [nvcc] + NVCC_FLAGS + ['-o', LIB] + sources() + \
        ['-L' + cuda_lib, '-lcufft', '-Xlinker', '-rpath,' + cuda_lib]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(HERE, 'csrc', 'build.log')
    with open(log, 'w') as f:
        f.write(' '.join(cmd) + '\n' + r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libvmadpm.so (see %s)" % log)
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
