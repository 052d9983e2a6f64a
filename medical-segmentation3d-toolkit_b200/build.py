"""Build the C-ABI CUDA library in-tree:  python medical-segmentation3d-toolkit_b200/build.py

nvcc cross-compiles for sm_100a without a GPU.  Output: lib/libseg3d_b200.so next to this file
(git-ignored; it travels to the GPU box with the gpurun snapshot).
"""
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
LIB = os.path.join(LIBDIR, 'libseg3d_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
         '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def _digest():
    h = hashlib.sha256()
    for f in _sources() + sorted(glob.glob(os.path.join(CSRC, '*.cuh'))) + \
            [os.path.join(HERE, '..', 'include', 'seg3d_b200.h')]:
        h.update(open(f, 'rb').read())
    h.update(' '.join(FLAGS).encode())
    return h.hexdigest()


def _file_digest(src):
    """one object file depends on its source, every header, and the flags"""
    h = hashlib.sha256()
    for f in [src] + sorted(glob.glob(os.path.join(CSRC, '*.cuh'))) + [os.path.join(HERE, '..', 'include', 'seg3d_b200.h')]:
        h.update(open(f, 'rb').read())
    h.update(' '.join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, 'build.sha256')
    dig = _digest()
    if not force and os.path.isfile(LIB) and os.path.isfile(stamp) and open(stamp).read().strip() == dig:
        return LIB
    objs = []
    procs = []
    for src in _sources():
        obj = os.path.join(LIBDIR, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        fd, fstamp = _file_digest(src), obj + '.sha256'
        if not force and not verbose and os.path.isfile(obj) and os.path.isfile(fstamp) and open(fstamp).read().strip() == fd:
            continue                                  # unchanged translation unit: keep its object file
        cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
        procs.append((src, fstamp, fd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, fstamp, fd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write('== %s ==\n%s\n' % (os.path.basename(src), out))
        if p.returncode == 0:
            open(fstamp, 'w').write(fd)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed')
    cmd = [NVCC, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIB] + objs + ['-lcudart']
    subprocess.check_call(cmd)
    open(stamp, 'w').write(dig)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
