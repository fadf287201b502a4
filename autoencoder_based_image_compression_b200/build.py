"""Builds ``libeae_b200.so`` in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m autoencoder_based_image_compression_b200.build [--force]

The shared library links the CUDA runtime statically and nothing else (no torch, no libcuda at link
time: the driver entry point for TMA descriptors is resolved through ``cudaGetDriverEntryPoint``).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_PATH = os.path.join(HERE, 'libeae_b200.so')
SOURCES = ['runtime.cu', 'coder.cu', 'glue.cu', 'transforms_simt.cu', 'conv_umma.cu', 'codec.cu']
NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-std=c++17', '-lineinfo',
    '--fmad=true',            # explicit __f*_rn intrinsics guard every step that must not contract
    '-Xcompiler', '-fPIC',
    '-Xptxas', '-v',
    '-cudart', 'static',
]


def nvcc_path():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError('nvcc not found')


def _newest_source_mtime():
    newest = os.path.getmtime(os.path.abspath(__file__))
    for root in (CSRC, os.path.join(HERE, '..', 'include')):
        for name in os.listdir(root):
            newest = max(newest, os.path.getmtime(os.path.join(root, name)))
    return newest


def needs_build():
    return not os.path.isfile(LIB_PATH) or os.path.getmtime(LIB_PATH) < _newest_source_mtime()


def build(force=False, verbose=False):
    """Compiles every CUDA source for sm_100a into one shared library. Returns its path."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = nvcc_path()
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace('.cu', '.o'))
        cmd = [nvcc] + NVCC_FLAGS + ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = False
    for (src, obj, proc) in procs:
        out = proc.communicate()[0]
        log.append('==== {} ====\n{}'.format(src, out))
        if proc.returncode != 0:
            failed = True
        objs.append(obj)
    with open(os.path.join(objdir, 'nvcc.log'), 'w') as f:
        f.write('\n'.join(log))
    if failed or verbose:
        sys.stderr.write('\n'.join(log))
    if failed:
        raise RuntimeError('nvcc failed; see {}'.format(os.path.join(objdir, 'nvcc.log')))
    link = [nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-cudart', 'static',
            '-Xcompiler', '-fPIC', '-o', LIB_PATH + '.tmp'] + objs + ['-ldl', '-lpthread', '-lrt']
    subprocess.check_call(link)
    os.replace(LIB_PATH + '.tmp', LIB_PATH)
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
