"""Build libdam_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m deep_audio_mixer_b200.build [--force] [--verbose]

The .so is git-ignored but travels with the gpurun snapshot, so the GPU box only loads it.
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
BUILD = os.path.join(HERE, 'csrc', 'build')
LIB = os.path.join(HERE, 'libdam_b200.so')

SOURCES = ['dam_api.cu', 'stft_db.cu', 'stft_tma.cu', 'stft_tma16.cu', 'mix_kernels.cu', 'loudness.cu']

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xptxas', '-v']


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return 'nvcc'


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, 'rb') as f:
            h.update(p.encode())
            h.update(f.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _deps():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cu', '.cuh', '.h'))]
    deps.append(os.path.join(os.path.dirname(HERE), 'include', 'dam_b200.h'))
    return deps


def build_library(force: bool = False, verbose: bool = False) -> str:
    sources = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    stamp = os.path.join(BUILD, 'stamp.txt')
    digest = _digest(_deps())
    if not force and os.path.exists(LIB) and os.path.exists(stamp):
        with open(stamp) as f:
            if f.read().strip() == digest:
                return LIB
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(BUILD, src.replace('.cu', '.o'))
        cmd = [nvcc] + NVCC_FLAGS + ['-c', os.path.join(CSRC, src), '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, r

    objs = []
    with concurrent.futures.ThreadPoolExecutor(max_workers=max(1, min(8, len(sources)))) as ex:
        for src, obj, r in ex.map(compile_one, sources):
            log = os.path.join(BUILD, src.replace('.cu', '.ptxas.log'))
            with open(log, 'w') as f:
                f.write(r.stdout + r.stderr)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError('nvcc failed on %s' % src)
            if verbose:
                sys.stderr.write(r.stderr)
            objs.append(obj)
    cmd = [nvcc, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError('link failed')
    with open(stamp, 'w') as f:
        f.write(digest)
    return LIB


def build_variant(name: str, defines, source: str = 'stft_tma.cu') -> str:
    """A/B build: libdam_b200_<name>.so = the shipped objects with `source` recompiled under extra -D flags.
    Loaded with DAM_LIB_VARIANT=<name> (see _lib.py); never the default library."""
    build_library()
    nvcc = _nvcc()
    obj = os.path.join(BUILD, source.replace('.cu', '_%s.o' % name))
    cmd = [nvcc] + NVCC_FLAGS + ['-D' + d for d in defines] + ['-c', os.path.join(CSRC, source), '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError('nvcc failed on %s (%s)' % (source, name))
    sys.stderr.write(''.join(l + '\n' for l in (r.stdout + r.stderr).splitlines() if 'spill' in l and ' 0 bytes spill stores' not in l))
    objs = [os.path.join(BUILD, s.replace('.cu', '.o')) for s in SOURCES if s != source] + [obj]
    lib = os.path.join(HERE, 'libdam_b200_%s.so' % name)
    r = subprocess.run([nvcc, '-shared', '-o', lib] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart'],
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError('link failed')
    return lib


if __name__ == '__main__':
    if '--variant' in sys.argv:
        i = sys.argv.index('--variant')
        rest = sys.argv[i + 2:]
        src = 'stft_tma.cu'
        if '--source' in rest:                      # e.g. --source mix_kernels.cu
            j = rest.index('--source')
            src = rest[j + 1]
            rest = rest[:j] + rest[j + 2:]
        print(build_variant(sys.argv[i + 1], [d[2:] if d.startswith('-D') else d for d in rest], src))
    else:
        path = build_library(force='--force' in sys.argv, verbose='--verbose' in sys.argv)
        print(path)
