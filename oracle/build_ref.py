"""TEST INFRASTRUCTURE -- builds the reference's own CUDA grid encoder as a second oracle.

Compiles /root/reference/NeRF_LiDAR/zipnerf/gridencoder/src/{gridencoder.cu,bindings.cpp}
WHERE THEY LIE (no reference source is copied into this repository) into the git-ignored
`oracle/_ref/_gridencoder_ref.so`, a torch extension exporting the reference's three functions
(Z/gridencoder/src/bindings.cpp:5-9).  The only change against the reference's own build recipe
(Z/gridencoder/backend.py:6-12) is `-std=c++17` (torch 2.11 headers need it) and the explicit
`-gencode arch=compute_100a,code=sm_100a`; its `-O3` and the three `-U__CUDA_NO_HALF*` flags are kept,
and no fast-math flag is added, so the arithmetic of the binary is the reference's.

nvcc cross-compiles here without a GPU; the .so travels to the GPU box with the snapshot
(`oracle/_ref/` is git-ignored, not gpurun-ignored).  On the box `/root/reference` does not exist:
tests only ever LOAD the prebuilt file (oracle/ref_grid.py) and skip when it is absent.

Only tests/, __graft_entry__.build()/smoke() and bench.py's reference-kernel leg may touch this."""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, '_ref')
REF_SRC = '/root/reference/NeRF_LiDAR/zipnerf/gridencoder/src'
NAME = '_gridencoder_ref'
SO = os.path.join(OUT_DIR, NAME + '.so')


def available() -> bool:
    return os.path.exists(SO)


def build(force: bool = False, verbose: bool = False) -> str | None:
    """Returns the path of the built extension, or None when the reference sources are not present
    (the GPU box) and nothing prebuilt exists."""
    srcs = [os.path.join(REF_SRC, 'gridencoder.cu'), os.path.join(REF_SRC, 'bindings.cpp')]
    if not all(os.path.exists(s) for s in srcs):
        return SO if os.path.exists(SO) else None
    if os.path.exists(SO) and not force and all(os.path.getmtime(SO) >= os.path.getmtime(s) for s in srcs):
        return SO
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(OUT_DIR, exist_ok=True)
    inc = []
    for p in ce.include_paths('cuda') + [sysconfig.get_paths()['include']]:
        inc += ['-isystem', p]
    defs = [f'-DTORCH_EXTENSION_NAME={NAME}', '-DTORCH_API_INCLUDE_EXTENSION_H',
            f'-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}']
    nvcc = os.environ.get('NVCC') or '/usr/local/cuda/bin/nvcc'
    o_cu, o_cpp = os.path.join(OUT_DIR, 'gridencoder.o'), os.path.join(OUT_DIR, 'bindings.o')
    cmds = [
        [nvcc, '-O3', '-std=c++17', '-U__CUDA_NO_HALF_OPERATORS__', '-U__CUDA_NO_HALF_CONVERSIONS__',
         '-U__CUDA_NO_HALF2_OPERATORS__', '-gencode', 'arch=compute_100a,code=sm_100a', '--expt-relaxed-constexpr',
         '-Xcompiler', '-fPIC', '-w'] + defs + inc + ['-c', srcs[0], '-o', o_cu],
        ['g++', '-O3', '-std=c++17', '-fPIC', '-w'] + defs + inc + ['-c', srcs[1], '-o', o_cpp],
    ]
    procs = [subprocess.Popen(c, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for c in cmds]
    for c, p in zip(cmds, procs):
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError('reference grid encoder failed to compile:\n' + ' '.join(c) + '\n' + out)
        if verbose:
            sys.stderr.write(out)
    libs = []
    for p in ce.library_paths('cuda'):
        libs += ['-L' + p, '-Wl,-rpath,' + p]
    link = ['g++', '-shared', o_cu, o_cpp, '-o', SO] + libs + \
           ['-lc10', '-lc10_cuda', '-ltorch_cpu', '-ltorch_cuda', '-ltorch', '-ltorch_python', '-lcudart']
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError('reference grid encoder failed to link:\n' + r.stdout)
    for o in (o_cu, o_cpp):
        os.remove(o)
    return SO


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
