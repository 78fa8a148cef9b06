"""Build libbetacores.so (sm_100a) in-tree:  python beta-cores_b200/build.py [--force]

nvcc cross-compiles without a GPU.  The .so lands in beta-cores_b200/lib/ (git-ignored, but it
travels to the GPU box with the gpurun snapshot)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'lib')
OUT = os.path.join(LIB, 'libbetacores.so')
UNITS = ['bc_project.cu', 'bc_project_q.cu', 'bc_small.cu', 'bc_sampler.cu', 'bc_dense.cu', 'bc_api.cu', 'bc_hostrng.cpp']
# bc_hostrng.cpp is host-only code (g++ directly) that must reproduce numpy's floating-point results bit for bit: no FMA contraction
CXX = os.environ.get('CXX', 'g++')
CXXFLAGS = ['-O3', '-std=c++17', '-fPIC', '-ffp-contract=off', '-pthread']
HEADERS = ['bc_common.cuh', 'bc_npmean.h', 'bc_umma.cuh', 'bc_fastmath.cuh', 'bc_models.cuh', 'bc_kernels.h', os.path.join('..', '..', 'include', 'betacores.h')]
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC']


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(LIB, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objs, jobs = [], []
    for u in UNITS:
        src = os.path.join(CSRC, u)
        obj = os.path.join(LIB, os.path.splitext(u)[0] + '.o')
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            if u.endswith('.cpp'):
                jobs.append([CXX] + CXXFLAGS + ['-c', src, '-o', obj])
            else:
                jobs.append([NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj])
    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed: %s\n%s' % (' '.join(cmd), r.stderr))
        return r.stderr
    with ThreadPoolExecutor(max_workers=4) as ex:
        logs = list(ex.map(run, jobs))
    if jobs or not os.path.exists(OUT):
        run([NVCC, '-shared', '-o', OUT] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a'])
    if verbose:
        sys.stderr.write(''.join(logs))
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
