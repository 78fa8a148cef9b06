"""Build beta-cores_b200/lib/lib_trace.so: libbetacores with clock64 stamps in k_project_q (CTA 0 only) for tools/q_trace.py.

The stamps are patched into a temporary copy of csrc/bc_project_q.cu (the shipped kernel carries no tracing code):
  g_trace[0..3][chunk]   MMA issuer: before / after the tmem_empty wait, after the full_b wait, after issue + commit
  g_trace[4+g][i]        epilogue group g (warp quarter 0, lane 0): before the tmem_full wait of its i-th chunk
  g_trace[8+g][i]        ... after it;  g_trace[12+g][i] accumulator buffer handed back;  g_trace[16+g][i] chunk done
Usage:  python tools/make_trace_lib.py && gpurun -- 'BC_LIB_PATH=$PWD/beta-cores_b200/lib/lib_trace.so python tools/q_trace.py'
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, 'beta-cores_b200', 'csrc')
LIB = os.path.join(ROOT, 'beta-cores_b200', 'lib')


def patch(t):
    def sub(old, new):
        nonlocal t
        if old not in t:
            raise SystemExit('make_trace_lib: the kernel source changed, anchor not found:\n' + old)
        t = t.replace(old, new, 1)
    sub('template <class F, int MODE, int NS>\n__global__ void __launch_bounds__(kQThreads, 1) k_project_q(',
        '__device__ unsigned long long g_trace[24][512];\n'
        '#define TR(kind, idx) do { if (blockIdx.x == 0 && (idx) < 512) g_trace[kind][idx] = clock64(); } while (0)\n'
        'extern "C" int bc_trace_read(unsigned long long* out) { return (int)cudaMemcpyFromSymbol(out, g_trace, sizeof(g_trace)); }\n\n'
        'template <class F, int MODE, int NS>\n__global__ void __launch_bounds__(kQThreads, 1) k_project_q(')
    sub('          mbar_wait_relaxed(tmem_empty + buf, (use & 1) ^ 1);\n          if (buf) ++use1; else ++use0;\n          mbar_wait_relaxed(full_b + st, ph);\n',
        '          TR(0, itb);\n          mbar_wait_relaxed(tmem_empty + buf, (use & 1) ^ 1);\n          TR(1, itb);\n          if (buf) ++use1; else ++use0;\n'
        '          mbar_wait_relaxed(full_b + st, ph);\n          TR(2, itb);\n')
    sub('            if (c == nchunks - 1) umma_commit(empty_a);\n          }\n          __syncwarp();\n',
        '            if (c == nchunks - 1) umma_commit(empty_a);\n          }\n          __syncwarp();\n          TR(3, itb);\n')
    sub('        mbar_wait(tmem_full + buf, use & 1);\n        ++use;\n',
        '        if (q == 0 && lane == 0) TR(4 + grp, use);\n        mbar_wait(tmem_full + buf, use & 1);\n'
        '        if (q == 0 && lane == 0) TR(8 + grp, use);\n        ++use;\n')
    sub('              if (lane == 0) mbar_arrive(tmem_empty + buf);\n',
        '              if (lane == 0) mbar_arrive(tmem_empty + buf);\n              if (q == 0 && lane == 0) TR(12 + grp, use - 1);\n')
    sub('          if (lane == 0) mbar_arrive(ring_full + slot);\n        }\n',
        '          if (lane == 0) mbar_arrive(ring_full + slot);\n        }\n        if (q == 0 && lane == 0) TR(16 + grp, use - 1);\n')
    return t


def main():
    sys.path.insert(0, os.path.join(ROOT, 'beta-cores_b200'))
    import build as bld
    bld.build()
    src = os.path.join(CSRC, '_trace_q.cu')
    with open(os.path.join(CSRC, 'bc_project_q.cu')) as f:
        t = patch(f.read())
    with open(src, 'w') as f:
        f.write(t)
    try:
        obj = os.path.join(LIB, '_trace_q.o')
        subprocess.run([bld.NVCC] + bld.FLAGS + ['-c', src, '-o', obj], check=True)
        objs = [os.path.join(LIB, os.path.splitext(u)[0] + '.o') for u in bld.UNITS if u != 'bc_project_q.cu'] + [obj]
        out = os.path.join(LIB, 'lib_trace.so')
        subprocess.run([bld.NVCC, '-shared', '-o', out] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a'], check=True)
        print(out)
    finally:
        os.remove(src)


if __name__ == '__main__':
    main()
