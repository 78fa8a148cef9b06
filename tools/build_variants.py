"""Build experiment / debug variants of libbetacores.so into beta-cores_b200/lib/variants/.
    python tools/build_variants.py NAME="-DFLAG ..." [NAME2="..."]          every unit recompiled with the flags
    python tools/build_variants.py q:NAME="-DFLAG"                          only bc_project_q.cu differs (quick kernel experiments)
Select one at run time with BC_LIB_PATH=beta-cores_b200/lib/variants/libbetacores_NAME.so
    debug="-DBC_DEBUG"   device-side invariant checks + mbarrier watchdogs (csrc/bc_common.cuh); the GPU suite is run against it
                         once per round (compute-sanitizer is closed on this pool)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'beta-cores_b200')
sys.path.insert(0, PKG)
import build as B
B.build()
out = os.path.join(B.LIB, 'variants')
os.makedirs(out, exist_ok=True)
jobs = []
for spec in sys.argv[1:]:
    name, flags = spec.split('=', 1)
    only_q = name.startswith('q:')
    name = name[2:] if only_q else name
    units = ['bc_project_q.cu'] if only_q else B.UNITS
    objs = []
    for u in B.UNITS:
        if u in units and u.endswith('.cu'):
            obj = os.path.join(out, '%s_%s.o' % (u[:-3], name))
            jobs.append((name, subprocess.Popen([B.NVCC] + B.FLAGS + flags.split() + ['-c', os.path.join(B.CSRC, u), '-o', obj])))
        else:
            obj = os.path.join(B.LIB, os.path.splitext(u)[0] + '.o')
        objs.append(obj)
    jobs.append((name, objs))
links = {}
for name, j in jobs:
    if isinstance(j, list):
        links[name] = j
    else:
        assert j.wait() == 0, name
for name, objs in links.items():
    so = os.path.join(out, 'libbetacores_%s.so' % name)
    subprocess.run([B.NVCC, '-shared', '-o', so] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a'], check=True)
    print(so)
