"""Build experiment variants of libbetacores.so (only bc_project_q.cu differs) into beta-cores_b200/lib/variants/.
    python tools/build_variants.py NAME="-DFLAG ..." [NAME2="..."]
Select one at run time with BC_LIB_PATH=beta-cores_b200/lib/variants/libbetacores_NAME.so"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'beta-cores_b200')
sys.path.insert(0, PKG)
import build as B
B.build()
out = os.path.join(B.LIB, 'variants')
os.makedirs(out, exist_ok=True)
procs = []
for spec in sys.argv[1:]:
    name, flags = spec.split('=', 1)
    obj = os.path.join(out, 'bc_project_q_%s.o' % name)
    procs.append((name, obj, subprocess.Popen([B.NVCC] + B.FLAGS + flags.split() + ['-c', os.path.join(B.CSRC, 'bc_project_q.cu'), '-o', obj])))
for name, obj, p in procs:
    assert p.wait() == 0, name
    objs = [obj if u == 'bc_project_q.cu' else os.path.join(B.LIB, u.replace('.cu', '.o')) for u in B.UNITS]
    so = os.path.join(out, 'libbetacores_%s.so' % name)
    subprocess.run([B.NVCC, '-shared', '-o', so] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a'], check=True)
    print(so)
