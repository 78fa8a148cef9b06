"""timing only: colsum pass of the tensor-core route at N rows (library chosen by BC_LIB_PATH)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'beta-cores_b200')]
import numpy as np, torch, ctypes
from bayesiancoresets import _native as nv
from bayesiancoresets._device import Engine, ptr, stream_ptr
eng = Engine.get(); ctx = eng.ctx('t'); dev = eng.device
N, D, S = 1_000_000, 128, 1024
X = torch.randn(N, D, dtype=torch.float64, device=dev)
T = torch.randn(S, D, dtype=torch.float64, device=dev)/np.sqrt(D)
nv.call('bc_set_potential', ctx, nv.MODEL_LOGISTIC, nv.KIND_BETALIK, D, nv.params8([0.1, 11.0, 0, 0, 0, 0, 0, 0]), None)
nv.call('bc_set_samples', ctx, ptr(T), S, D, stream_ptr())
nb = ctypes.c_int64(); nv.call('bc_q_image_bytes', N, ctypes.byref(nb))
img = torch.empty(nb.value, dtype=torch.uint8, device=dev); rs = torch.empty(N, dtype=torch.float64, device=dev)
nv.call("bc_quantise_rows", ctx, ptr(X), D, N, D, 0, ptr(img), ptr(rs), None, None, stream_ptr())
o = torch.empty(2*(S+1), dtype=torch.float64, device=dev)
fn = lambda: nv.call('bc_project_colsum_q', ctx, ptr(img), ptr(rs), N, None, ptr(o), stream_ptr())
for _ in range(2): fn()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): fn()
e1.record(); torch.cuda.synchronize()
print(os.environ.get('BC_LIB_PATH', 'default'), '%.3f ms/pass' % (e0.elapsed_time(e1)/5), flush=True)
