"""warm timings of the per-step small kernels (set_samples, finalize/combine, core-side step); run on a GPU box"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'beta-cores_b200')]
import numpy as np, torch, ctypes
from bayesiancoresets import _native as nv
from bayesiancoresets._device import Engine, ptr, stream_ptr
eng = Engine.get(); ctx = eng.ctx('t'); dev = eng.device
N, D, S = 200_000, 128, 1024
X = torch.randn(N, D, dtype=torch.float64, device=dev)
T = torch.randn(S, D, dtype=torch.float64, device=dev)/np.sqrt(D)
nv.call('bc_set_potential', ctx, nv.MODEL_LOGISTIC, nv.KIND_BETALIK, D, nv.params8([0.1, 11.0, 0, 0, 0, 0, 0, 0]), None)
def tm(fn, n=200):
    for _ in range(5): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return 1e3*e0.elapsed_time(e1)/n
out = {}
out['set_samples_us'] = tm(lambda: nv.call('bc_set_samples', ctx, ptr(T), S, D, stream_ptr()))
nb = ctypes.c_int64(); nv.call('bc_q_image_bytes', N, ctypes.byref(nb))
img = torch.empty(nb.value, dtype=torch.uint8, device=dev); rs = torch.empty(N, dtype=torch.float64, device=dev)
nv.call("bc_quantise_rows", ctx, ptr(X), D, N, D, 0, ptr(img), ptr(rs), None, None, stream_ptr())
o = torch.empty(2*(S+1), dtype=torch.float64, device=dev)
n_small = 148*128
out['colsum_q_one_tile_per_sm_us'] = tm(lambda: nv.call('bc_project_colsum_q', ctx, ptr(img), ptr(rs), n_small, None, ptr(o), stream_ptr()), 50)
out['colsum_q_tiny_us'] = tm(lambda: nv.call('bc_project_colsum_q', ctx, ptr(img), ptr(rs), 128, None, ptr(o), stream_ptr()), 50)
M = 16
V = torch.empty(M, S, dtype=torch.float64, device=dev)
out['materialise_M16_us'] = tm(lambda: nv.call('bc_project_materialise', ctx, ptr(X), D, None, M, None, ptr(V), S, None, None, 0, stream_ptr()), 100)
cs = torch.empty(S, dtype=torch.float64, device=dev)
out['colsum_combine_us'] = tm(lambda: nv.call('bc_colsum_combine', ctx, ptr(o), 1, S, ptr(cs), stream_ptr()))
print(json.dumps(out, indent=1))
