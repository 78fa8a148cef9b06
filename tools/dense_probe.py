"""GPU probe: HBM throughput of the stage-2 kernels on a materialised n x S matrix (snnls / Hilbert path)."""
import ctypes, os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'beta-cores_b200'), os.path.join(ROOT, 'beta-cores_b200', 'examples', 'common')]
import torch
from bayesiancoresets import _native as nv
from bayesiancoresets._device import Engine, ptr, stream_ptr
eng = Engine.get(); ctx = eng.ctx('dense'); dev = eng.device
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
S = 1024
V = torch.randn(n, S, dtype=torch.float64, device=dev)
norms = torch.empty(n, dtype=torch.float64, device=dev)
u = torch.randn(2*S, dtype=torch.float64, device=dev)
out = torch.zeros(4, dtype=torch.float64, device=dev)
dd = torch.empty(2*(S+1), dtype=torch.float64, device=dev)
def timeit(fn, reps=10):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps
res = {}
byt = 8.*n*S
for name, fn in [
    ('rownorms', lambda: nv.call('bc_dense_rownorms', ctx, ptr(V), n, S, S, ptr(norms), stream_ptr())),
    ('score_fw', lambda: nv.call('bc_dense_score', ctx, nv.SCORE_FW, ptr(V), n, S, S, ptr(norms), ptr(u), None, 0, ptr(out), None, stream_ptr())),
    ('score_giga', lambda: nv.call('bc_dense_score', ctx, nv.SCORE_GIGA, ptr(V), n, S, S, ptr(norms), ptr(u), None, 0, ptr(out), None, stream_ptr())),
    ('score_corr', lambda: nv.call('bc_dense_score', ctx, nv.SCORE_CORR, ptr(V), n, S, S, None, ptr(u), None, 0, ptr(out), None, stream_ptr())),
    ('colsum', lambda: nv.call('bc_dense_colsum', ctx, ptr(V), n, S, S, ptr(dd), stream_ptr())),
]:
    ms = timeit(fn)
    res[name] = {'ms': ms, 'GBs': byt/ms/1e6}
    print(name, '%.3f ms  %.0f GB/s' % (ms, byt/ms/1e6), flush=True)
b = torch.empty_like(V)
ms = timeit(lambda: b.copy_(V), 5)
print('torch copy', '%.3f ms  %.0f GB/s (read+write)' % (ms, 2*byt/ms/1e6))
