"""a few warm calls of bc_laplace_logistic_factor (D=128, M=10: the north-star optimiser step's sampler kernel) for ncu"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'beta-cores_b200'), os.path.join(ROOT, 'beta-cores_b200', 'examples', 'common')):
    sys.path.insert(0, p)
import numpy as np, torch
from bayesiancoresets import _native as nv
from bayesiancoresets._device import Engine, ptr, stream_ptr
eng = Engine.get(); ctx = eng.ctx('sampler')
D, M, S = 128, int(sys.argv[1]) if len(sys.argv) > 1 else 10, 1024
r = np.random.RandomState(M)
Zd = eng.upload(r.randn(M, D)); wd = eng.upload(r.rand(M)*1e4)
mu = eng.zeros(D); C = eng.empty(D, D); info = torch.zeros(2, dtype=torch.int32, device=eng.device)
Rd = eng.upload(r.randn(S, D)); th = eng.empty(S, D)
for k in range(3):
    nv.call('bc_laplace_logistic_factor', ctx, ptr(Zd), D, ptr(wd), M, D, ptr(mu), ptr(C), 200, 1e-13, ptr(info), stream_ptr())
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
torch.cuda.profiler.start()
e[0].record()
for k in range(10):
    nv.call('bc_laplace_logistic_factor', ctx, ptr(Zd), D, ptr(wd), M, D, ptr(mu), ptr(C), 200, 1e-13, ptr(info), stream_ptr())
e[1].record()
for k in range(10):
    nv.call('bc_sample_solve', ctx, ptr(mu), ptr(C), ptr(Rd), S, D, ptr(th), D, stream_ptr())
e[2].record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('laplace_factor %.1f us (steps %d), sample_solve %.1f us' % (100*e[0].elapsed_time(e[1]), int(info.cpu()[1]), 100*e[1].elapsed_time(e[2])))
