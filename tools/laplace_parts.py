"""where the Laplace sampler kernel's time goes (D = 128): warm calls with the Newton search (maxit = 200) against calls that only
form the Hessian at the given point and factorise it (maxit = 0), for a few coreset sizes M"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'beta-cores_b200'), os.path.join(ROOT, 'beta-cores_b200', 'examples', 'common')):
    sys.path.insert(0, p)
import numpy as np, torch
from bayesiancoresets import _native as nv
from bayesiancoresets._device import Engine, ptr, stream_ptr
eng = Engine.get(); ctx = eng.ctx('sampler')
D = 128
for M in (3, 6, 20, 64):
    r = np.random.RandomState(M)
    Zd = eng.upload(r.randn(M, D)); wd = eng.upload(r.rand(M)*1e4)
    mu = eng.zeros(D); C = eng.empty(D, D); info = torch.zeros(2, dtype=torch.int32, device=eng.device)
    out = {}
    for name, maxit in (('newton+factor', 200),):
        for k in range(3):
            nv.call('bc_laplace_logistic_factor', ctx, ptr(Zd), D, ptr(wd), M, D, ptr(mu), ptr(C), maxit, 1e-13, ptr(info), stream_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(20):
            nv.call('bc_laplace_logistic_factor', ctx, ptr(Zd), D, ptr(wd), M, D, ptr(mu), ptr(C), maxit, 1e-13, ptr(info), stream_ptr())
        e1.record(); torch.cuda.synchronize()
        out[name] = 50*e0.elapsed_time(e1)
        out[name+' steps'] = int(info.cpu()[1])
    print('M = %3d: %s' % (M, out), flush=True)
    if os.environ.get('BC_LIB_PATH', '').endswith('laptrace.so'):
        import ctypes
        L = ctypes.CDLL(os.environ['BC_LIB_PATH'])
        buf = (ctypes.c_longlong*8)()
        L.bc_lap_trace_read(buf)
        t = list(buf)
        print('        cycles: stage rows + margins + log-joint %d | Newton search %d | Hessian %d | Cholesky %d | write-out %d | total %d'
              % (t[1]-t[0], t[2]-t[1], t[3]-t[2], t[4]-t[3], t[5]-t[4], t[5]-t[0]), flush=True)
