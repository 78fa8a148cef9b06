"""time the device Laplace sampler (csrc/bc_sampler.cu) against the host one; run on a GPU box"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'beta-cores_b200'), os.path.join(ROOT, 'beta-cores_b200', 'examples', 'common')):
    sys.path.insert(0, p)
import numpy as np
import torch
from threadpoolctl import threadpool_limits
import model_lr

D, S = 128, 1024
with threadpool_limits(1, 'blas'):
    for M in (5, 50, 300):
        r = np.random.RandomState(M)
        Z = r.randn(M, D); w = r.rand(M)*1e4
        res = {'M': M}
        for method in ('newton', 'hybrid', 'device'):
            np.random.seed(0)
            s = model_lr.make_laplace_sampler(D, method=method, prefetch=True)
            for k in range(3):
                out = s(S, w*(1+0.01*k), Z)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for k in range(20):
                out = s(S, w*(1+0.01*(k+3)), Z)
                if method == 'newton':  # host samples are uploaded by the projector
                    out = torch.from_numpy(out).cuda()
            torch.cuda.synchronize()
            res[method+'_ms'] = 1e3*(time.perf_counter()-t0)/20
            if method == 'device':
                res['newton_steps'] = s.status()[1]
            s.drain()
        print(json.dumps(res))

# kernel-only timing through the C ABI
import ctypes
from bayesiancoresets import _native as nv
from bayesiancoresets._device import Engine, ptr, stream_ptr
eng = Engine.get(); ctx = eng.ctx('sampler')
for M in (5, 50, 300):
    r = np.random.RandomState(M)
    Zd = eng.upload(r.randn(M, D)); wd = eng.upload(r.rand(M)*1e4)
    mu = eng.zeros(D); L = eng.empty(D, D); info = torch.zeros(2, dtype=torch.int32, device=eng.device)
    Rd = eng.upload(r.randn(S, D)); th = eng.empty(S, D)
    nv.call('bc_laplace_logistic', ctx, ptr(Zd), D, ptr(wd), M, D, ptr(mu), ptr(L), 200, 1e-13, ptr(info), stream_ptr())
    cold_steps = int(info.cpu()[1])
    e0, e1, e2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    torch.cuda.synchronize()
    e0.record()
    for k in range(10):
        nv.call('bc_laplace_logistic', ctx, ptr(Zd), D, ptr(wd), M, D, ptr(mu), ptr(L), 200, 1e-13, ptr(info), stream_ptr())
    e1.record()
    for k in range(10):
        nv.call('bc_sample_affine', ctx, ptr(mu), ptr(L), ptr(Rd), S, D, ptr(th), D, stream_ptr())
    e2.record()
    torch.cuda.synchronize()
    print(json.dumps({'M': M, 'laplace_kernel_ms_warm': e0.elapsed_time(e1)/10, 'steps_warm': int(info.cpu()[1]), 'steps_cold': cold_steps,
                      'affine_kernel_ms': e1.elapsed_time(e2)/10}))
