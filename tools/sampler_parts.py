"""component times of the host / hybrid Laplace sampler (run on a GPU box)"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'beta-cores_b200'), os.path.join(ROOT, 'beta-cores_b200', 'examples', 'common')):
    sys.path.insert(0, p)
import numpy as np
import torch
from threadpoolctl import threadpool_limits
import model_lr
import scipy.linalg as sl

D, S, M = 128, 1024, 8
r = np.random.RandomState(0)
Z = r.randn(M, D); w = r.rand(M)*1e6
def tm(fn, n=50):
    fn(); t0 = time.perf_counter()
    for _ in range(n): fn()
    return 1e3*(time.perf_counter()-t0)/n
with threadpool_limits(1, 'blas'):
    mu, LSig, LSigInv = model_lr.get_laplace(w, Z, np.zeros(D), method='newton')
    R = r.randn(S, D)
    out = {}
    out['get_laplace_warm'] = tm(lambda: model_lr.get_laplace(w*1.01, Z, mu, method='newton'))
    out['newton_mode_warm'] = tm(lambda: model_lr._newton_mode(Z, w*1.01, mu))
    H = -model_lr.hess_th_log_joint(Z, mu, w)
    out['hess'] = tm(lambda: model_lr.hess_th_log_joint(Z, mu, w))
    out['grad'] = tm(lambda: model_lr.grad_th_log_joint(Z, mu, w))
    out['log_joint'] = tm(lambda: model_lr.log_joint(Z, mu, w))
    out['cho_factor'] = tm(lambda: sl.cho_factor(H, lower=True, check_finite=False))
    out['cholesky'] = tm(lambda: np.linalg.cholesky(H))
    L = np.linalg.cholesky(H)
    out['tri_inverse'] = tm(lambda: sl.solve_triangular(L, np.eye(D), lower=True, check_finite=False))
    out['dtrtri'] = tm(lambda: sl.lapack.dtrtri(L, lower=1))
    out['gemm'] = tm(lambda: mu + R.dot(LSig.T))
    pin = torch.empty(S, D, dtype=torch.float64).pin_memory()
    def up():
        pin.numpy()[...] = R
        d = pin.to('cuda', non_blocking=True); torch.cuda.synchronize()
    out['pinned_stage_and_upload'] = tm(up)
    def up2():
        d = torch.from_numpy(R).to('cuda'); torch.cuda.synchronize()
    out['pageable_upload'] = tm(up2)
    out['randn'] = tm(lambda: np.random.randn(S, D), 10)
with threadpool_limits(4, 'blas'):
    out['gemm_4t'] = tm(lambda: mu + R.dot(LSig.T))
print(json.dumps(out, indent=1))
