"""host-side profile of BASELINE config 1 (examples/zellner_gaussian: N=5700 d=100 S=200, 1000 ADAM steps per point, sub-samples
1000/200): where a 0.87 ms optimiser step goes.  cProfile over two build steps after one warm-up step."""
import os, sys, time, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'beta-cores_b200'), os.path.join(ROOT, 'beta-cores_b200', 'examples', 'common'), os.path.join(ROOT, 'tests', 'golden')):
    sys.path.insert(0, p)
import numpy as np, torch
from threadpoolctl import threadpool_limits
threadpool_limits(limits=1, user_api='blas')
import problems, bayesiancoresets as bc, gaussian
case = [c for c in problems.coreset_cases(True) if c['name'] == 'c1_zellner_gaussian'][0]
prob = case['make']()
bl = gaussian.gaussian_beta_likelihood.bind(**prob['params']); ll = gaussian.gaussian_loglikelihood.bind(**prob['params'])
sampler = gaussian.make_conjugate_sampler(prob['prior']['mu0'], prob['prior']['Sig0inv'], prob['params']['Siginv'], device=True, prefetch=True)
prj = bc.BetaBlackBoxProjector(sampler, case['S'], bl, ll, None)
alg = bc.BetaCoreset(prob['data'], prj, n_subsample_select=case['n_sel'], n_subsample_opt=case['n_opt'], opt_itrs=case['opt_itrs'],
                     step_sched=case['sched'], beta=case['beta'], learn_beta=False)
alg.build(1, 1)
torch.cuda.synchronize()
t0 = time.perf_counter()
alg.build(1, 2); alg.build(1, 3)
torch.cuda.synchronize()
print('2 build steps, no profiler: %.3f ms per optimiser step' % (1e3*(time.perf_counter()-t0)/2002))
# the native legacy generator against numpy's on this host (one optimiser step draws S x D = 200 x 100 normals)
import ctypes
from bayesiancoresets import _native as nv
from bayesiancoresets.util import rng as _rng
sampler.drain()
m = _rng._checkout(); buf = np.empty(20000)
for thr in (1, 2, 4, 8):
    nv.call('bc_mt_randn', ctypes.byref(m), buf.ctypes.data, 20000, thr)
    t = time.perf_counter()
    for _ in range(200):
        nv.call('bc_mt_randn', ctypes.byref(m), buf.ctypes.data, 20000, thr)
    print('bc_mt_randn(20000) %d thread(s): %.1f us' % (thr, (time.perf_counter()-t)/200*1e6))
t = time.perf_counter()
for _ in range(200):
    np.random.randn(200, 100)
print('np.random.randn(200, 100): %.1f us   (RNG_THREADS in use: %d)' % ((time.perf_counter()-t)/200*1e6, _rng.RNG_THREADS))
t0 = time.perf_counter()
pr = cProfile.Profile(); pr.enable()
alg.build(1, 4); alg.build(1, 5)
torch.cuda.synchronize()
pr.disable()
print('2 build steps: %.3f s (%.3f ms per optimiser step incl. profiler overhead)' % (time.perf_counter()-t0, 1e3*(time.perf_counter()-t0)/2002))
pstats.Stats(pr).sort_stats('tottime').print_stats(28)
