"""In-situ kernel timeline of a few optimiser steps (CUPTI through torch.profiler; ncu serialises launches and cannot show the
spacing between them): start, duration and the gap to the previous kernel's end, for one BetaCoreset.build(1, m) step of the
north-star workload on a row shard.      python tools/step_timeline.py [rows]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'beta-cores_b200'), os.path.join(ROOT, 'beta-cores_b200', 'examples', 'common')):
    sys.path.insert(0, p)
import numpy as np, torch
import bayesiancoresets as bc, model_lr
from bayesiancoresets._device import Engine, DeviceRows
from torch.profiler import profile, ProfilerActivity
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
D, S, beta = 128, 1024, 0.1
eng = Engine.get()
g = torch.Generator(device='cuda'); g.manual_seed(0)
X = torch.randn(N, D, dtype=torch.float64, device='cuda', generator=g)
y = torch.where(torch.rand(N, device='cuda', generator=g) < torch.sigmoid(X.sum(1)/np.sqrt(D)), 1., -1.).double()
Z = X*y[:, None]
np.random.seed(1)
sampler = model_lr.make_laplace_sampler(D, method='hybrid', prefetch=True)
prj = bc.BetaBlackBoxProjector(sampler, S, model_lr.beta_likelihood, model_lr.log_likelihood, None)
alg = bc.BetaCoreset(DeviceRows.from_device(eng, Z, row0=0, n_total=N), prj, opt_itrs=20, step_sched=lambda i: 1./(1.+i), beta=beta, learn_beta=False)
for m in (1, 2, 3):
    alg.build(1, m)
torch.cuda.synchronize()
sampler.state['prof'] = []
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    alg.build(1, 4)
    torch.cuda.synchronize()
evs = []
for e in prof.events():
    if str(e.device_type).endswith('CUDA'):
        evs.append((e.time_range.start, e.time_range.end - e.time_range.start, e.name))
evs.sort()
print('%d device activities' % len(evs))
# the window of three optimiser steps in the middle of the build step
big = [i for i, e in enumerate(evs) if 'k_project_q' in e[2]]
lo, hi = big[8], big[11]
prev_end = evs[lo][0] + evs[lo][1]
print('%10s %10s %10s  %s' % ('start us', 'dur us', 'gap us', 'activity'))
t0 = evs[lo][0]
for s, d, n in evs[lo:hi+1]:
    print('%10.1f %10.1f %10.1f  %s' % (s - t0, d, s - prev_end, n[:70]))
    prev_end = max(prev_end, s + d)

hp = np.array(sampler.state['prof'])
print('host ms inside device_step (median over %d calls): prepare+alloc %.3f | factor-kernel launch call %.3f | normals() %.3f' % (
    len(hp), *(1e3*np.median(hp, axis=0))))
