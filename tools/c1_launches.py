"""a few optimiser steps of BASELINE config 1 (for an ncu launch list: which kernels a 0.46 ms step is made of)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'beta-cores_b200'), os.path.join(ROOT, 'beta-cores_b200', 'examples', 'common'), os.path.join(ROOT, 'tests', 'golden')):
    sys.path.insert(0, p)
import numpy as np, torch
import problems, bayesiancoresets as bc, gaussian
case = [c for c in problems.coreset_cases(True) if c['name'] == 'c1_zellner_gaussian'][0]
prob = case['make']()
bl = gaussian.gaussian_beta_likelihood.bind(**prob['params']); ll = gaussian.gaussian_loglikelihood.bind(**prob['params'])
sampler = gaussian.make_conjugate_sampler(prob['prior']['mu0'], prob['prior']['Sig0inv'], prob['params']['Siginv'], device=True, prefetch=True)
prj = bc.BetaBlackBoxProjector(sampler, case['S'], bl, ll, None)
itrs = int(sys.argv[1]) if len(sys.argv) > 1 else 6
alg = bc.BetaCoreset(prob['data'], prj, n_subsample_select=case['n_sel'], n_subsample_opt=case['n_opt'], opt_itrs=itrs,
                     step_sched=case['sched'], beta=case['beta'], learn_beta=False)
alg.build(1, 1)
torch.cuda.synchronize()
torch.cuda.profiler.start()
alg.build(1, 2)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
sampler.drain()
print('done', alg.idcs)
