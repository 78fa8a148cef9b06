import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests', 'golden'), os.path.join(ROOT, 'beta-cores_b200'), os.path.join(ROOT, 'beta-cores_b200', 'examples', 'common')]
import numpy as np
import problems
import bayesiancoresets as bc, gaussian
from oracle import np_coresets as oc
case = [c for c in problems.coreset_cases(True) if c['name'] == 'c1_zellner_gaussian'][0]
IT = int(sys.argv[1]) if len(sys.argv) > 1 else 30
prob = case['make']()
state = np.random.get_state()
o = oc.GreedyVI(prob['data'], prob['sampler'], case['S'], prob['oracle_betalik'](case['beta']), n_sub_select=case['n_sel'], n_sub_opt=case['n_opt'],
                opt_itrs=IT, sched=case['sched'])
for m in range(1, 4):
    o.build(1, m)
    print('oracle', m, o.idcs, o.wts, o.log[-1])
np.random.set_state(state)
bl = gaussian.gaussian_beta_likelihood.bind(**prob['params']); ll = gaussian.gaussian_loglikelihood.bind(**prob['params'])
prj = bc.BetaBlackBoxProjector(prob['sampler'], case['S'], bl, ll, None)
alg = bc.BetaCoreset(prob['data'], prj, n_subsample_select=case['n_sel'], n_subsample_opt=case['n_opt'], opt_itrs=IT, step_sched=case['sched'],
                     beta=case['beta'], learn_beta=False)
for m in range(1, 4):
    alg.build(1, m)
    print('device', m, alg.idcs, alg.wts, alg._last_select)
