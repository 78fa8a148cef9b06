"""dev aid: run one golden case through the product and the oracle side by side (GPU box)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'beta-cores_b200'), os.path.join(ROOT, 'beta-cores_b200', 'examples', 'common'), os.path.join(ROOT, 'tests', 'golden'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import problems
import bayesiancoresets as bc
import model_lr, gaussian, model_neurlinr
from oracle import np_coresets as oc, np_snnls as osn
from test_gpu_parity import _device_potentials

name = sys.argv[1]
case = [c for c in problems.coreset_cases(True) if c['name'] == name][0]
models = (model_lr, gaussian, model_neurlinr)
np.set_printoptions(precision=12, linewidth=200)

prob = case['make']()
np.random.seed(case['seed'])
bl, ll = _device_potentials(prob, models)
if case['alg'] == 'beta':
    prj = bc.BetaBlackBoxProjector(prob['sampler'], case['S'], bl, ll, None)
    alg = bc.BetaCoreset(prob['data'], prj, n_subsample_select=case['n_sel'], n_subsample_opt=case['n_opt'], opt_itrs=case['opt_itrs'], step_sched=case['sched'], beta=case['beta'], learn_beta=False)
elif case['alg'] == 'svi':
    prj = bc.BlackBoxProjector(prob['sampler'], case['S'], ll, None)
    alg = bc.SparseVICoreset(prob['data'], prj, n_subsample_select=case['n_sel'], n_subsample_opt=case['n_opt'], opt_itrs=case['opt_itrs'], step_sched=case['sched'])
else:
    prj = bc.BlackBoxProjector(prob['sampler'], case['S'], ll, None)
    alg = bc.HilbertCoreset(prob['data'], prj, n_subsample=case['n_sel'], snnls=getattr(bc.snnls, case['solver']))
hist = []
for m in range(1, case['M']+1):
    alg.build(1, m)
    if case['alg'] in ('beta', 'svi'):
        hist.append((dict(alg._last_select), alg.wts.copy(), alg.idcs.copy()))
    else:
        hist.append((None, alg.snnls.weights()[alg.snnls.weights() != 0], np.nonzero(alg.snnls.weights())[0], alg.snnls.error()))

prob = case['make']()
np.random.seed(case['seed'])
if case['alg'] in ('beta', 'svi'):
    pot = prob['oracle_betalik'](case['beta']) if case['alg'] == 'beta' else prob['oracle_loglik']()
    o = oc.GreedyVI(prob['data'], prob['sampler'], case['S'], pot, n_sub_select=case['n_sel'], n_sub_opt=case['n_opt'], opt_itrs=case['opt_itrs'], sched=case['sched'])
else:
    o = oc.Hilbert(prob['data'], prob['sampler'], case['S'], prob['oracle_loglik'](), n_sub=case['n_sel'], solver={'GIGA': 'giga', 'FrankWolfe': 'fw', 'OrthoPursuit': 'omp'}[case['solver']])
with np.errstate(all='ignore'):
    for m in range(1, case['M']+1):
        o.build(1, m)
        print('--- step', m)
        if case['alg'] in ('beta', 'svi'):
            print(' oracle ', o.log[-1], o.wts, o.idcs)
            print(' product', hist[m-1][0], hist[m-1][1], hist[m-1][2])
        else:
            w = o.solver.w
            print(' oracle ', o.solver.trace[-1:], w[w != 0], np.nonzero(w)[0], o.solver.error())
            print(' product', hist[m-1][1], hist[m-1][2], hist[m-1][3])
