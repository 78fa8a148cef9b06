import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'beta-cores_b200'), os.path.join(ROOT, 'beta-cores_b200', 'examples', 'common'), os.path.join(ROOT, 'tests', 'golden'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import problems
import bayesiancoresets as bc
from bayesiancoresets._fused import FusedProjection
from bayesiancoresets._device import Engine, DeviceRows
import model_neurlinr
from oracle import np_coresets as oc, np_models as om
np.set_printoptions(precision=17, linewidth=200)
case = [c for c in problems.coreset_cases(True) if c['name'] == 'nl_beta_small'][0]
prob = case['make']()
np.random.seed(case['seed'])
o = oc.GreedyVI(prob['data'], prob['sampler'], case['S'], prob['oracle_betalik'](case['beta']), opt_itrs=case['opt_itrs'], sched=case['sched'])
with np.errstate(all='ignore'):
    for m in range(1, 5):
        o.build(1, m)
    th = prob['sampler'](case['S'], o.wts, o.pts)
    F = om.nl_betalik(prob['data'], th, case['beta'], 0.5)
    V = oc.centred(F.copy())
    core = oc.centred(om.nl_betalik(o.pts, th, case['beta'], 0.5))
    resid = V.sum(axis=0) - o.wts.dot(core)
    corrs = V.dot(resid)/np.sqrt((V**2).sum(axis=1))/V.shape[1]
print('oracle nan rows', np.nonzero(np.isnan(corrs))[0], 'argmax', np.argmax(corrs))
bad = np.nonzero(np.isnan(corrs))[0]
print('F rows', F[bad[0]][:6], 'data', prob['data'][bad[0]])
eng = Engine.get()
pot = model_neurlinr.neurlinr_beta_likelihood.bind(sigsq=0.5)
fp = FusedProjection(eng, pot, prob['data'].shape[1])
fp.configure(case['beta']); fp.set_samples(th)
rows = DeviceRows(eng, prob['data'])
out = eng.zeros(4); sc = eng.empty(prob['data'].shape[0])
fp.score(rows, None, eng.upload(np.concatenate((resid, [resid.sum()]))), 0, out, scores=sc)
s = sc.cpu().numpy()
print('device nan rows', np.nonzero(np.isnan(s))[0], 'at bad', s[bad], 'out', out.cpu().numpy())
dF, _, _ = fp.materialise(rows, raw=True)
dF = dF.cpu().numpy()
print('device F at bad', dF[bad[0]][:6], 'max |dF-F| at bad', np.abs(dF[bad]-F[bad]).max())
dV, dn, _ = fp.materialise(rows, want_norms=True)
print('device centred at bad', dV.cpu().numpy()[bad[0]][:6], 'norm', dn.cpu().numpy()[bad])
