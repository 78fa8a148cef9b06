"""CPU: the numpy oracle replayed against fixtures produced by the unmodified
reference (tests/golden/make_golden.py).  Bit-exact for the model functions and
snnls solvers (same numpy build wrote the fixtures); index sequences exact and
weights to 1e-9 for the greedy builds (scipy.optimize.minimize inside the host
sampler may differ in the last bits across machines)."""
import os
import numpy as np
import pytest

import problems
from oracle import np_models as om, np_snnls as osn, np_coresets as oc

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def test_model_functions_match_reference():
    g = np.load(os.path.join(G, 'g1_models.npz'))
    p = problems.model_function_inputs()
    for k in ('Z', 'Th', 'Xg', 'Thg', 'Siginv', 'Zn', 'Thn'):
        np.testing.assert_array_equal(p[k], g[k])      # fixture inputs are the seeded inputs
    beta = float(g['beta'])
    with np.errstate(all='ignore'):
        pairs = [
            (om.lr_loglik(p['Z'], p['Th']), 'lr_loglik'),
            (om.lr_betalik(p['Z'], p['Th'], beta), 'lr_betalik'),
            (om.gauss_loglik(p['Xg'], p['Thg'], p['Siginv'], p['logdetSig']), 'gauss_loglik'),
            (om.gauss_betalik(p['Xg'], p['Thg'], beta, p['Siginv'], p['logdetSig']), 'gauss_betalik'),
            (om.gauss_betagrad(p['Xg'], p['Thg'], beta, p['Siginv'], p['logdetSig']), 'gauss_betagrad'),
            (om.nl_loglik(p['Zn'], p['Thn'], p['sigsq']), 'nl_loglik'),
            (om.nl_betalik(p['Zn'], p['Thn'], beta, p['sigsq']), 'nl_betalik'),
            (oc.centred(om.lr_betalik(p['Z'], p['Th'], beta)), 'lr_project_f'),
        ]
    for got, key in pairs:
        np.testing.assert_allclose(got, g[key], rtol=1e-13, atol=1e-300, err_msg=key)


@pytest.mark.parametrize('name', ['giga', 'fw', 'omp'])
def test_snnls_fingerprint(name):
    g = np.load(os.path.join(G, 'g2_snnls.npz'))
    V = problems.snnls_matrix()
    o = osn.SOLVERS[name](V.T, V.sum(axis=0))
    fs = []
    for _ in range(100):
        f = o.select(); fs.append(int(f)); o.reweight(f)
    np.testing.assert_array_equal(fs, g[name+'_trace'])
    np.testing.assert_allclose(o.w, g[name+'_w'], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(o.error(), g[name+'_error'], rtol=1e-9, atol=1e-9)
    o2 = osn.SOLVERS[name](V.T, V.sum(axis=0)); o2.run(100)
    np.testing.assert_allclose(o2.w, g[name+'_build_w'], rtol=1e-10, atol=1e-12)
    assert bool(o2.hit_limit) == bool(g[name+'_build_limit'])
    o2.polish()
    np.testing.assert_allclose(o2.w, g[name+'_opt_w'], rtol=1e-9, atol=1e-10)


def test_snnls_small_cases():
    g = np.load(os.path.join(G, 'g2_snnls.npz'))
    for tag, A in problems.snnls_small_cases():
        for name in ('giga', 'fw', 'omp'):
            o = osn.SOLVERS[name](A.T, A.sum(axis=0)); o.run(A.shape[0])
            np.testing.assert_allclose(o.w, g['small_%s_%s_w' % (tag, name)], rtol=1e-9, atol=1e-11, err_msg=tag+name)
            assert bool(o.hit_limit) == bool(g['small_%s_%s_limit' % (tag, name)])


def run_oracle_case(case):
    prob = case['make']()
    problems.reseed(case)
    if case['alg'] == 'bpsvi':
        o = oc.BatchPSVI(prob['data'], prob['sampler'], case['S'], prob['oracle_loglik'](), prob['oracle_gradll'](), case['opt_itrs'],
                         n_sub_opt=case['n_opt'], sched=lambda m: case['sched'])
    elif case['alg'] in ('beta', 'svi'):
        pot = prob['oracle_betalik'](case['beta']) if case['alg'] == 'beta' else prob['oracle_loglik']()
        o = oc.GreedyVI(prob['data'], prob['sampler'], case['S'], pot, n_sub_select=case['n_sel'], n_sub_opt=case['n_opt'],
                        opt_itrs=case['opt_itrs'], sched=case['sched'], groups=case['groups'])
    else:
        o = oc.Hilbert(prob['data'], prob['sampler'], case['S'], prob['oracle_loglik'](), n_sub=case['n_sel'],
                       solver={'GIGA': 'giga', 'FrankWolfe': 'fw', 'OrthoPursuit': 'omp'}[case['solver']])
    sizes, sumw = [], []
    with np.errstate(all='ignore'):
        for m in ([case['M']] if case['alg'] == 'bpsvi' else range(1, case['M']+1)):
            o.build(1, problems.build_size(case, m))
            if case['alg'] in ('beta', 'svi', 'bpsvi'):
                w, _, i = o.get()
            else:
                w, i = o.wts, o.idcs
            sizes.append(len(i)); sumw.append(w.sum())
    run_oracle_case.last = o
    return w, i, np.array(sizes), np.array(sumw)


@pytest.mark.parametrize('case', problems.coreset_cases(heavy=False), ids=lambda c: c['name'])
def test_coreset_builds_match_reference(case):
    g = np.load(os.path.join(G, 'g3_coresets.npz'))
    w, i, sizes, sumw = run_oracle_case(case)
    nm = case['name']
    np.testing.assert_array_equal(i, g[nm+'_idcs'])
    np.testing.assert_array_equal(sizes, g[nm+'_sizes'])
    np.testing.assert_allclose(w, g[nm+'_wts'], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(sumw, g[nm+'_sumw'], rtol=1e-9, atol=1e-12)
    if case['alg'] == 'bpsvi':
        np.testing.assert_allclose(run_oracle_case.last.get()[1], g[nm+'_pts'], rtol=1e-9, atol=1e-12)


def test_nan_rows_poison_selection_like_reference():
    """Rows whose centred vector is exactly zero score 0/0: np.argmax picks the first
    NaN on the empty coreset and `nan > x` blocks every later addition (bcores.py:78-81)."""
    case = [c for c in problems.coreset_cases(False) if c['name'] == 'lr_beta_zero_rows'][0]
    w, i, sizes, _ = run_oracle_case(case)
    assert len(i) == 0 and np.all(sizes == 0)
