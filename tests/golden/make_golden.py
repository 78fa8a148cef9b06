"""Generate the golden fixtures in tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):   python tests/golden/make_golden.py

It imports dionman/beta-cores through three in-process shims (no
reference file is edited or copied), runs the hot path on small seeded problems,
checks that the numpy oracle in oracle/ reproduces the reference BIT FOR BIT on
the same inputs, and stores inputs' seeds + the reference outputs as .npz.
The fixtures are what `tests/test_oracle_golden.py` (CPU) and the `-m gpu`
parity tests replay.
"""
import os
import sys
import types
import time
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import problems  # shared seeded problem definitions (tests/golden/problems.py)


def import_reference():
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, 'examples', 'common'))
    sys.modules['iwg'] = types.ModuleType('iwg')                    # imported by util/opt.py:5, exists nowhere
    stub = types.ModuleType('bayesiancoresets.coreset.dpbpsvi')     # imported by coreset/__init__.py:6, file absent
    class DiffPrivBatchPSVICoreset(object):
        pass
    stub.DiffPrivBatchPSVICoreset = DiffPrivBatchPSVICoreset
    sys.modules['bayesiancoresets.coreset.dpbpsvi'] = stub
    import bayesiancoresets as bc
    import model_lr, gaussian, model_neurlinr
    return bc, model_lr, gaussian, model_neurlinr


def same(a, b, what):
    a = np.asarray(a); b = np.asarray(b)
    if a.shape != b.shape or not np.array_equal(a, b, equal_nan=True):
        raise SystemExit('oracle != reference for %s (max abs diff %g)' % (what, np.max(np.abs(a-b)) if a.shape == b.shape else -1))


def empty_kw():
    # dodge the reference's shared mutable default arrays (coreset.py:8)
    return dict(wts=np.array([]), idcs=np.array([], dtype=np.int64), pts=np.array([]))


def main():
    import contextlib, io
    bc, model_lr, gaussian, model_neurlinr = import_reference()
    from oracle import np_models as om, np_snnls as osn, np_coresets as oc
    out = {}
    # --only name1,name2 : regenerate just these coreset cases and merge them into the existing g3 fixture
    only = None
    if '--only' in sys.argv:
        only = set(sys.argv[sys.argv.index('--only')+1].split(','))
    if only is not None:
        g = dict(np.load(os.path.join(HERE, 'g3_coresets.npz')))
        coreset_section(bc, model_lr, gaussian, model_neurlinr, oc, g, only)
        np.savez_compressed(os.path.join(HERE, 'g3_coresets.npz'), **g)
        print('updated', sorted(only))
        return

    # ---- G1: model functions --------------------------------------------------
    p = problems.model_function_inputs()
    g = {}
    g['lr_loglik'] = model_lr.log_likelihood(p['Z'], p['Th'])
    g['lr_betalik'] = model_lr.beta_likelihood(p['Z'], p['Th'], p['beta'])
    with contextlib.redirect_stdout(io.StringIO()):
        g['gauss_loglik'] = gaussian.gaussian_loglikelihood(p['Xg'], p['Thg'], p['Siginv'], p['logdetSig'])
    g['gauss_betalik'] = gaussian.gaussian_beta_likelihood(p['Xg'], p['Thg'], p['beta'], p['Siginv'], p['logdetSig'])
    g['gauss_betagrad'] = gaussian.gaussian_beta_gradient(p['Xg'], p['Thg'], p['beta'], p['Siginv'], p['logdetSig'])
    g['nl_loglik'] = model_neurlinr.neurlinr_loglikelihood(p['Zn'], p['Thn'], p['sigsq'])
    g['nl_betalik'] = model_neurlinr.neurlinr_beta_likelihood(p['Zn'], p['Thn'], p['beta'], p['sigsq'])
    same(om.lr_loglik(p['Z'], p['Th']), g['lr_loglik'], 'lr_loglik')
    same(om.lr_betalik(p['Z'], p['Th'], p['beta']), g['lr_betalik'], 'lr_betalik')
    same(om.gauss_loglik(p['Xg'], p['Thg'], p['Siginv'], p['logdetSig']), g['gauss_loglik'], 'gauss_loglik')
    same(om.gauss_betalik(p['Xg'], p['Thg'], p['beta'], p['Siginv'], p['logdetSig']), g['gauss_betalik'], 'gauss_betalik')
    same(om.gauss_betagrad(p['Xg'], p['Thg'], p['beta'], p['Siginv'], p['logdetSig']), g['gauss_betagrad'], 'gauss_betagrad')
    same(om.nl_loglik(p['Zn'], p['Thn'], p['sigsq']), g['nl_loglik'], 'nl_loglik')
    same(om.nl_betalik(p['Zn'], p['Thn'], p['beta'], p['sigsq']), g['nl_betalik'], 'nl_betalik')
    # centred projections through the reference projector classes
    prj = bc.BetaBlackBoxProjector(lambda S, w, pts: p['Th'], p['Th'].shape[0], model_lr.beta_likelihood, model_lr.log_likelihood, None)
    g['lr_project_f'] = prj.project_f(p['Z'], p['beta'])
    same(oc.centred(om.lr_betalik(p['Z'], p['Th'], p['beta'])), g['lr_project_f'], 'project_f')
    np.savez_compressed(os.path.join(HERE, 'g1_models.npz'), **p, **g)
    print('G1 models ok')

    # ---- G2: snnls on the C2 problem (SURVEY 8c fingerprint) --------------------
    V = problems.snnls_matrix()
    g = {}
    for name, cls in (('giga', bc.snnls.GIGA), ('fw', bc.snnls.FrankWolfe), ('omp', bc.snnls.OrthoPursuit)):
        alg = cls(V.T, V.sum(axis=0))
        fs = []
        for i in range(100):
            f = alg._select(); fs.append(int(f)); alg._reweight(f)
        o = osn.SOLVERS[name](V.T, V.sum(axis=0))
        ofs = []
        for i in range(100):
            f = o.select(); ofs.append(int(f)); o.reweight(f)
        same(ofs, fs, name+' trace'); same(o.w, alg.w, name+' w'); same(o.error(), alg.error(), name+' error')
        g[name+'_trace'] = np.array(fs); g[name+'_w'] = alg.weights(); g[name+'_error'] = alg.error()
        # build() path (monotone check + retry + latch)
        alg2 = cls(V.T, V.sum(axis=0)); alg2.build(100)
        o2 = osn.SOLVERS[name](V.T, V.sum(axis=0)); o2.run(100)
        same(o2.w, alg2.w, name+' build w')
        g[name+'_build_w'] = alg2.weights(); g[name+'_build_error'] = alg2.error()
        g[name+'_build_limit'] = np.array(alg2.reached_numeric_limit)
        # optimize()
        alg2.optimize(); o2.polish()
        same(o2.w, alg2.w, name+' optimize w')
        g[name+'_opt_w'] = alg2.weights()
        print('G2', name, 'first12', fs[:12], 'size', alg.size(), 'err %.12f' % alg.error(), 'sumw %.12f' % alg.w.sum())
    # small-grid property cases (reference tests/test_snnls/test_deterministic.py:18-35 data generators)
    for tag, A in problems.snnls_small_cases():
        for name, cls in (('giga', bc.snnls.GIGA), ('fw', bc.snnls.FrankWolfe), ('omp', bc.snnls.OrthoPursuit)):
            alg = cls(A.T, A.sum(axis=0)); alg.build(A.shape[0])
            o = osn.SOLVERS[name](A.T, A.sum(axis=0)); o.run(A.shape[0])
            same(o.w, alg.w, tag+name)
            g['small_%s_%s_w' % (tag, name)] = alg.weights()
            g['small_%s_%s_limit' % (tag, name)] = np.array(alg.reached_numeric_limit)
    np.savez_compressed(os.path.join(HERE, 'g2_snnls.npz'), **g)
    print('G2 snnls ok')

    # ---- G3..: greedy coreset builds ------------------------------------------
    g = {}
    coreset_section(bc, model_lr, gaussian, model_neurlinr, oc, g, None)
    np.savez_compressed(os.path.join(HERE, 'g3_coresets.npz'), **g)
    print('all golden fixtures written')


def coreset_section(bc, model_lr, gaussian, model_neurlinr, oc, g, only):
    import contextlib, io
    for case in problems.coreset_cases():
        nm = case['name']
        if only is not None and nm not in only:
            continue
        t0 = time.time()
        # reference run
        prob = case['make']()
        problems.reseed(case)
        if case['alg'] == 'beta':
            prj = bc.BetaBlackBoxProjector(prob['sampler'], case['S'], prob['ref_betalik'](model_lr, gaussian, model_neurlinr),
                                           prob['ref_loglik'](model_lr, gaussian, model_neurlinr), None)
            alg = bc.BetaCoreset(prob['data'], prj, n_subsample_select=case['n_sel'], n_subsample_opt=case['n_opt'],
                                 opt_itrs=case['opt_itrs'], step_sched=case['sched'], beta=case['beta'], learn_beta=False,
                                 groups=case['groups'], **empty_kw())
        elif case['alg'] == 'bpsvi':
            prj = bc.BlackBoxProjector(prob['sampler'], case['S'], prob['ref_loglik'](model_lr, gaussian, model_neurlinr),
                                       prob['ref_gradll'](model_lr, gaussian, model_neurlinr))
            alg = bc.BatchPSVICoreset(prob['data'], prj, opt_itrs=case['opt_itrs'], n_subsample_opt=case['n_opt'],
                                      step_sched=lambda m: case['sched'], **empty_kw())
        elif case['alg'] == 'svi':
            prj = bc.BlackBoxProjector(prob['sampler'], case['S'], prob['ref_loglik'](model_lr, gaussian, model_neurlinr), None)
            alg = bc.SparseVICoreset(prob['data'], prj, n_subsample_select=case['n_sel'], n_subsample_opt=case['n_opt'],
                                     opt_itrs=case['opt_itrs'], step_sched=case['sched'], groups=case['groups'], **empty_kw())
        else:
            prj = bc.BlackBoxProjector(prob['sampler'], case['S'], prob['ref_loglik'](model_lr, gaussian, model_neurlinr), None)
            alg = bc.HilbertCoreset(prob['data'], prj, n_subsample=case['n_sel'], snnls=getattr(bc.snnls, case['solver']), **empty_kw())
        hist_i, hist_w = [], []
        steps = [case['M']] if case['alg'] == 'bpsvi' else list(range(1, case['M']+1))     # bpsvi: one build of M pseudo-points
        with contextlib.redirect_stdout(io.StringIO()):
            for m in steps:
                alg.build(1, problems.build_size(case, m))
                r = alg.get()
                hist_i.append(np.array(r[2]).copy()); hist_w.append(np.array(r[0]).copy())
        ref_pts = np.array(alg.get()[1]).copy()
        # oracle run, same seeds
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
        for k, m in enumerate(steps):
            o.build(1, problems.build_size(case, m))
            if case['alg'] in ('beta', 'svi', 'bpsvi'):
                ow, _, oi = o.get()
            else:
                ow, oi = o.wts, o.idcs
            same(oi, hist_i[k], nm+' idcs step %d' % m)
            same(ow, hist_w[k], nm+' wts step %d' % m)
        if case['alg'] == 'bpsvi':
            same(o.get()[1], ref_pts, nm+' pseudo-points')
            g[nm+'_pts'] = ref_pts
        if nm == 'c1_zellner_gaussian':     # the restated setup reproduces the unmodified driver's selections
            same(hist_i[-1][:len(problems.C1_DRIVER_FIRST_ROWS)], problems.C1_DRIVER_FIRST_ROWS, 'C1 driver fingerprint')
        g[nm+'_idcs'] = np.array(hist_i[-1]); g[nm+'_wts'] = np.array(hist_w[-1])
        g[nm+'_sizes'] = np.array([len(h) for h in hist_i])
        g[nm+'_first_idcs'] = np.array([h[-1] if len(h) else -1 for h in hist_i])
        g[nm+'_sumw'] = np.array([h.sum() for h in hist_w])
        print('G3 %s ok (%.1fs) idcs %s sumw %.9f' % (nm, time.time()-t0, [int(v) for v in hist_i[-1][:12]], hist_w[-1].sum()))


if __name__ == '__main__':
    main()
