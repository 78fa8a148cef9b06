"""Device-side Laplace sampler (csrc/bc_sampler.cu) against the host one (examples/common/model_lr.py), same numpy stream."""
import numpy as np
import pytest


pytestmark = pytest.mark.gpu


def _problem(M, D, seed):
    r = np.random.RandomState(seed)
    Z = r.randn(M, D)
    w = r.rand(M)*50
    w[1::5] = 0.         # rows the optimiser has clamped to zero do not contribute
    return Z, w


@pytest.mark.parametrize('M,D,S', [(1, 3, 7), (40, 10, 64), (200, 128, 1024), (300, 51, 100)])
@pytest.mark.parametrize('method', ['device', 'hybrid'])
def test_device_sampler_matches_host(M, D, S, method):
    import torch
    import model_lr
    Z, w = _problem(M, D, 3)
    np.random.seed(11)
    host = model_lr.make_laplace_sampler(D, method='newton')
    th_h = [host(S, w, Z), host(S, w*1.1, Z)]
    np.random.seed(11)
    dev = model_lr.make_laplace_sampler(D, method=method)
    th_d = [dev(S, w, Z).cpu().numpy(), dev(S, w*1.1, Z).cpu().numpy()]
    assert dev.status()[0] == 0
    for a, b in zip(th_h, th_d):
        assert a.shape == b.shape
        np.testing.assert_allclose(b, a, rtol=0, atol=1e-9*max(1., np.abs(a).max()))
    torch.cuda.synchronize()


def test_device_sampler_empty_coreset_is_prior():
    import model_lr
    D, S = 6, 50
    np.random.seed(5)
    R = np.random.randn(S, D)
    np.random.seed(5)
    dev = model_lr.make_laplace_sampler(D, method='device')
    th = dev(S, np.zeros(0), np.zeros((0, D))).cpu().numpy()
    np.testing.assert_array_equal(th, R)


def test_build_with_device_sampler_selects_like_host():
    """a short beta-coreset build: same first selections and weights to 1e-6 with either sampler"""
    import bayesiancoresets as bc
    import model_lr
    r = np.random.RandomState(0)
    N, D, S = 3000, 12, 128
    Z = r.randn(N, D)
    out = []
    for method in ('newton', 'hybrid', 'device'):
        np.random.seed(2)
        prj = bc.BetaBlackBoxProjector(model_lr.make_laplace_sampler(D, method=method), S, model_lr.beta_likelihood, model_lr.log_likelihood, None)
        alg = bc.BetaCoreset(Z, prj, opt_itrs=10, step_sched=lambda i: 1./(1.+i), beta=0.3, learn_beta=False)
        for m in range(1, 6):
            alg.build(1, m)
        out.append((alg.idcs.copy(), alg.wts.copy()))
    for o in out[1:]:
        np.testing.assert_array_equal(out[0][0], o[0])
        np.testing.assert_allclose(o[1], out[0][1], rtol=1e-6)


@pytest.mark.parametrize('model', ['gaussian', 'neurlin'])
def test_conjugate_samplers_device_form(model):
    """make_conjugate_sampler(device=True): the last line of the reference's samplers, mu + randn(S, D).dot(L.T), formed on
    the GPU from the same numpy stream"""
    import gaussian, model_neurlinr
    r = np.random.RandomState(2)
    D, M, S = 9, 30, 77
    A = r.randn(D, D)
    Sig0inv = np.eye(D) + 0.1*A.dot(A.T)
    mu0 = r.randn(D)
    w = r.rand(M)*3
    if model == 'gaussian':
        B = r.randn(D, D)
        Siginv = np.linalg.inv(np.eye(D)*4 + B.dot(B.T))
        pts = r.randn(M, D)
        make = lambda dev: gaussian.make_conjugate_sampler(mu0, Sig0inv, Siginv, device=dev)
    else:
        pts = np.hstack((np.maximum(r.randn(M, D), 0), r.randn(M, 1)))
        make = lambda dev: model_neurlinr.make_conjugate_sampler(mu0, Sig0inv, 0.6, device=dev)
    np.random.seed(21)
    host = [make(False)(S, w, pts), make(False)(S, np.zeros(0), np.zeros((0, pts.shape[1])))]
    np.random.seed(21)
    dev = [make(True)(S, w, pts).cpu().numpy(), make(True)(S, np.zeros(0), np.zeros((0, pts.shape[1]))).cpu().numpy()]
    for a, b in zip(host, dev):
        np.testing.assert_allclose(b, a, rtol=0, atol=1e-12*max(1., np.abs(a).max()))


def test_affine_samples_rejects_a_full_matrix():
    from bayesiancoresets.util.samplers import affine_samples
    with pytest.raises(ValueError):
        affine_samples(np.zeros(3), np.ones((3, 3)), np.zeros((4, 3)))


@pytest.mark.parametrize('name', ['gauss_beta_sub', 'nl_beta_small', 'gauss_svi_full'])
def test_golden_builds_with_device_samplers_and_look_ahead_rng(name):
    """the package's conjugate samplers in their optimiser-loop form (host factors the D x D precision, samples formed on the
    device by a triangular solve, normals AND the sub-sample indices drawn one call ahead on a helper thread): the builds
    must reproduce the golden coresets the unmodified reference produced with its plain host sampler on the same seeds --
    i.e. the global numpy stream is consumed in exactly the reference's order (util/rng.py) and the samples agree"""
    import os
    import bayesiancoresets as bc
    import gaussian, model_neurlinr
    import problems
    from bayesiancoresets.util import rng
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'g3_coresets.npz'))
    case = [c for c in problems.coreset_cases(False) if c['name'] == name][0]
    prob = case['make']()
    rng.drain()
    problems.reseed(case)
    if prob['model'] == 'gauss':
        sampler = gaussian.make_conjugate_sampler(prob['prior']['mu0'], prob['prior']['Sig0inv'], prob['params']['Siginv'], device=True, prefetch=True)
        bl, ll = gaussian.gaussian_beta_likelihood.bind(**prob['params']), gaussian.gaussian_loglikelihood.bind(**prob['params'])
    else:
        sampler = model_neurlinr.make_conjugate_sampler(prob['prior']['mu0'], prob['prior']['Sig0inv'], prob['params']['sigsq'], device=True, prefetch=True)
        bl, ll = model_neurlinr.neurlinr_beta_likelihood.bind(**prob['params']), model_neurlinr.neurlinr_loglikelihood.bind(**prob['params'])
    if case['alg'] == 'beta':
        alg = bc.BetaCoreset(prob['data'], bc.BetaBlackBoxProjector(sampler, case['S'], bl, ll, None), n_subsample_select=case['n_sel'],
                             n_subsample_opt=case['n_opt'], opt_itrs=case['opt_itrs'], step_sched=case['sched'], beta=case['beta'], learn_beta=False)
    else:
        alg = bc.SparseVICoreset(prob['data'], bc.BlackBoxProjector(sampler, case['S'], ll, None), n_subsample_select=case['n_sel'],
                                 n_subsample_opt=case['n_opt'], opt_itrs=case['opt_itrs'], step_sched=case['sched'])
    for m in range(1, case['M']+1):
        alg.build(1, m)
    sampler.drain()
    w, _, i = alg.get()[:3]
    np.testing.assert_array_equal(i, g[name+'_idcs'])
    np.testing.assert_allclose(w, g[name+'_wts'], rtol=1e-6, atol=1e-9)
    ahead = rng.active()
    assert ahead is not None and ahead.hits > case['M']*case['opt_itrs']//2       # the look-ahead did serve most draws


@pytest.mark.parametrize('M,D', [(1, 3), (7, 10), (40, 128), (128, 128), (150, 100)])
def test_laplace_factor_kernel_matches_host(M, D):
    """bc_laplace_logistic_factor (dual-space Newton steps while M <= D, Cholesky FACTOR out) against the host get_laplace"""
    import torch
    import model_lr
    from bayesiancoresets import _native as nv
    from bayesiancoresets._device import Engine, DeviceRows, ptr, stream_ptr
    eng = Engine.get()
    Z, w = _problem(M, D, 11)
    mu, _, C = model_lr.get_laplace(w, Z, np.zeros(D), method='newton', want_inverse=False)
    core = DeviceRows(eng, Z)
    d_mu, d_C = eng.zeros(D), eng.empty(D, D)
    info = torch.zeros(2, dtype=torch.int32, device=eng.device)
    d_w = eng.upload(w)
    nv.call('bc_laplace_logistic_factor', eng.ctx('sampler'), ptr(core.t), core.ld, ptr(d_w), M, D, ptr(d_mu), ptr(d_C), 200, 1e-13,
            ptr(info), stream_ptr())
    assert int(info.cpu()[0]) == 0
    np.testing.assert_allclose(d_mu.cpu().numpy(), mu, rtol=0, atol=1e-10*max(1., np.abs(mu).max()))
    np.testing.assert_allclose(d_C.cpu().numpy(), C, rtol=0, atol=1e-9*np.abs(C).max())


@pytest.mark.parametrize('model', ['gaussian', 'neurlin'])
def test_conjugate_factor_kernel_matches_host(model):
    """bc_conjugate_factor: precision factor and the reference's mean C^-1 C^-T v (gaussian.py:28-32, model_neurlinr.py:115-122)"""
    import torch
    import gaussian, model_neurlinr
    from bayesiancoresets import _native as nv
    from bayesiancoresets._device import Engine, DeviceRows, ptr, stream_ptr
    eng = Engine.get()
    r = np.random.RandomState(4)
    D, M = 23, 57
    A = r.randn(D, D)
    Sig0inv = np.eye(D) + 0.1*A.dot(A.T)
    mu0 = r.randn(D)
    w = r.rand(M)*3
    w[::7] = 0.
    if model == 'gaussian':
        B = r.randn(D, D)
        Siginv = np.linalg.inv(np.eye(D)*4 + B.dot(B.T))
        pts = r.randn(M, D)
        mu, _, C = gaussian.weighted_post(mu0, Sig0inv, Siginv, pts, w)
        args = (nv.MODEL_GAUSSIAN, Sig0inv, Siginv, 1.0)
    else:
        pts = np.hstack((np.maximum(r.randn(M, D), 0), r.randn(M, 1)))
        mu, _, C = model_neurlinr.weighted_post(mu0, Sig0inv, 0.6, pts, w)
        args = (nv.MODEL_NEURLIN, Sig0inv, None, 0.6)
    core = DeviceRows(eng, pts)
    d_mu, d_C = eng.empty(D), eng.empty(D, D)
    info = torch.zeros(2, dtype=torch.int32, device=eng.device)
    A1 = eng.upload(args[2]) if args[2] is not None else None
    d_w, A0, v0 = eng.upload(w), eng.upload(args[1]), eng.upload(Sig0inv.dot(mu0))      # named: the buffers must outlive the call
    nv.call('bc_conjugate_factor', eng.ctx('sampler'), args[0], ptr(core.t), core.ld, ptr(d_w), M, D, ptr(A0), ptr(A1),
            ptr(v0), args[3], ptr(d_mu), ptr(d_C), ptr(info), stream_ptr())
    assert int(info.cpu()[0]) == 0
    np.testing.assert_allclose(d_C.cpu().numpy(), C, rtol=0, atol=1e-12*np.abs(C).max())
    np.testing.assert_allclose(d_mu.cpu().numpy(), mu, rtol=0, atol=1e-11*max(1., np.abs(mu).max()))


@pytest.mark.parametrize('D', [5, 37, 100])
def test_diagonal_precision_shortcuts_change_no_bit(D):
    """Diagonal prior and noise precisions (the reference's Gaussian example: Sig0 = I, Sig = 500 I) take shortcuts in three
    kernels -- bc_conjugate_factor skips the pivoting, bc_sample_solve_hinted the substitution, the sample preparation takes
    one term per dot product.  Each is checked against the general code path, forced by one denormal off-diagonal entry that
    cannot change any result: bit-identical output."""
    import torch
    import gaussian
    from bayesiancoresets import _native as nv
    from bayesiancoresets._device import Engine, DeviceRows, ptr, stream_ptr
    eng = Engine.get()
    ctx = eng.ctx('sampler')
    r = np.random.RandomState(D)
    M, S = 9, 77
    Sig0inv = np.diag(0.5 + r.rand(D))
    Siginv = np.diag(1./(100. + 400.*r.rand(D)))
    mu0, w, pts = r.randn(D), 3.*r.rand(M), 5.*r.randn(M, D)
    core, d_w, v0 = DeviceRows(eng, pts), eng.upload(w), eng.upload(Sig0inv.dot(mu0))
    R = eng.upload(r.randn(S, D))
    res = []
    for dense in (False, True):
        A0h, A1h = Sig0inv.copy(), Siginv.copy()
        if dense and D > 1:
            A0h[1, 0] = A0h[0, 1] = 5e-324
            A1h[D-1, 0] = A1h[0, D-1] = 5e-324
        A0, A1 = eng.upload(A0h), eng.upload(A1h)
        d_mu, d_C, th, th2 = eng.empty(D), eng.empty(D, D), eng.empty(S, D), eng.empty(S, D)
        info = torch.zeros(2, dtype=torch.int32, device=eng.device)
        nv.call('bc_conjugate_factor', ctx, nv.MODEL_GAUSSIAN, ptr(core.t), core.ld, ptr(d_w), M, D, ptr(A0), ptr(A1), ptr(v0), 1.0,
                ptr(d_mu), ptr(d_C), ptr(info), stream_ptr())
        nv.call('bc_sample_solve_hinted', ctx, ptr(d_mu), ptr(d_C), ptr(R), S, D, ptr(th), D, ptr(info), stream_ptr())
        nv.call('bc_sample_solve', ctx, ptr(d_mu), ptr(d_C), ptr(R), S, D, ptr(th2), D, stream_ptr())
        inf = info.cpu().numpy()
        assert inf[0] == 0 and inf[1] == (0 if dense else 1)
        np.testing.assert_array_equal(th.cpu().numpy(), th2.cpu().numpy())      # hinted == plain on the same factor
        # the prepared samples / potential under this Siginv
        pot = gaussian.gaussian_beta_likelihood.bind(Siginv=A1h, logdetSig=float(-np.log(np.diag(Siginv)).sum()))
        V = pot(pts, th.cpu().numpy(), 0.3)
        res.append((d_mu.cpu().numpy(), d_C.cpu().numpy(), th.cpu().numpy(), V))
    (mu_a, C_a, th_a, V_a), (mu_b, C_b, th_b, V_b) = res
    C_b[1, 0], C_b[D-1, 0] = C_a[1, 0], C_a[D-1, 0]        # the planted denormals' own images
    np.testing.assert_array_equal(C_a, C_b)
    np.testing.assert_array_equal(mu_a, mu_b)
    np.testing.assert_array_equal(th_a, th_b)
    np.testing.assert_array_equal(V_a, V_b)
    mu_h, _, C_h = gaussian.weighted_post(mu0, Sig0inv, Siginv, pts, w)
    np.testing.assert_allclose(C_a, C_h, rtol=0, atol=1e-14*np.abs(C_h).max())
    np.testing.assert_allclose(mu_a, mu_h, rtol=0, atol=1e-12*max(1., np.abs(mu_h).max()))


@pytest.mark.parametrize('model', ['gaussian_sub', 'logistic_full'])
def test_graph_replayed_optimiser_loop_changes_no_bit(monkeypatch, model):
    """coreset/_greedy.py::_optimize_graph_steps: the device side of an optimiser step captured into CUDA graphs and replayed
    (step size / bias corrections from a device-resident schedule, normals and sub-sample indices through fixed pinned
    buffers) builds bit for bit the coreset the kernel-by-kernel loop builds, and consumes the numpy stream alike."""
    import bayesiancoresets as bc
    import model_lr, gaussian
    from bayesiancoresets.coreset import _greedy
    from bayesiancoresets.util import rng
    r = np.random.RandomState(3)
    out = []
    for graph in (True, False):
        monkeypatch.setattr(_greedy, 'STEP_GRAPH', graph)
        monkeypatch.setattr(_greedy, 'STEP_GRAPH_MIN_ITRS', 8)
        rng.drain()
        np.random.seed(7)
        if model == 'gaussian_sub':
            N, D, S = 3000, 20, 64
            X = r.randn(N, D)*3. if graph else X
            Siginv = np.eye(D)/9.
            smp = gaussian.make_conjugate_sampler(np.zeros(D), np.eye(D), Siginv, device=True, prefetch=True)
            bl = gaussian.gaussian_beta_likelihood.bind(Siginv=Siginv, logdetSig=float(D*np.log(9.)))
            ll = gaussian.gaussian_loglikelihood.bind(Siginv=Siginv, logdetSig=float(D*np.log(9.)))
            prj = bc.BetaBlackBoxProjector(smp, S, bl, ll, None)
            alg = bc.BetaCoreset(X, prj, opt_itrs=45, n_subsample_opt=500, n_subsample_select=800, step_sched=lambda i: .1/(1.+i), beta=0.2,
                                 learn_beta=False)
        else:
            N, D, S = 4000, 24, 96
            X = r.randn(N, D) if graph else X
            smp = model_lr.make_laplace_sampler(D, method='hybrid', prefetch=True)
            prj = bc.BetaBlackBoxProjector(smp, S, model_lr.beta_likelihood, model_lr.log_likelihood, None)
            alg = bc.BetaCoreset(X, prj, opt_itrs=33, step_sched=lambda i: 1./(1.+i), beta=0.3, learn_beta=False)
        for m in range(1, 5):
            alg.build(1, m)
        smp.drain()
        w, _, idcs = alg.get()[:3]
        out.append((np.array(w), np.array(idcs), np.random.rand(3)))
    np.testing.assert_array_equal(out[0][1], out[1][1])
    np.testing.assert_array_equal(out[0][0], out[1][0])
    np.testing.assert_array_equal(out[0][2], out[1][2])
    assert len(out[0][1]) >= 3 and (out[0][0] > 0).any()


def test_device_sampler_loop_matches_host_protocol_loop(monkeypatch):
    """the optimiser loop with the sampler's device_step (no host round trip per step) builds the coreset the loop builds
    when the same sampler is called through the reference's host protocol sampler(S, wts, pts)"""
    import bayesiancoresets as bc
    import model_lr
    from bayesiancoresets.coreset import _greedy
    r = np.random.RandomState(0)
    N, D, S = 4000, 24, 96
    Z = r.randn(N, D)
    out = []
    for loop in (True, False):
        monkeypatch.setattr(_greedy, 'DEVICE_SAMPLER_LOOP', loop)
        np.random.seed(2)
        smp = model_lr.make_laplace_sampler(D, method='hybrid', prefetch=True)
        prj = bc.BetaBlackBoxProjector(smp, S, model_lr.beta_likelihood, model_lr.log_likelihood, None)
        alg = bc.BetaCoreset(Z, prj, opt_itrs=12, n_subsample_opt=1500, step_sched=lambda i: 1./(1.+i), beta=0.3, learn_beta=False)
        for m in range(1, 7):
            alg.build(1, m)
        smp.drain()
        out.append((alg.idcs.copy(), alg.wts.copy()))
    np.testing.assert_array_equal(out[0][0], out[1][0])
    np.testing.assert_allclose(out[0][1], out[1][1], rtol=1e-7, atol=1e-10)
