"""Bayesian logistic regression: device potentials + the host-side pieces of the coreset-posterior
sampler.  Drop-in for the hot-path functions of examples/common/model_lr.py.

Data rows are z_n = y_n x_n (model_lr.py:29); with m = -z.theta
    log_likelihood(z, th)        = -log(1 + e^m)                                     (model_lr.py:72-79)
    beta_likelihood(z, th, beta) = -((b+1)/b (1+e^m)^-b - ((1+e^m)^(-b-1) + (1+e^-m)^(-b-1)))   (:81-86)
Both are `DevicePotential`s: callable with the reference signature (returning the (n, S) host
array, evaluated by the CUDA materialise kernel) and recognised by the projectors, which then run
the fused kernels instead of forming the matrix.

The functions below the potentials are the weighted-posterior Laplace approximation the drivers
use as sampler (examples/zellner_logreg/main.py:86-111,139-144; bayesiancoresets/util/opt.py:10-33).
They act on the M coreset points only (M x D, a few hundred rows at most) and stay on the host.
"""
import os
import numpy as np
import scipy.linalg as sl
from scipy.optimize import minimize

from bayesiancoresets.potentials import DevicePotential, DeviceGradient

log_likelihood = DevicePotential('logistic', 'loglik', name='log_likelihood')
beta_likelihood = DevicePotential('logistic', 'betalik', name='beta_likelihood')
grad_z_log_likelihood = DeviceGradient('logistic', name='grad_z_log_likelihood')        # model_lr.py:107-114 (BatchPSVI)


# ----------------------------------------------------------------- host: data utilities --
def gen_synthetic_outliers(N, D, seed=0, flip_rate=0.1, intercept=False):
    """SURVEY 8d north-star generator: X ~ N(0, I), theta* = 1/sqrt(D), y ~ Bernoulli(sigmoid(X theta*)),
    `flip_rate` of the labels flipped, Z = y X.  Returns (Z, X, y, flipped_mask)."""
    r = np.random.RandomState(seed)
    X = r.randn(N, D)
    if intercept:
        X[:, -1] = 1.
    th = np.ones(D)/np.sqrt(D)
    y = np.where(r.rand(N) < 1./(1.+np.exp(-X.dot(th))), 1., -1.)
    flipped = r.rand(N) < flip_rate
    y[flipped] *= -1.
    return y[:, np.newaxis]*X, X, y, flipped


# ----------------------------------------------- host: weighted posterior (M x D, tiny) --
def _sigma_of_margin(Z, th):
    m = -Z.dot(th)
    out = np.ones_like(m)
    ok = m < 100
    e = np.exp(m[ok])
    out[ok] = e/(1.+e)
    return m, out


def log_joint(Z, th, wts):
    """weighted log-likelihood + N(0, I) log-prior at ONE parameter vector (model_lr.py:88-93)."""
    m = -Z.dot(th)
    ll = np.where(m < 100, -np.log1p(np.exp(np.minimum(m, 100.))), -m)
    return wts.dot(ll) - 0.5*th.shape[0]*np.log(2.*np.pi) - 0.5*(th**2).sum()


def grad_th_log_joint(Z, th, wts):
    """gradient of log_joint in theta (model_lr.py:98-105,116-121)."""
    _, s = _sigma_of_margin(Z, th)
    return -th + Z.T.dot(wts*s)


def hess_th_log_joint(Z, th, wts):
    """Hessian of log_joint in theta (model_lr.py:123-137)."""
    m, s = _sigma_of_margin(Z, th)
    c = np.where(m < 100, s*(1.-s), 0.)
    return -np.eye(th.shape[0]) - (Z*(wts*c)[:, np.newaxis]).T.dot(Z)


def _margin_terms(m):
    """sigma(m) and sigma(m)(1 - sigma(m)) with the reference's m < 100 guard (model_lr.py:98-105,123-137)"""
    s = np.ones_like(m)
    ok = m < 100
    e = np.exp(m[ok])
    s[ok] = e/(1.+e)
    return s, np.where(ok, s*(1.-s), 0.)


_MASKS = {}


def _lower_mask(D):
    if D not in _MASKS:
        _MASKS[D] = np.tril(np.ones((D, D)))
    return _MASKS[D]


def _neg_hessian(Zw, ww, c):
    return np.eye(Zw.shape[1]) + (Zw*(ww*c)[:, np.newaxis]).T.dot(Zw)


def _log_joint_from_margins(m, th, ww):
    ll = np.where(m < 100, -np.log1p(np.exp(np.minimum(m, 100.))), -m)
    return ww.dot(ll) - 0.5*th.shape[0]*np.log(2.*np.pi) - 0.5*(th**2).sum()


def _newton_mode(Zw, ww, mu0, tol=1e-13, maxit=200, want_margins=False):
    """mode of the (strictly concave) weighted log-joint by damped Newton steps: a few D x D Cholesky solves instead of
    scipy's BFGS iterations on a dense D x D inverse-Hessian estimate (tens of ms at D = 128, every optimiser step).
    The margins -Z th are computed once per point and shared by value, gradient and Hessian; LAPACK is called through
    scipy's thin f2py wrappers (dpotrf / dpotrs) -- at D = 128 the checked high-level wrappers cost as much as the solves."""
    th = np.array(mu0, dtype=np.float64)
    m = -Zw.dot(th)
    f = _log_joint_from_margins(m, th, ww)
    # Few coreset points (M < D/2, the usual case while a coreset is being built): the negative Hessian I + Z^T diag(d) Z is
    # a rank-M update of the identity, so the Newton step comes from an M x M system (Woodbury),
    #   H^-1 g = g - Z^T r (I + r K r)^-1 r Z g,   r = sqrt(d),  K = Z Z^T (formed once per call),
    # instead of a D x D factorisation per iteration.  Same iteration, same stopping rule; the iterates agree to rounding.
    M, D = Zw.shape
    dual = 0 < 2*M < D
    K = Zw.dot(Zw.T) if dual else None
    for _ in range(maxit):
        s, c = _margin_terms(m)
        g = -th + Zw.T.dot(ww*s)
        if dual:
            r = np.sqrt(ww*c)
            A = K*r[:, np.newaxis]*r[np.newaxis, :]
            A[np.diag_indices_from(A)] += 1.
            La, info = sl.lapack.dpotrf(A, lower=1, overwrite_a=1)
            if info != 0:
                raise np.linalg.LinAlgError('negative Hessian not positive definite (dual dpotrf info %d)' % info)
            v, _ = sl.lapack.dpotrs(La, r*Zw.dot(g), lower=1)
            step = g - Zw.T.dot(r*v)
        else:
            L, info = sl.lapack.dpotrf(_neg_hessian(Zw, ww, c), lower=1, overwrite_a=1)
            if info != 0:
                raise np.linalg.LinAlgError('negative Hessian not positive definite (dpotrf info %d)' % info)
            step, _ = sl.lapack.dpotrs(L, g, lower=1)
        t = 1.
        while True:
            th_new = th + t*step
            m_new = -Zw.dot(th_new)
            f_new = _log_joint_from_margins(m_new, th_new, ww)
            if f_new >= f - 1e-15*abs(f) or t < 1e-8:
                break
            t *= .5
        dm, am = np.abs(th_new-th).max(), np.abs(th_new).max()
        # a full step this small is in the quadratic regime: the point it lands on is within ~|z| dm^2 of the mode
        done = dm <= tol*(1.+am) or (t == 1. and dm <= 1e-7*(1.+am))
        th, f, m = th_new, f_new, m_new
        if done:
            break
    return (th, m) if want_margins else th


def get_laplace(wts, Z, mu0, diag=False, method='bfgs', want_inverse=True):
    """N(mu, L L^T) Laplace approximation of the weighted posterior; returns (mu, L, Linv^T-factor)
    like bayesiancoresets/util/opt.py:10-33.  method='bfgs': the reference's optimiser (scipy BFGS with analytic
    gradient, up to 10 restarts from a jittered start); method='newton': the same mode by damped Newton steps."""
    keep = wts > 0
    Zw, ww = Z[keep, :], wts[keep]
    if method == 'newton':
        mu, m = _newton_mode(Zw, ww, mu0, want_margins=True)
        if not diag:
            # factor and invert through the LAPACK routines directly: L = chol(-Hessian), LSig = L^-1 (both lower)
            L, info = sl.lapack.dpotrf(_neg_hessian(Zw, ww, _margin_terms(m)[1]), lower=1, overwrite_a=1)
            if info != 0:
                raise np.linalg.LinAlgError('negative Hessian not positive definite (dpotrf info %d)' % info)
            mask = _lower_mask(L.shape[0])         # dpotrf / dtrtri leave the other triangle untouched
            LSigInv = L*mask
            if not want_inverse:                   # the caller solves against the factor itself (bc_sample_solve)
                return mu, None, LSigInv
            LSig, info = sl.lapack.dtrtri(LSigInv, lower=1)
            return mu, LSig*mask, LSigInv
    else:
        res = None
        for _ in range(10):
            try:
                res = minimize(lambda mu: -log_joint(Zw, mu, ww), mu0, jac=lambda mu: -grad_th_log_joint(Zw, mu, ww))
                break
            except Exception:
                mu0 = mu0 + np.sqrt((mu0**2).sum())*0.1*np.random.randn(mu0.shape[0])
        if res is None:
            raise RuntimeError('Laplace optimisation failed 10 times')
        mu = res.x
    H = -hess_th_log_joint(Zw, mu, ww)
    if diag:
        LSigInv = np.sqrt(np.diag(H))
        return mu, 1./LSigInv, LSigInv
    LSigInv = np.linalg.cholesky(H)
    LSig = sl.solve_triangular(LSigInv, np.eye(LSigInv.shape[0]), lower=True, overwrite_b=True, check_finite=False)
    return mu, LSig, LSigInv


def make_laplace_sampler(D, mu0=None, method='bfgs', prefetch=False):
    """sampler(S, wts, pts) -> (S, D): the callback the logistic drivers hand to the projector
    (examples/zellner_logreg/main.py:139-144).  Empty coreset -> the N(0, I) prior.

    method='hybrid' / 'device' form the samples on the GPU and return a device tensor (see `_device_laplace_sampler`).
    method='newton' additionally warm-starts each mode search at the previous mode (consecutive optimiser steps move the
    weights a little).  prefetch=True draws the NEXT call's S x D standard normals -- and whatever the coreset class draws
    between two calls, the sub-sample indices of bcores.py:53 -- on a helper thread while the caller is busy (the draws do
    not depend on the weights); the global numpy stream is consumed in exactly the same order and amounts as without it
    (bayesiancoresets/util/rng.py)."""
    mu0 = np.zeros(D) if mu0 is None else mu0
    state = {'mu': None}
    ahead = None
    if prefetch:
        from bayesiancoresets.util import rng
        ahead = rng.activate()

    def normals(S, d, stage=None):
        """S x d standard normals from numpy's global stream; `stage` (optional) post-processes a draw on the thread that made it"""
        if ahead is not None:
            return ahead.randn(S, d, stage)
        r = np.random.randn(S, d)
        return r if stage is None else stage(r)

    def new_call():
        if ahead is not None:
            ahead.begin_cycle()

    def sampler(S, wts, pts):
        new_call()
        if pts.shape[0] == 0:
            wts = np.zeros(1)
            pts = np.zeros((1, D))
        start = state['mu'] if (method == 'newton' and state['mu'] is not None) else mu0
        mu, LSig, _ = get_laplace(wts, pts, start, method=method)
        state['mu'] = mu
        return mu + normals(S, mu.shape[0]).dot(LSig.T)

    def drain():
        """stop the look-ahead and leave numpy's global stream where this sampler's own draws have left it (call before
        re-seeding it, or before drawing from it directly)"""
        if ahead is not None:
            ahead.drain()

    if method in ('device', 'hybrid'):
        sampler = _device_laplace_sampler(D, mu0, normals, host_factor=(method == 'hybrid'), new_call=new_call,
                                          peek=(lambda: ahead.peek_next()) if ahead is not None else (lambda: None))
    sampler.drain = drain
    return sampler


def _device_laplace_sampler(D, mu0, normals, host_factor, new_call=lambda: None, peek=lambda: None):
    """method='device' / 'hybrid': the S x D samples are formed on the GPU and returned as a device tensor (the projector's
    fused path takes it without a host round trip).  The standard normals still come from numpy's global stream, so the
    sampler consumes it exactly like the host one.
      'hybrid': mode and Cholesky factor C of the negative Hessian on the host (`get_laplace(method='newton')`, warm-started;
                the Newton steps come from an M x M system while the coreset is small), the samples mu + C^-1 R^T on the device
                by forward substitution (csrc/bc_sampler.cu::k_sample_solve) -- the host neither inverts the factor (dtrtri is
                the slower LAPACK call of the two) nor multiplies, and the 8 S D-byte sample upload leaves its critical path
                (the normals go up instead, and do not wait for the weights);
      'device': the Newton mode search and the factorisations too (k_laplace_logistic, one CTA; D <= 160).  Same iteration and
                tolerance as `_newton_mode`; the two agree to rounding.  At D = 128 the single-CTA factorisations are
                latency-bound and no faster than the host's LAPACK, so 'hybrid' is the quicker of the two."""
    import torch
    from bayesiancoresets import _native as nv
    from bayesiancoresets._device import Engine, ptr, stream_ptr
    st = {'mu': None, 'L': None, 'info': None, 'pin': [None, None], 'ev': [None, None], 'k': 0, 'mu_host': None, 'ml': None}

    def stage_buf(shape):
        """the next of two pinned staging buffers, free again (the upload that last used it has left it)"""
        k = st['k'] = st['k'] ^ 1
        if st['pin'][k] is None or tuple(st['pin'][k].shape) != tuple(shape):
            st['pin'][k] = torch.empty(*shape, dtype=torch.float64).pin_memory()
        if st['ev'][k] is not None:
            st['ev'][k].synchronize()
        return k, st['pin'][k]

    def stage(r):
        """copy a draw into a pinned staging buffer (runs on the thread that drew it)"""
        k, pin = stage_buf(r.shape)
        pin.numpy()[...] = r
        return k, pin

    def stage_into(shape):
        """... or have the native generator write the draw there directly (bayesiancoresets/util/rng.py)"""
        k, pin = stage_buf(shape)
        return (k, pin), pin.numpy()
    stage.into = stage_into

    def sampler(S, wts, pts):
        new_call()
        eng = Engine.get()
        ctx = eng.ctx('sampler')
        wts = np.asarray(wts, dtype=np.float64)
        pts = np.atleast_2d(np.asarray(pts, dtype=np.float64))
        if st['mu'] is None:
            st['ml'] = eng.empty(D*D + D)                  # [mu | L] in one buffer: one upload in 'hybrid'
            st['mu'], st['L'] = st['ml'][:D], st['ml'][D:].view(D, D)
            st['mu'].copy_(torch.from_numpy(np.asarray(mu0, dtype=np.float64)))
            st['info'] = torch.zeros(2, dtype=torch.int32, device=eng.device)
            st['ml_pin'] = torch.empty(D*D + D, dtype=torch.float64).pin_memory()
        theta = eng.empty(S, D)
        keep = wts > 0
        st['factor'] = False
        if pts.shape[0] == 0 or not keep.any():           # empty coreset: the N(0, I) prior
            st['mu'].zero_()
            st['L'].copy_(torch.eye(D, dtype=torch.float64, device=eng.device))
            st['mu_host'] = None
        elif host_factor:
            if st.get('device_mode'):                      # the optimiser loop has moved the mode on the device: resume from it
                st['mu_host'] = st['mu'].cpu().numpy().copy()
                st['device_mode'] = False
            start = st['mu_host'] if st['mu_host'] is not None else mu0
            solve = D <= 160                               # bc_sample_solve's limit; beyond it: invert on the host as before
            mu, LSig, LSigInv = get_laplace(wts, pts, start, method='newton', want_inverse=not solve)
            st['mu_host'] = mu
            st['factor'] = solve                           # st['L'] holds the Cholesky factor, not its inverse
            if not solve:
                LSigInv = LSig
            if st.get('ml_ev') is not None:
                st['ml_ev'].synchronize()
            buf = st['ml_pin'].numpy()
            buf[:D] = mu
            buf[D:] = LSigInv.ravel()
            st['ml'].copy_(st['ml_pin'], non_blocking=True)
            st['ml_ev'] = torch.cuda.Event()
            st['ml_ev'].record()
        else:
            Zd = eng.upload(pts[keep])
            wd = eng.upload(wts[keep])
            nv.call('bc_laplace_logistic', ctx, ptr(Zd), int(Zd.stride(0)), ptr(wd), int(Zd.shape[0]), D, ptr(st['mu']), ptr(st['L']),
                    200, 1e-13, ptr(st['info']), stream_ptr())
        # the normals last: the next draw starts on the helper thread here, after the Python-heavy part of this call
        k, pin = normals(S, D, stage)
        Rd = pin.to(eng.device, non_blocking=True)
        st['ev'][k] = torch.cuda.Event()
        st['ev'][k].record()
        nv.call('bc_sample_solve' if st['factor'] else 'bc_sample_affine', ctx, ptr(st['mu']), ptr(st['L']), ptr(Rd), S, D, ptr(theta),
                int(theta.stride(0)), stream_ptr())
        return theta

    def device_step(S, w_dev, core):
        """the sampler call of one optimiser step with the weights already on the device (`w_dev`) and the coreset points as
        resident rows (`core`, a DeviceRows): mode (warm-started dual Newton steps), Cholesky factor and samples are formed by
        kernels (bc_laplace_logistic_factor, bc_sample_solve); nothing is read back, the host only queues work and draws the
        normals.  Rows with weight 0 do not contribute, exactly like `keep = wts > 0` on the host."""
        prof = st.get('prof')                  # (tools/step_timeline.py: host seconds spent in each part of this call)
        if prof is not None:
            import time
            tp = [time.perf_counter()]
        graph_prepare()
        eng = Engine.get()
        ctx = eng.ctx('sampler')
        theta = eng.empty(S, D)
        if prof is not None:
            tp.append(time.perf_counter())
        ring = st.get('ring')
        if ring is not None and ring.get('to_free') is not None:
            # the ring buffer the previous call's solve read is released HERE, at the head of this call's kernels: the upload
            # that waits for it (four uploads later, the host runs that far ahead) then runs beside the 140 us factor kernel.
            # Released right behind the solve, it ran beside the sample-preparation kernels and stretched them.
            jf, ring['to_free'] = ring['to_free'], None
            ring['free'][jf] = torch.cuda.Event()
            ring['free'][jf].record(torch.cuda.current_stream(eng.device))
        nv.call('bc_laplace_logistic_factor', ctx, ptr(core.t), core.ld, ptr(w_dev), core.n_local, D, ptr(st['mu']), ptr(st['L']), 200, 1e-13,
                ptr(st['info']), stream_ptr())
        if prof is not None:
            tp.append(time.perf_counter())
        tok = normals(S, D, stage)             # after the factor kernel is queued: the GPU works while the host waits for the draw
        k, pin = tok
        if prof is not None:
            tp.append(time.perf_counter())
            prof.append(tuple(b - a for a, b in zip(tp[:-1], tp[1:])))
        # The 8 S D bytes of normals do not depend on anything the device computes, and a single 1 MB pinned copy takes 90-150 us
        # to cross PCIe on the B200 boxes (tools/step_timeline.py): it must not sit in the serial chain of the step.  It goes up
        # on a copy stream of its own, and it goes up a STEP EARLY: the look-ahead generator (bayesiancoresets/util/rng.py) has
        # already drawn the next call's normals into its other pinned staging buffer, so this call uploads them behind its own
        # solve -- beside the data pass that follows -- and the next call finds them on the device.  Nothing is consumed from the
        # random stream early; if the next call is handed another draw (the speculation was rewound) it uploads that one itself.
        # Four device buffers in turn; one is reused once the solve that read it has run.
        if os.environ.get('BC_NORMALS_INLINE') == '1':    # (A/B switch: the upload in stream order, between the factor kernel and the solve)
            Rd = pin.to(eng.device, non_blocking=True)
            graph_launched(k)
            nv.call('bc_sample_solve', ctx, ptr(st['mu']), ptr(st['L']), ptr(Rd), S, D, ptr(theta), int(theta.stride(0)), stream_ptr())
            return theta
        ring = st.get('ring')
        if ring is None or tuple(ring['buf'][0].shape) != (S, D):
            ring = st['ring'] = {'buf': [eng.empty(S, D) for _ in range(4)], 'free': [None]*4, 'i': 0, 'stream': torch.cuda.Stream(device=eng.device),
                                 'pre': None, 'to_free': None}
        cur = torch.cuda.current_stream(eng.device)
        side = ring['stream']

        def upload(token):
            """enqueue the copy of a pinned draw into the next ring buffer on the copy stream -> (ring slot, its completion event)"""
            kk, src = token
            jj = ring['i'] % 4
            ring['i'] += 1
            if ring['free'][jj] is not None:
                side.wait_event(ring['free'][jj])
            with torch.cuda.stream(side):
                ring['buf'][jj].copy_(src, non_blocking=True)
                done = torch.cuda.Event()
                done.record(side)
            st['ev'][kk] = done                # the pinned staging buffer is free again once this copy has run
            return jj, done

        pre, ring['pre'] = ring['pre'], None
        if pre is not None and pre[0] is tok:
            j, ready = pre[1], pre[2]          # uploaded by the previous call
        else:
            j, ready = upload(tok)
        cur.wait_event(ready)
        Rd = ring['buf'][j]
        nv.call('bc_sample_solve', ctx, ptr(st['mu']), ptr(st['L']), ptr(Rd), S, D, ptr(theta), int(theta.stride(0)), stream_ptr())
        ring['to_free'] = j                    # (recorded by the next call; a buffer never released is simply not reused)
        if os.environ.get('BC_NORMALS_PREFETCH', '1') != '0':
            nxt = peek()
            if isinstance(nxt, tuple) and len(nxt) == 2 and isinstance(nxt[1], torch.Tensor) and tuple(nxt[1].shape) == (S, D):
                ring['pre'] = (nxt,) + upload(nxt)
        return theta

    # The same call in parts, so that the device work of a step can be captured into a CUDA graph and replayed
    # (bayesiancoresets/coreset/_greedy.py): only graph_enqueue touches the stream, and it names nothing but fixed buffers.
    def graph_prepare():
        new_call()
        eng = Engine.get()
        if st['mu'] is None:
            st['ml'] = eng.empty(D*D + D)
            st['mu'], st['L'] = st['ml'][:D], st['ml'][D:].view(D, D)
            st['mu'].copy_(torch.from_numpy(np.asarray(mu0, dtype=np.float64)))
            st['info'] = torch.zeros(2, dtype=torch.int32, device=eng.device)
            st['ml_pin'] = torch.empty(D*D + D, dtype=torch.float64).pin_memory()
        if st.get('mu_host') is not None:              # a host-side call came in between: warm-start from its mode
            st['mu'].copy_(torch.from_numpy(st['mu_host']))
            st['mu_host'] = None
        st['device_mode'] = True

    def graph_normals(S):
        """host part: this step's S x D normals, drawn in the reference's order, are in pinned staging buffer k on return"""
        graph_prepare()
        k, _ = normals(S, D, stage)
        return k

    def graph_enqueue(S, w_dev, core, k):
        eng = Engine.get()
        ctx = eng.ctx('sampler')
        theta = eng.empty(S, D)
        nv.call('bc_laplace_logistic_factor', ctx, ptr(core.t), core.ld, ptr(w_dev), core.n_local, D, ptr(st['mu']), ptr(st['L']), 200, 1e-13,
                ptr(st['info']), stream_ptr())
        Rd = st['pin'][k].to(eng.device, non_blocking=True)
        nv.call('bc_sample_solve', ctx, ptr(st['mu']), ptr(st['L']), ptr(Rd), S, D, ptr(theta), int(theta.stride(0)), stream_ptr())
        return theta

    def graph_launched(k):
        st['ev'][k] = torch.cuda.Event()
        st['ev'][k].record()

    def status():
        """(status, newton steps) of the last device mode search: 0 = converged"""
        return tuple(int(v) for v in st['info'].cpu()) if st['info'] is not None else (0, 0)
    sampler.status = status
    sampler.state = st
    sampler.device_step = device_step
    sampler.graph_normals, sampler.graph_enqueue, sampler.graph_launched = graph_normals, graph_enqueue, graph_launched
    sampler.supports_device_step = lambda: D <= 160
    return sampler
