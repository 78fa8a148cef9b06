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


def _newton_mode(Zw, ww, mu0, tol=1e-13, maxit=200):
    """mode of the (strictly concave) weighted log-joint by damped Newton steps: a few D x D Cholesky solves instead of
    scipy's BFGS iterations on a dense D x D inverse-Hessian estimate (tens of ms at D = 128, every optimiser step)"""
    th = np.array(mu0, dtype=np.float64)
    f = log_joint(Zw, th, ww)
    for _ in range(maxit):
        g = grad_th_log_joint(Zw, th, ww)
        H = -hess_th_log_joint(Zw, th, ww)
        step = sl.cho_solve(sl.cho_factor(H, lower=True, check_finite=False), g, check_finite=False)
        t = 1.
        while True:
            th_new = th + t*step
            f_new = log_joint(Zw, th_new, ww)
            if f_new >= f - 1e-15*abs(f) or t < 1e-8:
                break
            t *= .5
        done = np.abs(th_new-th).max() <= tol*(1.+np.abs(th_new).max())
        th, f = th_new, f_new
        if done:
            break
    return th


def get_laplace(wts, Z, mu0, diag=False, method='bfgs'):
    """N(mu, L L^T) Laplace approximation of the weighted posterior; returns (mu, L, Linv^T-factor)
    like bayesiancoresets/util/opt.py:10-33.  method='bfgs': the reference's optimiser (scipy BFGS with analytic
    gradient, up to 10 restarts from a jittered start); method='newton': the same mode by damped Newton steps."""
    keep = wts > 0
    Zw, ww = Z[keep, :], wts[keep]
    if method == 'newton':
        mu = _newton_mode(Zw, ww, mu0)
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

    method='newton' additionally warm-starts each mode search at the previous mode (consecutive optimiser steps move the
    weights a little).  prefetch=True draws the NEXT call's S x D standard normals on a helper thread while the caller is
    busy (the draw does not depend on the weights); the global numpy stream is consumed in exactly the same order and
    amounts as without it, provided nothing else draws from it between two sampler calls (true for full-data builds)."""
    mu0 = np.zeros(D) if mu0 is None else mu0
    state = {'mu': None, 'fut': None, 'shape': None}
    pool = None
    if prefetch:
        from concurrent.futures import ThreadPoolExecutor
        pool = ThreadPoolExecutor(max_workers=1)

    def normals(S, d):
        if pool is None:
            return np.random.randn(S, d)
        if state['fut'] is not None and state['shape'] == (S, d):
            out = state['fut'].result()
        else:
            if state['fut'] is not None:
                state['fut'].result()       # a draw of another shape was in flight: it has consumed the stream; keep order
            out = np.random.randn(S, d)
        state['fut'], state['shape'] = pool.submit(np.random.randn, S, d), (S, d)
        return out

    def sampler(S, wts, pts):
        if pts.shape[0] == 0:
            wts = np.zeros(1)
            pts = np.zeros((1, D))
        start = state['mu'] if (method == 'newton' and state['mu'] is not None) else mu0
        mu, LSig, _ = get_laplace(wts, pts, start, method=method)
        state['mu'] = mu
        return mu + normals(S, mu.shape[0]).dot(LSig.T)

    def drain():
        """wait for the draw in flight (call before re-seeding numpy's global stream)"""
        if state['fut'] is not None:
            state['fut'].result()
            state['fut'] = None
    sampler.drain = drain
    return sampler
