"""Oracle: per-datapoint potentials f(x_n, theta_s) -> (N, S) fp64 arrays.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Each function restates one
reference function; the floating-point expression order is the reference's so
results match bit for bit on the same numpy.
"""
import numpy as np


# ---------------------------------------------------------------- logistic --
def lr_margin(z, th):
    """m = -Z Theta^T  (examples/common/model_lr.py:73-75, :82-84)."""
    return -np.atleast_2d(z).dot(np.atleast_2d(th).T)


def lr_loglik(z, th):
    """examples/common/model_lr.py:72-79: -log1p(e^m) for m < 100, else -m."""
    m = lr_margin(z, th)
    small = m < 100
    big = np.logical_not(small)
    m[small] = -np.log1p(np.exp(m[small]))
    m[big] = -m[big]
    return m


def lr_betalik(z, th, beta):
    """examples/common/model_lr.py:81-86 (no overflow guard: e^m -> inf gives 0/1)."""
    m = lr_margin(z, th)
    with np.errstate(over='ignore'):
        em = np.exp(m)
        enm = np.exp(-m)
        m = -(((beta+1.)/beta)*(1+em)**(-beta) - ((1+em)**(-beta-1.) + (1+enm)**(-beta-1.)))
    return m


# ---------------------------------------------------------------- gaussian --
def _gauss_quad_terms(x, th, Siginv):
    """The three pieces of (x-th)^T Siginv (x-th) as the reference forms them
    (examples/common/gaussian.py:10-12, :38-40, :50-52)."""
    x = np.atleast_2d(x)
    th = np.atleast_2d(th)
    xSx = (x*(x.dot(Siginv))).sum(axis=1)
    tSt = (th*(th.dot(Siginv))).sum(axis=1)
    xSt = x.dot(Siginv.dot(th.T))
    return x, th, xSx, tSt, xSt


def lr_grad_z_loglik(z, th):
    """model_lr.py:107-114 grad_z_log_likelihood: sigma(m)[:, :, None] * th[None] -> (n, S, D)"""
    z = np.atleast_2d(z)
    th = np.atleast_2d(th)
    m = -z.dot(th.T)
    idcs = m < 100
    m[idcs] = np.exp(m[idcs])/(1.+np.exp(m[idcs]))
    m[np.logical_not(idcs)] = 1.
    return m[:, :, np.newaxis]*th[np.newaxis, :, :]


def gauss_grad_x_loglik(x, th, Siginv):
    """gaussian.py:17-20 gaussian_grad_x_loglikelihood -> (n, S, d)"""
    x = np.atleast_2d(x)
    th = np.atleast_2d(th)
    return th.dot(Siginv)[np.newaxis, :, :] - x.dot(Siginv)[:, np.newaxis, :]


def gauss_loglik(x, th, Siginv, logdetSig):
    """examples/common/gaussian.py:7-15 (the reference's debug print is dropped)."""
    x, th, xSx, tSt, xSt = _gauss_quad_terms(x, th, Siginv)
    return -x.shape[1]/2*np.log(2*np.pi) - 1./2.*logdetSig - 1./2.*(xSx[:, np.newaxis] + tSt - 2*xSt)


def gauss_betalik(x, th, beta, Siginv, logdetSig):
    """examples/common/gaussian.py:34-44.  `cnst` (:41) is computed but unused
    by the reference, so it does not appear in the value."""
    x, th, xSx, tSt, xSt = _gauss_quad_terms(x, th, Siginv)
    d = float(x.shape[1])
    t1 = (1./beta)*np.exp(-.5*beta*(xSx[:, np.newaxis] + tSt - 2*xSt))
    t2 = (1+beta)**(-.5*d-1)
    return t1 - t2


def gauss_betagrad(x, th, beta, Siginv, logdetSig):
    """examples/common/gaussian.py:46-62: d/dbeta of the beta-likelihood."""
    x, th, xSx, tSt, xSt = _gauss_quad_terms(x, th, Siginv)
    d = float(x.shape[1])
    logcnst = np.log((2*np.pi)**(-.5*d)*(np.exp(logdetSig)**(-.5)))
    q = xSx[:, np.newaxis] + tSt - 2*xSt
    gaussq = np.exp(-.5*beta*q)
    t11 = (1./beta)*gaussq
    t12 = (1+beta)**(-.5*d-1.)
    t1 = logcnst*(t11-t12)
    t2 = 1./(beta)**2*gaussq
    t3 = 1./(2.*beta)*q*gaussq
    t4 = (1+beta)**(-.5*d-1.)*np.log(1.+beta)
    return t1 - t2 - t3 - t4


def gauss_weighted_post(th0, Sig0inv, Siginv, x, w):
    """examples/common/gaussian.py:28-32 (host-side conjugate sampler parameters)."""
    import scipy.linalg as sl
    LinvT = np.linalg.cholesky(Sig0inv + w.sum()*Siginv)
    L = sl.solve_triangular(LinvT, np.eye(LinvT.shape[0]), lower=True, overwrite_b=True, check_finite=False)
    mu = np.dot(L.dot(L.T), np.dot(Sig0inv, th0) + np.dot(Siginv, (w[:, np.newaxis]*x).sum(axis=0)))
    return mu, L, LinvT


# ----------------------------------------------------------- neural-linear --
def nl_loglik(z, th, sigsq):
    """examples/common/model_neurlinr.py:90-97.  z = [phi(x), y]."""
    z = np.atleast_2d(z)
    x, y = z[:, :-1], z[:, -1]
    u = x.dot(np.atleast_2d(th).T)
    return -1./2.*np.log(2.*np.pi*sigsq) - 1./(2.*sigsq)*(y[:, np.newaxis]**2 - 2*u*y[:, np.newaxis] + u**2)


def nl_betalik(z, th, beta, sigsq):
    """examples/common/model_neurlinr.py:102-110."""
    z = np.atleast_2d(z)
    x, y = z[:, :-1], z[:, -1]
    u = x.dot(np.atleast_2d(th).T)
    return 1./(2*np.pi*sigsq)**(beta/2.)*(-(beta+1.)/beta*np.exp(-beta/(2.*sigsq)*(y[:, np.newaxis]**2 - 2*u*y[:, np.newaxis] + u**2))
                                          + 1./np.sqrt(1.+beta))


def nl_weighted_post(th0, Sig0inv, sigsq, z, w):
    """examples/common/model_neurlinr.py:115-122 (host-side conjugate sampler parameters)."""
    import scipy.linalg as sl
    z = np.atleast_2d(z)
    X, Y = z[:, :-1], z[:, -1]
    LinvT = np.linalg.cholesky(Sig0inv + (w[:, np.newaxis]*X).T.dot(X)/sigsq)
    L = sl.solve_triangular(LinvT, np.eye(LinvT.shape[0]), lower=True, overwrite_b=True, check_finite=False)
    mu = np.dot(L.dot(L.T), np.dot(Sig0inv, th0) + (w[:, np.newaxis]*Y[:, np.newaxis]*X).sum(axis=0)/sigsq)
    return mu, L, LinvT


# --------------------------------------------- logistic Laplace (host sampler) --
def lr_log_joint(z, th, wts):
    """examples/common/model_lr.py:88-93: weighted log-likelihood + N(0,I) log-prior."""
    th2 = np.atleast_2d(th)
    prior = -0.5*th2.shape[1]*np.log(2.*np.pi) - 0.5*(th2**2).sum(axis=1)
    return (wts[:, np.newaxis]*lr_loglik(z, th)).sum(axis=0) + prior


def _lr_sigma(z, th):
    m = lr_margin(z, th)
    small = m < 100
    m[small] = np.exp(m[small])/(1.+np.exp(m[small]))
    m[np.logical_not(small)] = 1.
    return m


def lr_grad_log_joint(z, th, wts):
    """examples/common/model_lr.py:98-105,116-121."""
    z = np.atleast_2d(z)
    g = _lr_sigma(z, th)[:, :, np.newaxis]*z[:, np.newaxis, :]
    return -np.atleast_2d(th) + (wts[:, np.newaxis, np.newaxis]*g).sum(axis=0)


def lr_hess_log_joint(z, th, wts):
    """examples/common/model_lr.py:123-137."""
    z = np.atleast_2d(z)
    th2 = np.atleast_2d(th)
    m = lr_margin(z, th)
    small = m < 100
    m[small] = np.exp(m[small])/(1.+np.exp(m[small]))**2
    m[np.logical_not(small)] = 0.
    H = -m[:, :, np.newaxis, np.newaxis]*z[:, np.newaxis, :, np.newaxis]*z[:, np.newaxis, np.newaxis, :]
    return np.tile(-np.eye(th2.shape[1]), (th2.shape[0], 1, 1)) + (wts[:, np.newaxis, np.newaxis, np.newaxis]*H).sum(axis=0)


def lr_laplace(wts, Z, mu0):
    """bayesiancoresets/util/opt.py:10-33 (full-covariance branch; the retry loop
    re-draws mu0 with np.random on failure exactly like the reference)."""
    import scipy.linalg as sl
    from scipy.optimize import minimize
    trials = 10
    Zw = Z[wts > 0, :]
    ww = wts[wts > 0]
    while True:
        try:
            res = minimize(lambda mu: -lr_log_joint(Zw, mu, ww)[0], mu0,
                           jac=lambda mu: -lr_grad_log_joint(Zw, mu, ww)[0, :])
        except Exception:
            mu0 = mu0.copy()
            mu0 += np.sqrt((mu0**2).sum())*0.1*np.random.randn(mu0.shape[0])
            trials -= 1
            if trials <= 0:
                break
            continue
        break
    mu = res.x
    LinvT = np.linalg.cholesky(-lr_hess_log_joint(Zw, mu, ww)[0, :, :])
    L = sl.solve_triangular(LinvT, np.eye(LinvT.shape[0]), lower=True, overwrite_b=True, check_finite=False)
    return mu, L, LinvT
