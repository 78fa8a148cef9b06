"""Neural-linear regression (Bayesian linear head on fixed features): device potentials + host
conjugate posterior.  Drop-in for the hot-path functions of examples/common/model_neurlinr.py.

Data rows are z_n = [phi(x_n), y_n]; with u = phi.theta and r2 = y^2 - 2 u y + u^2:
    neurlinr_loglikelihood(z, th, sigsq)          = -1/2 log(2 pi s2) - r2/(2 s2)                      (model_neurlinr.py:90-97)
    neurlinr_beta_likelihood(z, th, beta, sigsq)  = (2 pi s2)^(-beta/2) (-(b+1)/b e^{-b r2/(2 s2)} + 1/sqrt(1+b))   (:102-110)
"""
import numpy as np
import scipy.linalg as sl

from bayesiancoresets.potentials import DevicePotential

neurlinr_loglikelihood = DevicePotential('neurlin', 'loglik', name='neurlinr_loglikelihood')
neurlinr_beta_likelihood = DevicePotential('neurlin', 'betalik', name='neurlinr_beta_likelihood')


def weighted_post(th0, Sig0inv, sigsq, z, w):
    """conjugate weighted posterior of the linear head (model_neurlinr.py:115-122); host, D x D."""
    z = np.atleast_2d(z)
    X, Y = z[:, :-1], z[:, -1]
    LSigpInv = np.linalg.cholesky(Sig0inv + (w[:, np.newaxis]*X).T.dot(X)/sigsq)
    LSigp = sl.solve_triangular(LSigpInv, np.eye(LSigpInv.shape[0]), lower=True, overwrite_b=True, check_finite=False)
    mup = np.dot(LSigp.dot(LSigp.T), np.dot(Sig0inv, th0) + (w[:, np.newaxis]*Y[:, np.newaxis]*X).sum(axis=0)/sigsq)
    return mup, LSigp, LSigpInv


def make_conjugate_sampler(mu0, Sig0inv, sigsq):
    D = mu0.shape[0]

    def sampler(S, wts, pts):
        if pts.shape[0] == 0:
            wts = np.zeros(1)
            pts = np.zeros((1, D+1))
        mu, L, _ = weighted_post(mu0, Sig0inv, sigsq, pts, wts)
        return mu + np.random.randn(S, D).dot(L.T)
    return sampler
