"""Gaussian mean inference with known covariance: device potentials + host-side conjugate posterior.
Drop-in for the hot-path functions of examples/common/gaussian.py.

With q = (x-th)^T Siginv (x-th):
    gaussian_loglikelihood(x, th, Siginv, logdetSig)          = -d/2 log 2pi - 1/2 logdetSig - q/2   (gaussian.py:7-15)
    gaussian_beta_likelihood(x, th, beta, Siginv, logdetSig)  = e^{-beta q/2}/beta - (1+beta)^(-d/2-1) (gaussian.py:34-44)
    gaussian_beta_gradient(x, th, beta, Siginv, logdetSig)    = d/dbeta of the above as the reference writes it (:46-62)
Call them with the reference's positional signature, or `.bind(Siginv=..., logdetSig=...)` them
(instead of wrapping them in a lambda as examples/zellner_gaussian/main.py:58-67 does) so the
projector can run them fused.
"""
import numpy as np
import scipy.linalg as sl

from bayesiancoresets.potentials import DevicePotential, DeviceGradient

gaussian_loglikelihood = DevicePotential('gaussian', 'loglik', name='gaussian_loglikelihood')
gaussian_beta_likelihood = DevicePotential('gaussian', 'betalik', name='gaussian_beta_likelihood')
gaussian_beta_gradient = DevicePotential('gaussian', 'betagrad', name='gaussian_beta_gradient')
gaussian_grad_x_loglikelihood = DeviceGradient('gaussian', name='gaussian_grad_x_loglikelihood')   # gaussian.py:17-20 (BatchPSVI)


def weighted_post(th0, Sig0inv, Siginv, x, w):
    """conjugate weighted posterior N(mu, L L^T) (gaussian.py:28-32); host, d x d."""
    LSigpInv = np.linalg.cholesky(Sig0inv + w.sum()*Siginv)
    LSigp = sl.solve_triangular(LSigpInv, np.eye(LSigpInv.shape[0]), lower=True, overwrite_b=True, check_finite=False)
    mup = np.dot(LSigp.dot(LSigp.T), np.dot(Sig0inv, th0) + np.dot(Siginv, (w[:, np.newaxis]*x).sum(axis=0)))
    return mup, LSigp, LSigpInv


def gaussian_KL(mu0, Sig0, mu1, Sig1inv):
    """KL(N(mu0, Sig0) || N(mu1, Sig1)) (gaussian.py:22-26); host, evaluation only."""
    t1 = np.dot(Sig1inv, Sig0).trace()
    t2 = np.dot((mu1-mu0), np.dot(Sig1inv, mu1-mu0))
    t3 = -np.linalg.slogdet(Sig1inv)[1] - np.linalg.slogdet(Sig0)[1]
    return 0.5*(t1+t2+t3-mu0.shape[0])


def make_conjugate_sampler(mu0, Sig0inv, Siginv, device=False, prefetch=False):
    """sampler(S, wts, pts) as in examples/zellner_gaussian/main.py:87-92.
    device=True: the samples are formed on the GPU and stay there (bayesiancoresets/util/samplers.py): the host factors the
    d x d posterior precision, the triangular solve against the S normal vectors runs on the device; same numpy stream, same
    samples.  prefetch=True (device form): the normals -- and the coreset class's sub-sample indices -- are drawn one call
    ahead on a helper thread (bayesiancoresets/util/rng.py); the global stream is consumed in the same order."""
    d = mu0.shape[0]
    if device and d <= 160:
        from bayesiancoresets.util import rng
        from bayesiancoresets.util.samplers import ConjugateDeviceSampler
        Sig0inv_mu0 = np.dot(Sig0inv, mu0)

        def precision(wts, pts):
            if pts.shape[0] == 0:
                return Sig0inv + 0.*Siginv, Sig0inv_mu0 + 0.
            return Sig0inv + wts.sum()*Siginv, Sig0inv_mu0 + np.dot(Siginv, (wts[:, np.newaxis]*pts).sum(axis=0))
        return ConjugateDeviceSampler(d, precision, rng.activate() if prefetch else None,
                                      device_model=(1, Sig0inv, Siginv, Sig0inv_mu0, None))

    def sampler(S, wts, pts):
        if pts.shape[0] == 0:
            wts = np.zeros(1)
            pts = np.zeros((1, d))
        mu, L, _ = weighted_post(mu0, Sig0inv, Siginv, pts, wts)
        if device:
            from bayesiancoresets.util.samplers import affine_samples
            return affine_samples(mu, L, np.random.randn(S, d))
        return mu + np.random.randn(S, d).dot(L.T)
    return sampler
