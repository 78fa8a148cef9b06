"""Projectors: hold the user's sampler and likelihood callbacks and the S current posterior samples.

Drop-in for bayesiancoresets/coreset/projector.py (same class names, constructor and method
signatures).  `project()` / `project_f()` return the centred (n, S) matrix as a host ndarray like
the reference; when the likelihood is a bound `DevicePotential` the matrix is produced by the
materialise kernel, and the coreset classes bypass it entirely through `fused()`.
"""
import numpy as np

from .._device import Engine, DeviceRows, ptr, stream_ptr
from .._fused import FusedProjection
from ..potentials import DevicePotential
from .. import _native as nv


def evaluate_potential(pot, pts, samples, beta=None, centred=True):
    """(n, S) host array of `pot` on host inputs, through the device materialise pass."""
    pts = np.atleast_2d(np.asarray(pts, dtype=np.float64))
    if not hasattr(samples, 'is_cuda'):               # a device-side sampler hands its samples over as a CUDA tensor
        samples = np.atleast_2d(np.asarray(samples, dtype=np.float64))
    eng = Engine.get()
    # its own workspace: a coreset object may be in the middle of a fused pass on the main one (learn_beta evaluates the
    # beta-gradient of the coreset points between begin() and the data pass)
    fp = FusedProjection(eng, pot, pts.shape[1], ctx_name='eval')
    fp.configure(beta)
    fp.set_samples(samples)
    rows = DeviceRows(eng, pts)
    V, _, _ = fp.materialise(rows, raw=not centred)
    return V.cpu().numpy()


def centre_on_device(mat):
    """row-centre a host matrix produced by an opaque callback (projector.py:26 / :55)."""
    eng = Engine.get()
    t = eng.upload(np.atleast_2d(np.asarray(mat, dtype=np.float64)))
    nv.call('bc_dense_center', eng.ctx(), ptr(t), t.shape[0], t.shape[1], int(t.stride(0)), stream_ptr())
    return t


class Projector(object):
    def project(self, pts, grad=False):
        raise NotImplementedError

    def update(self, wts, pts):
        raise NotImplementedError


class _SamplingProjector(Projector):
    """sampler(S, wts, pts) -> (S, D) is the user's host callback; it is invoked exactly where the
    reference invokes it (constructor, then every update()), so RNG streams line up."""

    def _init_common(self, sampler, projection_dimension, kwargs):
        self.projection_dimension = projection_dimension
        self.sampler = sampler
        self._fused_cache = {}
        self.update(np.array([]), np.array([]))
        self.encoder = kwargs['nl'] if 'nl' in kwargs else None   # learned-feature encoder, forwarded to the callbacks

    def update(self, wts, pts):
        self.samples = self.sampler(self.projection_dimension, wts, pts)

    def _fused_for(self, fn, ncols):
        """FusedProjection for callback `fn`, or None when `fn` is an opaque Python callable."""
        if self.encoder or not isinstance(fn, DevicePotential) or not fn.is_bound():
            return None
        key = (id(fn), ncols)
        if key not in self._fused_cache:
            self._fused_cache[key] = FusedProjection(Engine.get(), fn, ncols)
        return self._fused_cache[key]


class BlackBoxProjector(_SamplingProjector):
    def __init__(self, sampler, projection_dimension, loglikelihood, grad_loglikelihood=None, **kwargs):
        self.loglikelihood = loglikelihood
        self.grad_loglikelihood = grad_loglikelihood
        self._init_common(sampler, projection_dimension, kwargs)

    def fused(self, ncols):
        return self._fused_for(self.loglikelihood, ncols)

    def host_matrix(self, pts, beta=None):
        """un-centred callback output for opaque callables (the black-box path)"""
        if self.encoder:
            return self.loglikelihood(pts, self.samples, self.encoder)
        return self.loglikelihood(pts, self.samples)

    def project(self, pts, grad=False):
        pts = np.atleast_2d(pts)
        f = self.fused(pts.shape[1])
        if f is not None:
            lls = evaluate_potential(self.loglikelihood, pts, self.samples, None, centred=True)
        else:
            lls = centre_on_device(self.host_matrix(pts)).cpu().numpy()
        if grad:
            if self.grad_loglikelihood is None:
                raise ValueError('grad_loglikelihood was requested but not initialized in BlackBoxProjector.project')
            # opaque gradient callback: (n, S, D) host tensor, centred over its LAST axis as the reference does
            # (projector.py:31) -- on the device, as n*S rows of length D.  BatchPSVICoreset does not come through here.
            glls = np.ascontiguousarray(self.grad_loglikelihood(pts, self.samples), dtype=np.float64)
            shp = glls.shape
            return lls, centre_on_device(glls.reshape(-1, shp[-1])).cpu().numpy().reshape(shp)
        return lls


class BetaBlackBoxProjector(_SamplingProjector):
    def __init__(self, sampler, projection_dimension, beta_likelihood, loglikelihood, beta_gradient, **kwargs):
        self.beta_likelihood = beta_likelihood
        self.loglikelihood = loglikelihood
        self.beta_gradient = beta_gradient
        self._init_common(sampler, projection_dimension, kwargs)

    def fused(self, ncols):
        return self._fused_for(self.beta_likelihood, ncols)

    def host_matrix(self, pts, beta=None):
        if self.encoder:
            return self.beta_likelihood(pts, self.samples, beta, self.encoder)
        return self.beta_likelihood(pts, self.samples, beta)

    def project_f(self, pts, beta, grad=False):
        pts = np.atleast_2d(pts)
        f = self.fused(pts.shape[1])
        if f is not None:
            bls = evaluate_potential(self.beta_likelihood, pts, self.samples, beta, centred=True)
        else:
            bls = centre_on_device(self.host_matrix(pts, beta)).cpu().numpy()
        if grad:
            if self.beta_gradient is None:
                raise ValueError('grad_loglikelihood was requested but not initialized in BlackBoxProjector.project')
            if isinstance(self.beta_gradient, DevicePotential) and self.beta_gradient.is_bound():
                glls = evaluate_potential(self.beta_gradient, pts, self.samples, beta, centred=True)
            else:
                glls = centre_on_device(self.beta_gradient(pts, self.samples, beta)).cpu().numpy()
            return bls, glls
        return bls
