"""Device side of the conjugate / Laplace samplers.

Every sampler of the reference's drivers ends in the same line, `mu + np.random.randn(S, D).dot(L.T)`
(examples/zellner_gaussian/main.py:87-92, zellner_logreg/main.py:139-144, model_neurlinr.py:115-122): D x D algebra for
(mu, L), then an S x D x D product.  `affine_samples` forms that product on the GPU (csrc/bc_sampler.cu::k_sample_affine)
from normals drawn on the host -- the numpy stream is consumed exactly as by the host line -- and returns the samples as a
device tensor, which the fused projection takes without a round trip through host memory.
"""
import numpy as np
import torch

from .. import _native as nv
from .._device import Engine, ptr, stream_ptr


def affine_samples(mu, L, R):
    """mu + R.dot(L.T) on the device.  mu (D,), L (D, D) lower triangular, R (S, D): host arrays.  Returns a (S, D) CUDA tensor."""
    eng = Engine.get()
    mu = np.ascontiguousarray(mu, dtype=np.float64)
    L = np.ascontiguousarray(L, dtype=np.float64)
    R = np.ascontiguousarray(R, dtype=np.float64)
    S, D = R.shape
    if L.shape != (D, D) or mu.shape != (D,):
        raise ValueError('affine_samples: mu (D,), L (D, D), R (S, D) expected')
    if np.any(np.triu(L, 1) != 0.):
        raise ValueError('affine_samples: L must be lower triangular (the kernel reads its lower triangle only)')
    d_mu, d_L, d_R = eng.upload(mu), eng.upload(L), eng.upload(R)
    out = eng.empty(S, D)
    nv.call('bc_sample_affine', eng.ctx('sampler'), ptr(d_mu), ptr(d_L), ptr(d_R), S, D, ptr(out), int(out.stride(0)), stream_ptr())
    return out
