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


class _PinnedStage(object):
    """where a draw of normals goes: one of the owner's two pinned staging buffers.  Called with a finished draw it copies;
    through `into` the native generator of util/rng.py writes the draw there directly (runs on the thread that draws)."""

    def __init__(self, owner):
        self.o = owner

    def __call__(self, r):
        k, pin = self.o._stage_buf(r.shape)
        pin.numpy()[...] = r
        return k, pin

    def into(self, shape):
        k, pin = self.o._stage_buf(shape)
        return (k, pin), pin.numpy()


class ConjugateDeviceSampler(object):
    """The device form of a conjugate sampler, built for the optimiser loop (one call per ADAM step):
        precision(wts, pts) -> (H, v)   the D x D posterior precision and the linear term (host, tiny)
        samples = mu + C^-1 R^T,  C = chol(H) (LAPACK dpotrf on the host), mu = C^-1 C^-T v (two dtrtrs), R = randn(S, D)
    -- the same distribution AND the same samples as the reference's `mu + randn(S, D).dot(LSigp.T)` with
    LSigp = C^-1 (gaussian.py:28-32, model_neurlinr.py:115-122), without forming the inverse: the triangular solve
    against the S normal vectors runs on the GPU (csrc/bc_sampler.cu::k_sample_solve).  [mu | C] goes up in one copy from
    pinned memory, the normals in another; with `ahead` (util/rng.py) they are drawn one call ahead on a helper thread,
    already into pinned memory."""

    def __init__(self, D, precision, ahead=None, device_model=None):
        self.D, self.precision, self.ahead = int(D), precision, ahead
        self.st = None
        # (model id, A0 = Sig0inv, A1 = Siginv or None, v0 = Sig0inv mu0, sigsq): lets the optimiser loop form the posterior on the
        # device from device-resident weights (device_step), with no host round trip per step
        self.device_model = device_model
        self._dm = None
        self._stage = _PinnedStage(self)

    def _state(self, eng):
        D = self.D
        if self.st is None:
            self.st = {'k': 0, 'pin': [None, None], 'ev': [None, None], 'ml': eng.empty(D*D + D),
                       'ml_pin': torch.empty(D*D + D, dtype=torch.float64).pin_memory(), 'ml_ev': None,
                       'info': torch.zeros(2, dtype=torch.int32, device=eng.device)}
        return self.st

    def _normals_to_device(self, eng, S):
        st = self.st
        if self.ahead is not None:
            k, pin = self.ahead.randn(S, self.D, self._stage)
        else:
            k, pin = self._stage(np.random.randn(S, self.D))
        Rd = pin.to(eng.device, non_blocking=True)
        st['ev'][k] = torch.cuda.Event()
        st['ev'][k].record()
        return Rd

    def device_step(self, S, w_dev, core):
        """the sampler call of one optimiser step with the weights ALREADY on the device (`w_dev`, M doubles) and the coreset
        points as resident rows (`core`, a DeviceRows): posterior factor, mean and samples are formed by kernels
        (bc_conjugate_factor, bc_sample_solve), nothing is read back -- the host only queues work and draws the normals."""
        k = self.graph_normals(S)
        theta = self.graph_enqueue(S, w_dev, core, k)
        self.graph_launched(k)
        return theta

    # The same call in three parts, so that the device work of a step can be CAPTURED into a CUDA graph and replayed
    # (coreset/_greedy.py): only graph_enqueue touches the stream, and it names nothing but fixed buffers.
    def graph_normals(self, S):
        """host part: this step's S x D normals, drawn in the reference's order, are in pinned staging buffer k on return"""
        if self.ahead is not None:
            self.ahead.begin_cycle()
        self._state(Engine.get())
        if self.ahead is not None:
            k, _ = self.ahead.randn(S, self.D, self._stage)
        else:
            k, _ = self._stage(np.random.randn(S, self.D))
        return k

    def graph_enqueue(self, S, w_dev, core, k):
        """stream part: factor kernel, upload of staging buffer k, solve kernel.  Returns the (S, D) samples."""
        eng = Engine.get()
        st = self._state(eng)
        D = self.D
        if self._dm is None:
            model, A0, A1, v0, sigsq = self.device_model
            self._dm = (model, eng.upload(np.ascontiguousarray(A0, dtype=np.float64)),
                        eng.upload(np.ascontiguousarray(A1, dtype=np.float64)) if A1 is not None else None,
                        eng.upload(np.ascontiguousarray(v0, dtype=np.float64)), float(sigsq) if sigsq is not None else 1.0)
        model, A0, A1, v0, sigsq = self._dm
        ctx = eng.ctx('sampler')
        nv.call('bc_conjugate_factor', ctx, model, ptr(core.t), core.ld, ptr(w_dev), core.n_local, D, ptr(A0), ptr(A1), ptr(v0), sigsq,
                ptr(st['ml'][:D]), ptr(st['ml'][D:]), ptr(st['info']), stream_ptr())
        Rd = st['pin'][k].to(eng.device, non_blocking=True)
        theta = eng.empty(S, D)
        nv.call('bc_sample_solve_hinted', ctx, ptr(st['ml'][:D]), ptr(st['ml'][D:]), ptr(Rd), S, D, ptr(theta), int(theta.stride(0)),
                ptr(st['info']), stream_ptr())
        return theta

    def graph_launched(self, k):
        """after the stream part (or a replay of it) has been queued: staging buffer k is free again once it has run"""
        self.st['ev'][k] = torch.cuda.Event()
        self.st['ev'][k].record()

    def supports_device_step(self):
        return self.device_model is not None and self.D <= 160

    def _stage_buf(self, shape):
        """the next of two pinned staging buffers, free again (the upload that last used it has left it)"""
        st = self.st
        k = st['k'] = st['k'] ^ 1
        if st['pin'][k] is None or tuple(st['pin'][k].shape) != tuple(shape):
            st['pin'][k] = torch.empty(*shape, dtype=torch.float64).pin_memory()
        if st['ev'][k] is not None:
            st['ev'][k].synchronize()
        return k, st['pin'][k]

    def __call__(self, S, wts, pts):
        import scipy.linalg as sl
        if self.ahead is not None:
            self.ahead.begin_cycle()
        eng = Engine.get()
        D = self.D
        st = self._state(eng)
        H, v = self.precision(np.asarray(wts, dtype=np.float64), pts)
        C, info = sl.lapack.dpotrf(H, lower=1, overwrite_a=0)
        if info != 0:
            raise np.linalg.LinAlgError('posterior precision not positive definite (dpotrf info %d)' % info)
        # the reference's mean is LSigp LSigp^T v with LSigp = C^-1, i.e. C^-1 (C^-T v) = (C^T C)^-1 v -- NOT H^-1 v = (C C^T)^-1 v
        # (gaussian.py:31, model_neurlinr.py:121 multiply by the inverse FACTOR's Gram matrix); two triangular solves
        y, info = sl.lapack.dtrtrs(C, v, lower=1, trans=1)
        mu, info = sl.lapack.dtrtrs(C, y, lower=1, trans=0)
        if st['ml_ev'] is not None:
            st['ml_ev'].synchronize()
        buf = st['ml_pin'].numpy()
        buf[:D] = mu
        buf[D:] = np.tril(C).ravel()
        st['ml'].copy_(st['ml_pin'], non_blocking=True)
        st['ml_ev'] = torch.cuda.Event()
        st['ml_ev'].record()
        Rd = self._normals_to_device(eng, S)
        theta = eng.empty(S, D)
        nv.call('bc_sample_solve', eng.ctx('sampler'), ptr(st['ml'][:D]), ptr(st['ml'][D:]), ptr(Rd), S, D, ptr(theta), int(theta.stride(0)),
                stream_ptr())
        return theta

    def drain(self):
        if self.ahead is not None:
            self.ahead.drain()
