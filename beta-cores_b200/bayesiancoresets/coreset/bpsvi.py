"""BatchPSVICoreset: batch pseudo-coreset (drop-in for bayesiancoresets/coreset/bpsvi.py:6-65).

sz pseudo-points are drawn from the data, then their weights AND locations are optimised jointly by projected ADAM
(only the weights are clamped).  Every optimiser step is one pass of the fused projection over the data rows -- the same
hot path as BetaCoreset / SparseVICoreset -- plus, on the pseudo-point side, the (M, S) projection of the points, the
weight gradient and the location gradient.  The latter is contracted with the residual on the device
(bc_core_pgrad for the built-in models, bc_dense_pgrad for opaque gradient callbacks); the reference's (M, S, D)
gradient tensor (bpsvi.py:39,54) is never formed for the built-in models.

DiffPrivBatchPSVICoreset: the reference imports coreset/dpbpsvi.py (coreset/__init__.py:6) but does not ship it; the
name is exported so that `import bayesiancoresets` keeps the reference's surface."""
import numpy as np
from ..util import rng
import torch

from .. import _native as nv
from .._device import DeviceRows, ptr, stream_ptr
from ..potentials import DeviceGradient
from ..util.opt import partial_nn_opt
from .coreset import Coreset
from ._greedy import GreedyVICoreset, _FusedTangent


class BatchPSVICoreset(GreedyVICoreset):
    def __init__(self, data, ll_projector, opt_itrs, n_subsample_opt=None, step_sched=lambda m: lambda i: 1./(1.+i), mup=None,
                 Zmean=None, SigpInv=None, diagnostics=False, **kw):
        self.mup = mup
        self.SigpInv = SigpInv
        self._init_greedy(data, ll_projector, None, n_subsample_opt, opt_itrs, step_sched, None, False, kw)

    def _host_project(self, pts):
        return self.ll_projector.project(pts)

    # bpsvi.py:17-24 (itrs is ignored, like in the reference)
    def _build(self, itrs, sz):
        rng.drain()      # np.random.choice draws on its own: no look-ahead in flight
        init_idcs = np.random.choice(self._n_total, size=sz, replace=False)
        if self.rows is not None and not isinstance(self.data, np.ndarray):
            self.pts = np.vstack([np.asarray(self.data[int(i)], dtype=np.float64).reshape(1, -1) for i in init_idcs])
        else:
            self.pts = np.asarray(self.data[init_idcs], dtype=np.float64)
        self.wts = self._n_total/sz*np.ones(sz)
        self.idcs = init_idcs
        self._optimize()

    # bpsvi.py:44-62
    def _optimize(self):
        t = self._get_tangent()
        sz = self.wts.shape[0]
        d = self._ncols
        if sz == 0:
            return
        prj = self.ll_projector
        gfn = getattr(prj, 'grad_loglikelihood', None)
        if gfn is None:
            raise ValueError('grad_loglikelihood was requested but not initialized in BlackBoxProjector.project')
        fused_grad = isinstance(gfn, DeviceGradient) and gfn.is_bound() and isinstance(t, _FusedTangent) \
            and gfn.model == t.fp.pot.model
        g = t.eng.empty(sz + sz*d)
        scaling = 1. if self.n_subsample_opt is None else self._n_total/self.n_subsample_opt

        def grd(x_host, x_dev):
            w_host = x_host[:sz]
            p_host = x_host[sz:].reshape((sz, d))
            t.begin(w_host, p_host, None)                                  # sampler first (bpsvi.py:28)
            sub_idcs = None if self.n_subsample_opt is None else rng.randint(self._n_total, self.n_subsample_opt)
            colsum = t.colsum(sub_idcs)
            core = DeviceRows(t.eng, p_host) if isinstance(t, _FusedTangent) else p_host
            Vc = t.core_rows(core)
            w_dev = x_dev[:sz]
            resid = t.residual(colsum, scaling, Vc, w_dev)
            t.grad(Vc, resid, g[:sz])                                      # wgrad (bpsvi.py:52)
            out = g[sz:].view(sz, d)
            if fused_grad:
                nv.call('bc_core_pgrad', t.fp.ctx, ptr(core.t), sz, core.ld, ptr(w_dev), ptr(resid), ptr(out), d, stream_ptr())
            else:
                G = np.ascontiguousarray(gfn(p_host, prj.samples), dtype=np.float64)      # (M, S, D) from the user's callback
                if G.shape != (sz, resid.shape[0]-1, d):
                    raise ValueError('grad_loglikelihood returned shape %s, expected %s' % (G.shape, (sz, resid.shape[0]-1, d)))
                Gd = t.eng.upload(G)
                nv.call('bc_dense_pgrad', t.ctx, ptr(Gd), sz, G.shape[1], d, ptr(w_dev), ptr(resid), 1, ptr(out), d, stream_ptr())
            return g
        grd.wants_device_iterate = True
        x0 = np.hstack((self.wts, self.pts.reshape(sz*d)))
        xf = partial_nn_opt(x0, grd, np.arange(sz), self.opt_itrs, step_sched=self.step_sched(sz))
        self.wts = xf[:sz]
        self.pts = xf[sz:].reshape((sz, d))

    def error(self):
        return 0.       # the reference has no KL estimate either (bpsvi.py:64-65)


class DiffPrivBatchPSVICoreset(Coreset):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError('DiffPrivBatchPSVICoreset: the reference imports coreset/dpbpsvi.py but does not ship it')
