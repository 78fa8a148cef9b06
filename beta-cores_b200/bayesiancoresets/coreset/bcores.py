"""BetaCoreset: beta-divergence robust coreset (drop-in for bayesiancoresets/coreset/bcores.py).
Greedy selection + projected-ADAM reweighting on the beta-likelihood tangent space; the loop itself is in
_greedy.py, shared with SparseVICoreset."""
import numpy as np
from ..util import rng
from ._greedy import GreedyVICoreset


class BetaCoreset(GreedyVICoreset):
    _uses_beta = True

    def __init__(self, data, ll_projector, n_subsample_select=None, n_subsample_opt=None,
                 opt_itrs=100, step_sched=lambda i: 1./(1.+i), mup=None, SigpInv=None,
                 beta=.5, learn_beta=True, groups=None, selected_groups=None, initialized=False, **kw):
        self.mup = mup
        self.SigpInv = SigpInv
        self.beta = beta
        self.learn_beta = learn_beta
        self._init_greedy(data, ll_projector, n_subsample_select, n_subsample_opt, opt_itrs, step_sched, groups, initialized, kw)

    def _beta(self):
        return self.beta

    def _host_project(self, pts):
        return self.ll_projector.project_f(pts, self.beta)

    def _optimize(self):
        if not self.learn_beta:
            return super()._optimize()
        # Joint (w, beta) optimisation, bcores.py:127-140.  The reference's branch calls `_get_projection_ii`, which it
        # never defines (AttributeError); what it must return is spelled out by the lines around the call and by
        # BetaBlackBoxProjector.project_f(grad=True) (projector.py:56-61): the projection at the ITERATE's beta plus the
        # centred beta-gradient of the coreset points.  Built to that reading; no reference output exists
        # to pin it, the oracle restates the same lines (oracle/np_coresets.py::GreedyVILearnBeta).
        from ..util.opt import partial_nn_opt
        t = self._get_tangent()
        M = self.wts.shape[0]
        if M == 0:
            for i in range(self.opt_itrs):             # gradient of nothing: the sampler and the subsample draw still run
                self.ll_projector.update(self.wts, self.pts)
                if self.n_subsample_opt is not None:
                    rng.randint(self._n_total, self.n_subsample_opt)
            return
        core = self._core_operand(t)
        gw, gb = t.eng.empty(M), t.eng.empty(M)
        scaling = 1. if self.n_subsample_opt is None else self._n_total/self.n_subsample_opt

        def grd(x):
            w, beta = np.ascontiguousarray(x[:-1]), float(x[-1])
            self.beta = beta                           # opaque callbacks read it through _host_project
            t.begin(w, self.pts, beta)
            if self.n_subsample_opt is not None:
                sub_idcs = rng.randint(self._n_total, self.n_subsample_opt)
            elif self.groups is not None and not self._groups_cover:
                sub_idcs = self._groups_flat
            else:
                sub_idcs = None
            colsum = t.colsum(sub_idcs)
            Vc = t.core_rows(core)
            resid = t.residual(colsum, scaling, Vc, t.eng.upload(w))
            t.grad(Vc, resid, gw)                                                   # -corevecs.resid / S       (:133)
            bgrads = self.ll_projector.project_f(self.pts, beta, grad=True)[1]      # centred d/dbeta, (M, S)
            t.grad(t.eng.upload(np.ascontiguousarray(bgrads, dtype=np.float64)), resid, gb)   # -betagrads.resid / S
            return np.hstack((gw.cpu().numpy(), 1e-5*w.dot(gb.cpu().numpy())))   # betagrad = -1e-5 w.(betagrads.resid)/S  (:134)
        x0 = np.hstack((self.wts, np.asarray([self.beta], dtype=np.float64)))
        xf = partial_nn_opt(x0, grd, np.arange(x0.shape[0]), self.opt_itrs, step_sched=self.step_sched)
        self.wts = xf[:-1]
        self.beta = xf[-1]                             # numpy float64, as in the reference (:140)

    def get(self):
        keep = self.wts > 0
        return self.wts[keep], self.pts[keep, :], self.idcs[keep], self.beta
