"""BetaCoreset: beta-divergence robust coreset (drop-in for bayesiancoresets/coreset/bcores.py).
Greedy selection + projected-ADAM reweighting on the beta-likelihood tangent space; the loop itself is in
_greedy.py, shared with SparseVICoreset."""
import numpy as np
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
        if self.learn_beta:
            # bcores.py:131 calls a method the reference never defines (AttributeError on every call);
            # joint (w, beta) optimisation is listed as a next step in SURVEY 8f.2
            raise NotImplementedError('learn_beta=True: the reference path is broken (bcores.py:131); pass learn_beta=False')
        super()._optimize()

    def get(self):
        keep = self.wts > 0
        return self.wts[keep], self.pts[keep, :], self.idcs[keep], self.beta
