"""SparseVICoreset: greedy KL coreset on the log-likelihood tangent space (drop-in for
bayesiancoresets/coreset/sparsevi.py).  Same device loop as BetaCoreset (_greedy.py), different epilogue."""
from ._greedy import GreedyVICoreset


class SparseVICoreset(GreedyVICoreset):
    def __init__(self, data, ll_projector, n_subsample_select=None, n_subsample_opt=None,
                 opt_itrs=100, step_sched=lambda i: 1./(1.+i), mup=None, SigpInv=None,
                 groups=None, selected_groups=None, initialized=False, enforce_new=False, **kwargs):
        self.mup = mup
        self.SigpInv = SigpInv
        self.enforce_new = enforce_new
        self._init_greedy(data, ll_projector, n_subsample_select, n_subsample_opt, opt_itrs, step_sched, groups, initialized, kwargs)

    def _host_project(self, pts):
        return self.ll_projector.project(pts)
