"""Random-sampling baselines (role of bayesiancoresets/snnls/sampling.py): host RNG bookkeeping on the column
norms, no pass over the matrix -- outside the accelerated path, kept for the drivers' baselines."""
import numpy as np
from .snnls import SparseNNLS


class ImportanceSampling(SparseNNLS):
    """sample datapoints with probability proportional to their column norm; weights = counts / (total * p)"""

    def __init__(self, A, b):
        super().__init__(A, b)
        self.cts = np.zeros(self._N)
        norms = self._norms_host.copy() if self._N else np.zeros(0)
        if np.any(norms > 0):
            self.ps = norms/norms.sum()
        else:
            self.ps = np.ones(self._N)/float(max(self._N, 1))
        self.check_error_monotone = False

    def reset(self):
        super().reset()
        self.cts = np.zeros(self._N)

    def _select(self):
        return np.random.choice(self.ps.shape[0], p=self.ps)

    def _reweight(self, f):
        self.cts[f] += 1
        with np.errstate(divide='ignore', invalid='ignore'):
            self.w = (self.cts/self.cts.sum())/self.ps


class UniformSampling(ImportanceSampling):
    def __init__(self, A, b):
        super().__init__(A, b)
        self.ps = np.ones(self._N)/float(max(self._N, 1))
