"""Uniform-subsampling baseline (role of bayesiancoresets/coreset/sampling.py): no arithmetic on
the hot path, pure host RNG bookkeeping -- kept so baselines in the drivers still run."""
import numpy as np
from .coreset import Coreset


class UniformSamplingCoreset(Coreset):
    def __init__(self, data, **kw):
        super().__init__(**kw)
        self.data = data
        self.cts = []
        self.ct_idcs = []

    def reset(self):
        self.cts = []
        self.ct_idcs = []
        super().reset()

    def _build(self, itrs, sz):
        if self.size()+itrs > sz:
            raise ValueError('%s._build(): # itrs + current size cannot exceed total desired size sz. # itr = %s cur sz: %s '
                             'desired sz: %s' % (self.alg_name, itrs, self.size(), sz))
        for i in range(itrs):
            f = np.random.randint(self.data.shape[0])
            if f in self.ct_idcs:
                self.cts[self.ct_idcs.index(f)] += 1
            else:
                self.ct_idcs.append(f)
                self.cts.append(1)
        self.wts = self.data.shape[0]*np.array(self.cts)/np.array(self.cts).sum()
        self.idcs = np.array(self.ct_idcs)
        self.pts = self.data[self.idcs]

    def error(self):
        return 0.

    def _optimize(self):
        pass
