"""Uniform-subsampling baseline (role of bayesiancoresets/coreset/sampling.py:5-47): no arithmetic on
the hot path, pure host RNG bookkeeping -- kept so the baselines in the drivers still run
(examples/zellner_neural_linear/main.py:150 passes `wts=`, `idcs=`, `pts=` and `groups=`).

Counts: every draw of an already-held point (or group) bumps its count; the weights are the counts rescaled to
sum to N.  A coreset warm-started through `wts=` starts with one count per initial point (sampling.py:9-11)."""
import numpy as np
from ..util import rng
from .coreset import Coreset


class UniformSamplingCoreset(Coreset):
    def __init__(self, data, groups=None, selected_groups=None, **kw):
        super().__init__(**kw)
        self.data = data
        warm = 'wts' in kw
        self.ct_idcs = [int(i) for i in np.asarray(self.idcs).ravel()] if warm else []
        self.cts = [1 for _ in self.ct_idcs]
        self.groups = groups
        self.selected_groups = []          # the reference ignores the constructor argument too (sampling.py:16)

    def reset(self):
        self.cts, self.ct_idcs = [], []
        super().reset()

    def _count_weights(self):
        c = np.asarray(self.cts, dtype=np.float64)
        return self.data.shape[0]*c/c.sum()

    def _draw_point(self):
        f = int(np.random.randint(self.data.shape[0]))
        try:
            self.cts[self.ct_idcs.index(f)] += 1
        except ValueError:
            self.ct_idcs.append(f)
            self.cts.append(1)

    def _draw_group(self):
        # sampling.py:33-47: a group is taken whole, once; its rows get one count each
        g = int(np.random.randint(len(self.groups)))
        if g in self.selected_groups:
            return
        rows = np.asarray(self.groups[g], dtype=np.int64)
        new = np.atleast_2d(self.data[rows, :])
        self.ct_idcs.append(self.groups[g])
        self.cts.extend([1]*new.shape[0])
        self.idcs = np.concatenate((np.asarray(self.idcs, dtype=np.int64).ravel(), rows))
        held = np.asarray(self.pts, dtype=np.float64).reshape(-1, self.data.shape[1])
        self.pts = np.vstack((held, new))
        self.wts = self._count_weights()
        self.selected_groups.append(g)

    def _build(self, itrs, sz):
        if self.size()+itrs > sz:
            raise ValueError('%s._build(): # itrs + current size cannot exceed total desired size sz. # itr = %s cur sz: %s '
                             'desired sz: %s' % (self.alg_name, itrs, self.size(), sz))
        rng.drain()      # plain np.random draws below: no sampler look-ahead may be in flight
        if self.groups is not None:
            for _ in range(itrs):
                self._draw_group()
            return
        for _ in range(itrs):
            self._draw_point()
        self.wts = self._count_weights()
        self.idcs = np.array(self.ct_idcs)
        self.pts = self.data[self.idcs]

    def error(self):
        return 0.

    def _optimize(self):
        pass
