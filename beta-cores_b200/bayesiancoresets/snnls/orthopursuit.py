"""Orthogonal matching pursuit with non-negative re-solve (drop-in for bayesiancoresets/snnls/orthopursuit.py).
The signed arg-max over all N datapoints is the device pass; the Lawson-Hanson re-solve on the S x m active block runs on
the device too (bc_nnls on the cached rows, warm-started from the current weights), scipy's routine on the host being the
fallback for blocks wider than the kernel takes (orthopursuit.py:40)."""
import numpy as np
import torch

from .. import _native as nv
from .._device import ptr
from .snnls import SparseNNLS


class OrthoPursuit(SparseNNLS):
    def __init__(self, A, b):
        super().__init__(A, b)
        self._setup()

    def _setup(self):
        self._require_nonzero_columns()
        if self._N:
            self._active_dev = self._eng.zeros(max(self._nl, 1), dtype=torch.uint8)     # this rank's block of datapoints

    def _sync_active(self):
        self._active_dev.zero_()
        nz = [i-self._row0 for i, x in zip(self._act, self._aw) if x > 0 and self._row0 <= i < self._row0+self._nl]
        if nz:
            self._active_dev[torch.as_tensor(nz, dtype=torch.int64, device=self._eng.device)] = 1

    def _select(self):
        # orthopursuit.py:17-35: positive direction over all datapoints, negative over the active set
        self._vec(nv.VEC_RESID, u=self._u)
        self._sync_active()
        self._score(nv.SCORE_OMP, self._u, self._active_dev)
        _, pos, fpos, neg, fneg = self._best()
        if self.size() == 0 or fneg < 0:
            return fpos
        return fpos if pos >= neg else fneg

    def _reweight(self, f):
        # orthopursuit.py:37-42
        self._aw[self._activate(int(f))] = 1.
        nz = sorted(i for i, x in zip(self._act, self._aw) if x > 0)
        self._assign(nz, self._nnls_on(nz, fresh=int(f)))
