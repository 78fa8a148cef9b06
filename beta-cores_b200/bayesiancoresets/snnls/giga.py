"""GIGA: greedy iterative geodesic ascent (drop-in for bayesiancoresets/snnls/giga.py)."""
import numpy as np

from .. import util
from .. import _native as nv
from ..util.errors import NumericalPrecisionError
from .snnls import SparseNNLS


class GIGA(SparseNNLS):
    _device_algo = 0

    def _device_run_operands(self):
        return self._bn_dev, self.bnorm

    def __init__(self, A, b):
        super().__init__(A, b)
        self._setup()

    def _setup(self):
        self._require_nonzero_columns()
        self.bnorm = np.sqrt(((self.b)**2).sum())                      # giga.py:15-18
        if self.bnorm == 0.:
            raise NumericalPrecisionError('norm of b must be > 0')
        self.bn = self.b / self.bnorm
        if self._N:
            self._bn_dev = self._eng.upload(self.bn)

    def _select(self):
        # giga.py:20-38: direction on the sphere, then score every datapoint against (cdir, xw)
        self._vec(nv.VEC_GIGA_DIR, u=self._u, b=self._bn_dev)
        self._score(nv.SCORE_GIGA, self._u)
        o, best, f, _, _ = self._best()
        cdirnrm = float(o[0])
        if cdirnrm < util.TOL:
            raise NumericalPrecisionError('cdirnrm < TOL: cdirnrm = ' + str(cdirnrm))
        return f

    def _reweight(self, f):
        # giga.py:40-64: closed-form geodesic line search
        self._vec(nv.VEC_GIGA_STEP, xf=self._row(f), aux=self.bnorm, b=self._bn_dev)
        o = self._out[:4].cpu().numpy()
        gA, gB, alpha, beta = (float(x) for x in o)
        if gA <= 0. or gB < 0:
            raise NumericalPrecisionError
        self._scale_and_add(alpha, f, beta)
