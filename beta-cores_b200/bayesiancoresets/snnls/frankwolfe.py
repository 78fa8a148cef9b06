"""Frank-Wolfe on the weighted-norm polytope (drop-in for bayesiancoresets/snnls/frankwolfe.py)."""
import numpy as np

from .. import _native as nv
from ..util.errors import NumericalPrecisionError
from .snnls import SparseNNLS


class FrankWolfe(SparseNNLS):
    _device_algo = 1

    def _device_run_operands(self):
        return self._b_dev, self.Anorms.sum()

    def __init__(self, A, b):
        super().__init__(A, b)
        self._setup()

    def _setup(self):
        self._require_nonzero_columns()
        self.Anorms = self._norms_host if self._N else np.zeros(0)     # frankwolfe.py:10

    def _select(self):
        # frankwolfe.py:15-17: residual, then arg-max of the normalised correlations
        self._vec(nv.VEC_RESID, u=self._u)
        self._score(nv.SCORE_FW, self._u)
        return self._best()[2]

    def _reweight(self, f):
        # frankwolfe.py:19-40
        if self.size() == 0:
            alpha = 0.
            beta = self.Anorms.sum() / self.Anorms[f]
        else:
            nsum = self.Anorms.sum()
            nf = self.Anorms[f]
            self._vec(nv.VEC_FW_STEP, xf=self._row(f), aux=nsum/nf)
            o = self._out[:2].cpu().numpy()
            gammanum, gammadenom = float(o[0]), float(o[1])
            if gammanum < 0. or gammadenom == 0. or gammanum > gammadenom:
                raise NumericalPrecisionError('precision loss in gammanum/gammadenom: num = %s denom = %s' % (gammanum, gammadenom))
            alpha = 1. - gammanum/gammadenom
            beta = nsum/nf*gammanum/gammadenom
        self._scale_and_add(alpha, f, beta)
