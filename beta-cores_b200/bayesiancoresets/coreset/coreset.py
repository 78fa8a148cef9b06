"""Common state and guards of every coreset construction (role of
bayesiancoresets/coreset/coreset.py:7-71): weights / indices / points, `build` pre-checks,
`get` filtering, `optimize` with revert-on-worse."""
import logging
import secrets
import numpy as np

from .. import util
from ..util.errors import NumericalPrecisionError


class Coreset(object):
    def __init__(self, initial_sz=10, wts=None, idcs=None, pts=None):
        # NB the reference shares one mutable default array between all instances
        # (coreset.py:8) and grows it in place; here every instance starts from fresh arrays.
        self.alg_name = '%s-%s' % (type(self).__name__, secrets.token_hex(3))
        self.log = logging.LoggerAdapter(logging.getLogger(), {'id': self.alg_name})
        self.reached_numeric_limit = False
        self.wts = np.array([]) if wts is None else wts
        self.idcs = np.array([], dtype=np.int64) if idcs is None else idcs
        self.pts = np.array([]) if pts is None else pts

    def reset(self):
        self.wts = np.array([])
        self.idcs = np.array([], dtype=np.int64)
        self.pts = np.array([])
        self.reached_numeric_limit = False

    def size(self):
        return (self.wts > 0).sum()

    def get(self):
        keep = self.wts > 0
        return self.wts[keep], self.pts[keep, :], self.idcs[keep]

    def error(self):
        raise NotImplementedError()

    def build(self, itrs, sz):
        """grow towards size sz with at most itrs greedy iterations (never shrinks)"""
        if self.reached_numeric_limit:
            return
        cur = self.size()
        if sz < cur:
            raise ValueError('%s.build(): requested coreset of size < the current size, but cannot shrink coresets; '
                             'returning. Requested size = %s current size = %s' % (self.alg_name, sz, cur))
        self._build(itrs, sz)
        if self.reached_numeric_limit:
            self.log.warning('the numeric limit has been reached. No more points will be added. size = %s, error = %s'
                             % (self.size(), self.error()))

    def optimize(self):
        """re-solve the weights only; keep the previous solution if the error got worse"""
        saved = (self.wts.copy(), self.idcs.copy(), self.pts.copy())
        try:
            before = self.error()
            self._optimize()
            after = self.error()
            if after > before*(1.+util.TOL):
                raise NumericalPrecisionError(
                    'self.optimize() returned a solution with increasing error. Numeric limit possibly reached: '
                    'preverr = %s err = %s. If the two errors are very close, try bc.util.set_tolerance(tol) with '
                    'tol > current tol = %s before running' % (before, after, util.TOL))
        except NumericalPrecisionError as e:
            self.log.warning(e)
            self.wts, self.idcs, self.pts = saved
            self.reached_numeric_limit = True

    def _optimize(self):
        raise NotImplementedError

    def _build(self, itrs, sz):
        raise NotImplementedError
