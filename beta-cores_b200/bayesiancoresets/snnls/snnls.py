"""SparseNNLS: greedy sparse non-negative least squares  min_w |A w - b|,  w >= 0, nnz(w) small
(drop-in for bayesiancoresets/snnls/snnls.py:8-106: same constructor, build loop, monotone-error check,
one retry, numeric-limit latch, optimize()).

Device layout: the reference's A is (S, N) with one column per datapoint; here the matrix lives in HBM
datapoint-major, V = A^T (N x S, row n contiguous), so that the per-iteration score pass streams each
datapoint's S values with coalesced 16-byte loads (bc_dense_score).  The iterate A w is formed from the
active rows only (bc_dense_combine); the S-length line-search algebra is one single-CTA kernel per step
(bc_vec_step); the host sees a handful of scalars per iteration and applies the reference's guards to them.
"""
import logging
import secrets
import numpy as np
import torch
from scipy.optimize import nnls

from .. import util
from .. import _native as nv
from .._device import Engine, ptr, stream_ptr
from ..util.errors import NumericalPrecisionError


class SparseNNLS(object):
    def __init__(self, A, b, check_error_monotone=True):
        self.alg_name = '%s-%s' % (type(self).__name__, secrets.token_hex(3))
        self.log = logging.LoggerAdapter(logging.getLogger(), {'id': self.alg_name})
        self.A = A
        self.b = np.asarray(b, dtype=np.float64)
        self.reached_numeric_limit = False
        self.check_error_monotone = check_error_monotone
        self._eng = Engine.get()
        self._ctx = self._eng.ctx()
        A = np.asarray(A, dtype=np.float64)
        self._S, self._N = (A.shape[0], A.shape[1]) if A.ndim == 2 else (0, 0)
        if A.size:
            At = self._eng.upload(A)                                   # (S, N) as the caller laid it out
            V = self._eng.empty(self._N, self._S)
            nv.call('bc_transpose', self._ctx, ptr(At), self._S, self._N, int(At.stride(0)), ptr(V), self._S, stream_ptr())
            del At
            self._attach(V, None)
        self._reset_weights()

    @classmethod
    def from_device(cls, V, norms, b_host):
        """build a solver directly on a device-resident datapoint-major matrix V (n x S), e.g. the output of
        the materialise pass (HilbertCoreset); `A` is exposed lazily as V^T on the host."""
        self = cls.__new__(cls)
        self.alg_name = '%s-%s' % (cls.__name__, secrets.token_hex(3))
        self.log = logging.LoggerAdapter(logging.getLogger(), {'id': self.alg_name})
        self.A = None
        self.b = np.asarray(b_host, dtype=np.float64)
        self.reached_numeric_limit = False
        self.check_error_monotone = True
        self._eng = Engine.get()
        self._ctx = self._eng.ctx()
        self._N, self._S = int(V.shape[0]), int(V.shape[1])
        if self._N:
            self._attach(V, norms)
        self._reset_weights()
        self._setup()
        return self

    def _attach(self, V, norms):
        self._V = V
        self._ldv = int(V.stride(0))
        if norms is None:
            norms = self._eng.empty(self._N)
            nv.call('bc_dense_rownorms', self._ctx, ptr(V), self._N, self._S, self._ldv, ptr(norms), stream_ptr())
        self._norms = norms
        self._norms_host = norms.cpu().numpy()                        # N doubles, once: zero-column check, FW's norm sum
        self._b_dev = self._eng.upload(self.b)
        self._xw = self._eng.zeros(self._S)
        self._u = self._eng.zeros(2*self._S)
        self._out = self._eng.zeros(8)
        self._xw_valid = False

    def _setup(self):
        pass

    def _require_nonzero_columns(self):
        if self._N and np.any(self._norms_host == 0):
            raise ValueError(self.alg_name+'.__init__(): A must not have any 0 columns')

    # ---- weights: sparse on the inside, the reference's dense `w` on the outside ----
    def _reset_weights(self):
        self._act = []            # indices that ever received weight, in order of first selection
        self._aw = []             # their weights
        self._xw_valid = False

    @property
    def w(self):
        w = np.zeros(self._N)
        if self._act:
            w[np.asarray(self._act, dtype=np.int64)] = np.asarray(self._aw)
        return w

    @w.setter
    def w(self, w):
        w = np.asarray(w, dtype=np.float64)
        nz = np.nonzero(w)[0]
        self._act = [int(i) for i in nz]
        self._aw = [float(w[i]) for i in nz]
        self._xw_valid = False

    def _scale_and_add(self, alpha, f, beta):
        """w = alpha*w ; w[f] = max(0, w[f] + beta)      (giga.py:63-64, frankwolfe.py:39-40)"""
        self._aw = [alpha*x for x in self._aw]
        f = int(f)
        if f in self._act:
            k = self._act.index(f)
            self._aw[k] = max(0., self._aw[k]+beta)
        else:
            self._act.append(f)
            self._aw.append(max(0., 0.*alpha+beta))
        self._xw_valid = False

    def _snapshot(self):
        return (list(self._act), list(self._aw))

    def _restore(self, snap):
        self._act, self._aw = list(snap[0]), list(snap[1])
        self._xw_valid = False

    def reset(self):
        self._reset_weights()
        self.reached_numeric_limit = False

    def size(self):
        return int(sum(1 for x in self._aw if x > 0))

    def weights(self):
        return self.w.copy()

    # ---- device steps ----
    def _iterate(self):
        """xw = A w on the device, from the active rows only"""
        if not self._xw_valid:
            m = len(self._act)
            idx = self._eng.upload(np.asarray(self._act, dtype=np.int64), dtype=torch.int64) if m else None
            aw = self._eng.upload(np.asarray(self._aw, dtype=np.float64)) if m else None
            nv.call('bc_dense_combine', self._ctx, ptr(self._V), self._ldv, self._S, ptr(idx), ptr(aw), m, ptr(self._xw), stream_ptr())
            self._xw_valid = True
        return self._xw

    def _vec(self, op, xf=None, aux=0.0, u=None, out=None, b=None):
        nv.call('bc_vec_step', self._ctx, op, ptr(self._iterate()), ptr(xf), ptr(self._b_dev if b is None else b), self._S, float(aux),
                ptr(u), ptr(self._out if out is None else out), stream_ptr())

    def _score(self, mode, u, active=None):
        nv.call('bc_dense_score', self._ctx, mode, ptr(self._V), self._N, self._S, self._ldv, ptr(self._norms), ptr(u), ptr(active), 0,
                ptr(self._out[4:]), None, stream_ptr())

    def _row(self, f):
        return self._V[int(f)]

    def error(self):
        if self._N == 0:
            return float(np.sqrt((self.b**2).sum()))
        self._vec(nv.VEC_RESID)
        return float(self._out[:1].cpu().numpy()[0])

    # ---- the reference's build loop (snnls.py:31-78) ----
    def build(self, itrs):
        if self.reached_numeric_limit:
            self.log.warning('the numeric limit was already reached; returning. size = %s, error = %s' % (self.size(), self.error()))
            return
        if self._N == 0 or self._S == 0:
            self.log.warning('there are no data, returning.')
            return
        retried_already = False
        for i in range(itrs):
            try:
                size_nonzero = self.size() > 0
                if self.check_error_monotone and size_nonzero:
                    prev_error = self.error()
                    prev_w = self._snapshot()
                f = self._select()
                self._reweight(f)
                if self.check_error_monotone and size_nonzero:
                    error = self.error()
                    if error > prev_error:
                        self._restore(prev_w)
                        raise NumericalPrecisionError('Error not monotone: curr error = %s prev error = %s' % (error, prev_error))
                    retried_already = False
            except NumericalPrecisionError as e:
                self.log.warning('numerical precision error: ' + str(e))
                if retried_already:
                    self.log.warning('iterative step failed a second time. Assuming numeric limit reached.')
                    self.reached_numeric_limit = True
                    break
                else:
                    self.log.warning('iterative step failed. Stabilizing and retrying...')
                    retried_already = True
                    self._stabilize()
            if self.reached_numeric_limit:
                break
        if self.reached_numeric_limit:
            self.log.warning('the numeric limit has been reached. No more points will be added. size = %s, error = %s'
                             % (self.size(), self.error()))

    def _active_columns(self, idx):
        """host (S, m) matrix of the columns idx (Lawson-Hanson runs on the host: S x m, tiny)"""
        idx = np.asarray(idx, dtype=np.int64)
        d_idx = self._eng.upload(idx, dtype=torch.int64)
        out = self._eng.empty(max(len(idx), 1), self._S)
        nv.call('bc_dense_gather', self._ctx, ptr(self._V), self._ldv, self._S, ptr(d_idx), len(idx), ptr(out), self._S, stream_ptr())
        return np.ascontiguousarray(out[:len(idx)].cpu().numpy().T)

    def _nnls_on(self, idx):
        res = nnls(self._active_columns(idx), self.b, maxiter=100*self._N)
        return res[0]

    # snnls.py:82-97
    def optimize(self):
        try:
            prev_cost = self.error()
            prev_w = self._snapshot()
            nz = sorted(i for i, x in zip(self._act, self._aw) if x > 0)
            sol = self._nnls_on(nz)
            lut = dict(zip(nz, sol))
            self._aw = [float(lut.get(i, x)) for i, x in zip(self._act, self._aw)]
            self._xw_valid = False
            new_cost = self.error()
            if new_cost > prev_cost*(1.+util.TOL):
                raise NumericalPrecisionError(
                    'self.optimize() returned a solution with increasing error. Numeric limit possibly reached: preverr = %s '
                    'err = %s. If the two errors are very close, try bc.util.set_tolerance(tol) with tol > current tol = %s'
                    % (prev_cost, new_cost, util.TOL))
        except NumericalPrecisionError as e:
            self.log.warning(e)
            self._restore(prev_w)
            self.reached_numeric_limit = True
            return

    def _stabilize(self):
        pass

    def _select(self):
        raise NotImplementedError

    def _reweight(self, f):
        raise NotImplementedError

    def _best(self):
        """(score, index) of the last score pass; plus the negative-direction pair for OrthoPursuit"""
        o = self._out.cpu().numpy()
        return o, float(o[4]), int(o[5:6].view(np.int64)[0]), float(o[6]), int(o[7:8].view(np.int64)[0])
