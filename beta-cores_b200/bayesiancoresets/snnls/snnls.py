"""SparseNNLS: greedy sparse non-negative least squares  min_w |A w - b|,  w >= 0, nnz(w) small
(drop-in for bayesiancoresets/snnls/snnls.py:8-106: same constructor, build loop, monotone-error check,
one retry, numeric-limit latch, optimize()).

Device layout: the reference's A is (S, N) with one column per datapoint; here the matrix lives in HBM
datapoint-major, V = A^T (N x S, row n contiguous), so that the per-iteration score pass streams each
datapoint's S values with coalesced 16-byte loads (bc_dense_score).

What an iteration touches:
  * the score pass reads V once (8 N S bytes, the HBM-bound part);
  * the rows that ever received weight are CACHED in a small replicated matrix (m x S): the iterate A w is
    re-formed from those m rows on the device (bc_dense_combine) -- exact, so the reference's strict error-monotone
    check (snnls.py:56-62) sees the errors it would see, at O(m S) instead of the reference's three N x S products per
    iteration; the weights go up through a pinned buffer without a synchronisation;
  * the S-length line-search algebra is one single-CTA kernel per step (bc_vec_step);
  * the host reads three small results per iteration (selection, step scalars, new error) and applies the
    reference's guards to them; the error of the current weights is remembered, not recomputed.

Datapoints (the rows of V) may be sharded over torch.distributed ranks (last sentence): every rank scores
its own block, the ranks exchange one (score, global index) pair per iteration (two for OrthoPursuit) and merge them
with numpy's arg-max rule; a newly selected row is broadcast once by its owner into every rank's cache, after which
the iterate, the line search and the error are computed redundantly and identically everywhere.
"""
import bisect
import logging
import secrets
import numpy as np
import torch
from scipy.optimize import nnls

from .. import util
from .. import _native as nv
from .._device import Engine, ptr, stream_ptr
from .._shard import Comm, merge_best, partition_rows
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
        self._comm = Comm.current()
        A = np.asarray(A, dtype=np.float64)
        self._S, self._N = (A.shape[0], A.shape[1]) if A.ndim == 2 else (0, 0)
        self._nl, self._row0 = self._N, 0
        if A.size:
            # in a multi-rank job every rank is handed the same host matrix and keeps its own block of columns (= datapoints)
            r0, nl = partition_rows(self._N, self._comm.world, self._comm.rank) if self._comm.world > 1 else (0, self._N)
            V = self._eng.empty(max(nl, 1), self._S)
            if nl:
                At = self._eng.upload(np.ascontiguousarray(A[:, r0:r0+nl]))          # (S, nl) as the caller laid it out
                nv.call('bc_transpose', self._ctx, ptr(At), self._S, nl, int(At.stride(0)), ptr(V), self._S, stream_ptr())
                del At
            self._attach(V[:nl], None, r0, self._N)
        self._reset_weights()

    @classmethod
    def from_device(cls, V, norms, b_host, row0=0, n_total=None):
        """build a solver directly on a device-resident datapoint-major matrix V (n_local x S), e.g. the output of the
        materialise pass (HilbertCoreset); rows [row0, row0 + n_local) of n_total in a sharded job."""
        self = cls.__new__(cls)
        self.alg_name = '%s-%s' % (cls.__name__, secrets.token_hex(3))
        self.log = logging.LoggerAdapter(logging.getLogger(), {'id': self.alg_name})
        self.A = None
        self.b = np.asarray(b_host, dtype=np.float64)
        self.reached_numeric_limit = False
        self.check_error_monotone = True
        self._eng = Engine.get()
        self._ctx = self._eng.ctx()
        self._comm = Comm.current()
        self._S = int(V.shape[1])
        self._N = int(n_total) if n_total is not None else int(V.shape[0])
        self._nl, self._row0 = int(V.shape[0]), int(row0)
        if self._N:
            self._attach(V, norms, row0, self._N)
        self._reset_weights()
        self._setup()
        return self

    def _attach(self, V, norms, row0, n_total):
        self._V = V
        self._nl, self._row0 = int(V.shape[0]), int(row0)
        self._ldv = int(V.stride(0)) if self._nl else self._S
        if norms is None:
            norms = self._eng.empty(max(self._nl, 1))
            if self._nl:
                nv.call('bc_dense_rownorms', self._ctx, ptr(V), self._nl, self._S, self._ldv, ptr(norms), stream_ptr())
        self._norms = norms
        # all N norms on every host, once: the zero-column check and Frank-Wolfe's sum of norms (frankwolfe.py:10)
        if self._comm.world > 1:
            parts = self._comm.allgather_var(norms[:self._nl])
            self._starts = list(np.cumsum([0] + [int(p.numel()) for p in parts])[:-1])
            self._norms_host = torch.cat(parts).cpu().numpy()
            if self._starts[self._comm.rank] != self._row0 or self._norms_host.shape[0] != n_total:
                raise ValueError('row blocks of the ranks do not tile [0, %d)' % n_total)
        else:
            self._starts = [0]
            self._norms_host = norms[:self._nl].cpu().numpy()
        self._b_dev = self._eng.upload(self.b)
        self._xw = self._eng.zeros(self._S)
        self._u = self._eng.zeros(2*self._S)
        self._out = self._eng.zeros(8)
        self._cache_cap = 0
        self._Vact = None
        self._aw_dev = None
        self._seq_dev = None
        self._pin = [torch.empty(1, dtype=torch.float64).pin_memory(), torch.empty(1, dtype=torch.float64).pin_memory()]
        self._pin_k = 0

    def _setup(self):
        pass

    def _require_nonzero_columns(self):
        if self._N and np.any(self._norms_host == 0):
            raise ValueError(self.alg_name+'.__init__(): A must not have any 0 columns')

    # ---- weights: sparse on the inside, the reference's dense `w` on the outside ----
    def _reset_weights(self):
        self._act = []            # global indices that ever received weight, in order of first selection
        self._pos = {}            # index -> position in _act (= row of the cache)
        self._aw = []             # their weights
        self._xw_valid = False
        self._err = None          # error() of the current weights, once computed

    @property
    def w(self):
        w = np.zeros(self._N)
        if self._act:
            w[np.asarray(self._act, dtype=np.int64)] = np.asarray(self._aw)
        return w

    @w.setter
    def w(self, w):
        w = np.asarray(w, dtype=np.float64)
        self._reset_weights()
        for i in np.nonzero(w)[0]:
            self._activate(int(i))
            self._aw[self._pos[int(i)]] = float(w[i])
        self._touch()

    def _touch(self):
        self._xw_valid = False
        self._err = None

    def _owner(self, f):
        return bisect.bisect_right(self._starts, f) - 1

    def _activate(self, f):
        """make sure datapoint f has a row in the replicated cache (fetched once from its owner) and a weight slot"""
        f = int(f)
        if f in self._pos:
            return self._pos[f]
        k = len(self._act)
        if k >= self._cache_cap:
            cap = max(64, 2*self._cache_cap)
            Vn = self._eng.empty(cap, self._S)
            if self._Vact is not None and k:
                Vn[:k].copy_(self._Vact[:k])
            self._Vact, self._cache_cap = Vn, cap
            self._aw_dev = self._eng.zeros(cap)
            self._seq_dev = torch.arange(cap, dtype=torch.int64, device=self._eng.device)
            self._pin = [torch.empty(cap, dtype=torch.float64).pin_memory(), torch.empty(cap, dtype=torch.float64).pin_memory()]
        own = self._owner(f) if self._comm.world > 1 else self._comm.rank
        if own == self._comm.rank:
            self._Vact[k].copy_(self._V[f-self._row0])
        if self._comm.world > 1:
            self._comm.broadcast(self._Vact[k], own)
        self._act.append(f)
        self._pos[f] = k
        self._aw.append(0.)
        return k

    def _scale_and_add(self, alpha, f, beta):
        """w = alpha*w ; w[f] = max(0, w[f] + beta)      (giga.py:63-64, frankwolfe.py:39-40)"""
        k = self._activate(f)
        self._aw = [alpha*x for x in self._aw]
        self._aw[k] = max(0., self._aw[k]+beta)
        self._touch()

    def _snapshot(self):
        return (list(self._aw), self._err)

    def _restore(self, snap):
        # rows selected since the snapshot stay cached (with weight 0, exactly the reference's w after `self.w = prev_w`)
        self._aw = list(snap[0]) + [0.]*(len(self._act)-len(snap[0]))
        self._touch()
        self._err = snap[1]

    def reset(self):
        self._reset_weights()
        self.reached_numeric_limit = False

    def size(self):
        return int(sum(1 for x in self._aw if x > 0))

    def weights(self):
        return self.w.copy()

    # ---- device steps ----
    def _iterate(self):
        """xw = A w on the device, re-formed exactly from the cached active rows"""
        if not self._xw_valid:
            m = len(self._act)
            if m:
                self._pin_k ^= 1
                pin = self._pin[self._pin_k]
                pin[:m] = torch.from_numpy(np.asarray(self._aw, dtype=np.float64))
                self._aw_dev[:m].copy_(pin[:m], non_blocking=True)
                nv.call('bc_dense_combine', self._ctx, ptr(self._Vact), self._S, self._S, ptr(self._seq_dev), ptr(self._aw_dev), m,
                        ptr(self._xw), stream_ptr())
            else:
                self._xw.zero_()
            self._xw_valid = True
        return self._xw

    def _vec(self, op, xf=None, aux=0.0, u=None, out=None, b=None):
        nv.call('bc_vec_step', self._ctx, op, ptr(self._iterate()), ptr(xf), ptr(self._b_dev if b is None else b), self._S, float(aux),
                ptr(u), ptr(self._out if out is None else out), stream_ptr())

    def _score(self, mode, u, active=None):
        if self._nl:
            nv.call('bc_dense_score', self._ctx, mode, ptr(self._V), self._nl, self._S, self._ldv, ptr(self._norms), ptr(u), ptr(active),
                    self._row0, ptr(self._out[4:]), None, stream_ptr())
        else:       # a rank without datapoints offers no candidate (index -1 loses every merge)
            empty = np.zeros(4)
            empty[1::2] = np.array([-1, -1], dtype=np.int64).view(np.float64)
            self._out[4:].copy_(torch.from_numpy(empty))

    def _row(self, f):
        k = self._activate(f)          # may (re)allocate the cache: look the tensor up afterwards
        return self._Vact[k]

    def error(self):
        if self._N == 0:
            return float(np.sqrt((self.b**2).sum()))
        if self._err is None:
            self._vec(nv.VEC_RESID)
            self._err = float(self._out[:1].cpu().numpy()[0])
        return self._err

    # ---- runs of iterations without a host round trip (GIGA, Frank-Wolfe; one rank) ----
    _device_algo = None            # bc_solver_iterations' algo id; None: this solver iterates on the host path only
    DEVICE_RUN_MAX = 256           # iterations queued per run (each is four launches; the state comes back once per run)

    def _device_run_operands(self):
        """(right-hand side of the line search, aux) of bc_solver_iterations for this solver"""
        raise NotImplementedError

    def _device_iterations(self, want):
        """queue up to `want` whole iterations on the device (bc_solver_iterations) and adopt the state they leave behind.
        Returns (iterations completed, a monotone-checked iteration completed, a guard stopped the run)."""
        import os
        if (self._device_algo is None or self._comm.world > 1 or self._nl == 0 or os.environ.get('BC_SOLVER_DEVICE_LOOP', '1') == '0'):
            return 0, False, True
        want = min(int(want), self.DEVICE_RUN_MAX)
        eng, S = self._eng, self._S
        m0 = len(self._act)
        # room in the row cache for every datapoint the run could add
        need = m0 + want
        if need > self._cache_cap:
            cap = max(64, 2*self._cache_cap, need)
            Vn = eng.empty(cap, S)
            if self._Vact is not None and m0:
                Vn[:m0].copy_(self._Vact[:m0])
            self._Vact, self._cache_cap = Vn, cap
            self._aw_dev = eng.zeros(cap)
            self._seq_dev = torch.arange(cap, dtype=torch.int64, device=eng.device)
            self._pin = [torch.empty(cap, dtype=torch.float64).pin_memory(), torch.empty(cap, dtype=torch.float64).pin_memory()]
        cap = self._cache_cap
        st = getattr(self, '_run_state', None)
        if st is None or st['cap'] != cap:
            st = self._run_state = {'cap': cap, 'dev': eng.empty(8 + 2*cap), 'prev': eng.empty(cap),
                                    'host': torch.empty(8 + 2*cap, dtype=torch.float64).pin_memory()}
        size_nonzero = self.size() > 0
        err = self.error() if size_nonzero else 0.          # (remembered, not recomputed, once an iteration has run)
        self._iterate()                                     # A w of the current weights in self._xw
        h = st['host'].numpy()
        h[:8] = [m0, 0., 0., 0., err, 1. if self.check_error_monotone else 0., cap, 0.]
        h[8:8+m0] = self._aw
        h[8+cap:8+cap+m0] = np.asarray(self._act, dtype=np.int64).view(np.float64)
        st['dev'].copy_(st['host'], non_blocking=True)
        ctl, aw, act = st['dev'][:8], st['dev'][8:8+cap], st['dev'][8+cap:]
        rhs, aux = self._device_run_operands()
        nv.call('bc_solver_iterations', self._ctx, self._device_algo, want, ptr(self._V), self._nl, S, self._ldv, ptr(self._norms),
                ptr(rhs), ptr(self._b_dev), float(aux), float(util.TOL), ptr(self._Vact), ptr(ctl), ptr(aw), ptr(st['prev']), ptr(act),
                ptr(self._xw), ptr(self._u), ptr(self._out), stream_ptr())
        st['host'].copy_(st['dev'])                         # the one read-back of the run (synchronises)
        m, status, done, checked = int(h[0]), int(h[1]), int(h[2]), h[3] != 0.
        self._act = [int(v) for v in h[8+cap:8+cap+m].view(np.int64)]
        self._pos = {f: k for k, f in enumerate(self._act)}
        self._aw = [float(v) for v in h[8:8+m]]
        # every exit leaves A w of the adopted weights in self._xw: a guard trips before the weights change, the monotone check
        # restores them and re-forms A w
        self._xw_valid = True
        self._err = float(h[4]) if any(x > 0 for x in self._aw) else None
        return done, checked, status != 0

    # ---- the reference's build loop (snnls.py:31-78) ----
    def build(self, itrs):
        if self.reached_numeric_limit:
            self.log.warning('the numeric limit was already reached; returning. size = %s, error = %s' % (self.size(), self.error()))
            return
        if self._N == 0 or self._S == 0:
            self.log.warning('there are no data, returning.')
            return
        retried_already = False
        i = 0
        while i < itrs:
            # whole iterations on the device for as long as no guard trips; the iteration that trips one is then taken through
            # the per-iteration path below, where the reference's error handling lives
            done, checked, stopped = self._device_iterations(itrs - i)
            i += done
            if checked:
                retried_already = False
            if not stopped or i >= itrs:
                continue
            i += 1
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
        """host (S, m) matrix of the columns idx, from the replicated cache (Lawson-Hanson runs on the host: S x m, tiny)"""
        ks = [self._activate(int(i)) for i in idx]
        if not ks:
            return np.zeros((self._S, 0))
        sel = self._Vact[torch.as_tensor(ks, dtype=torch.int64, device=self._eng.device)]
        return np.ascontiguousarray(sel.cpu().numpy().T)

    def _nnls_on(self, idx, fresh=None):
        """argmin |A[:, idx] x - b|, x >= 0 (orthopursuit.py:40, snnls.py:88).  On the device (bc_nnls: Lawson-Hanson on the
        cached rows, warm-started from the current weights -- `fresh` names an index that has only just been marked for
        inclusion and starts at 0); scipy's routine on the host when the block is wider than the kernel takes, when the
        kernel reports numerically dependent columns, or with BC_DEVICE_NNLS=0.  The reference pins no scipy version for this
        path; the two agree to rounding (tests/test_gpu_parity.py::test_device_nnls_matches_scipy)."""
        import os
        ks = [self._activate(int(i)) for i in idx]
        m = len(ks)
        if m and m <= nv.lib().bc_nnls_max_columns() and os.environ.get('BC_DEVICE_NNLS', '1') != '0':
            eng = self._eng
            pin = getattr(self, '_nnls_pin', None)          # one pinned staging buffer per solver (pinning per call costs ~0.1 ms)
            if pin is None or pin.numel() < 2*m:
                pin = self._nnls_pin = torch.empty(max(2*m, 256), dtype=torch.float64).pin_memory()
                self._nnls_pin_ev = None
            if self._nnls_pin_ev is not None:
                self._nnls_pin_ev.synchronize()             # the upload that last read the buffer has run
            buf = pin[:2*m]
            h = buf.numpy()
            h[:m] = np.asarray(ks, dtype=np.int64).view(np.float64)
            h[m:] = [0. if (fresh is not None and int(i) == int(fresh)) else max(0., self._aw[k]) for i, k in zip(idx, ks)]
            d = buf.to(eng.device, non_blocking=True)
            self._nnls_pin_ev = torch.cuda.Event()
            self._nnls_pin_ev.record()
            out = eng.empty(m + 1)
            info = out[m:].view(torch.int32)
            nv.call('bc_nnls', self._ctx, ptr(self._Vact), self._S, ptr(d[:m]), m, ptr(self._b_dev), ptr(d[m:]), ptr(out), 0, ptr(info),
                    stream_ptr())
            o = out.cpu()
            if int(o[m:].view(torch.int32)[0]) == 0:
                return o[:m].numpy().copy()
            self.log.warning('device nnls did not converge (status %d); solving on the host' % int(o[m:].view(torch.int32)[0]))
        res = nnls(self._active_columns(idx), self.b, maxiter=100*self._N)
        return res[0]

    def _assign(self, idx, sol):
        lut = dict(zip(idx, sol))
        self._aw = [float(lut.get(i, x)) for i, x in zip(self._act, self._aw)]
        self._touch()

    # snnls.py:82-97
    def optimize(self):
        try:
            prev_cost = self.error()
            prev_w = self._snapshot()
            nz = sorted(i for i, x in zip(self._act, self._aw) if x > 0)
            self._assign(nz, self._nnls_on(nz))
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
        """(out, score, index, neg-score, neg-index) of the last score pass, merged over the ranks with numpy's arg-max rule;
        out[0:4] = the scalars of the last bc_vec_step (identical on every rank)"""
        if self._comm.world == 1:
            o = self._out.cpu().numpy()
        else:
            allo = self._comm.allgather(self._out).cpu().numpy()
            o = allo[self._comm.rank].copy()
            bi = allo[:, 5:6].copy().view(np.int64)[:, 0]
            ni = allo[:, 7:8].copy().view(np.int64)[:, 0]
            bv, bidx = merge_best((allo[r, 4], int(bi[r])) for r in range(allo.shape[0]))
            nvv, nidx = merge_best((allo[r, 6], int(ni[r])) for r in range(allo.shape[0]))
            return o, float(bv), int(bidx), float(nvv), int(nidx)
        return o, float(o[4]), int(o[5:6].view(np.int64)[0]), float(o[6]), int(o[7:8].view(np.int64)[0])
