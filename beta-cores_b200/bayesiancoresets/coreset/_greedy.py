"""Greedy variational coreset construction on the device: the loop BetaCoreset and SparseVICoreset
share (bayesiancoresets/coreset/bcores.py:27-150, sparsevi.py:26-136; individual-point mode).

Per `build` iteration:
  _select   : sampler -> Theta (host callback) ; pass A over the data rows = column sum of the centred
              projection (bc_project_colsum) ; coreset rows materialised (M x S) ; residual ; pass B over the
              data rows = per-row correlation + arg-max (bc_project_score) ; gate against the coreset rows.
  _optimize : opt_itrs x [sampler -> Theta ; pass A ; coreset rows ; residual ; gradient ; ADAM step on device ;
              weights back to the host for the next sampler call].
The (n, S) matrix of the data rows is never formed when the projector's likelihood is a bound
DevicePotential ("fused" tangent).  Opaque Python callbacks take the "dense" tangent: the callback's host
matrix is uploaded and the same reductions run on it with the materialised-matrix kernels.

Rows may be sharded over torch.distributed ranks (DeviceRows.row0 / n_total): per step the ranks exchange one
2 x (S+1) double-double part, per selection one (score, position) pair (_shard.py).
"""
import ctypes
import numpy as np
import torch

from .. import _native as nv
from .._device import Engine, DeviceRows, ptr, stream_ptr, padded_ld
from .._shard import Comm, local_subsample, merge_best, owner_of
from ..util.opt import nn_opt, partial_nn_opt
from ..util import rng
from .coreset import Coreset


def _as_f64_2d(a):
    return np.atleast_2d(np.asarray(a, dtype=np.float64))


# one library call per optimiser step (bc_greedy_opt_step) instead of one per kernel; BC_FUSED_STEP=0 keeps the call-by-call loop
import os
import time
FUSED_STEP_CALL = os.environ.get('BC_FUSED_STEP', '1') != '0'
# samplers that offer device_step() run the optimiser loop without any host synchronisation; BC_DEVICE_SAMPLER_LOOP=0 keeps
# calling them through the reference's host protocol sampler(S, wts, pts)
DEVICE_SAMPLER_LOOP = os.environ.get('BC_DEVICE_SAMPLER_LOOP', '1') != '0'
# single-rank host-free loops replay the device work of a step as a CUDA graph (two graphs, one per staging-buffer parity)
# once the loop is long enough to pay for the two captures (about 3 ms against 30-40 us saved per step: measured on the
# Gaussian example, 1000 steps per point, 0.157 -> 0.125 ms per step; the 20-step loops of the logistic / neural-linear
# examples lose); BC_STEP_GRAPH=0 keeps launching kernel by kernel
# instrumentation (bench.py): a list that collects, per optimiser step of a row-sharded job, four CUDA events -- step begin,
# first half queued, exchange done, second half done -- for the per-step overhead breakdown; None = off
STEP_EVENTS = None
STEP_GRAPH = os.environ.get('BC_STEP_GRAPH', '1') != '0'
STEP_GRAPH_MIN_ITRS = int(os.environ.get('BC_STEP_GRAPH_MIN_ITRS', '200'))


class _Tangent(object):
    """Common device-side algebra once a column sum / coreset-row matrix is available."""

    def __init__(self, owner):
        self.o = owner
        self.eng = Engine.get()
        self.ctx = self.eng.ctx()
        self.comm = Comm.current()
        self.resid = None
        self.out4 = self.eng.zeros(4)
        self.core_out = self.eng.zeros(1)      # max |correlation| of the coreset rows (kept apart from out4: the dense
                                               # score kernel writes all four entries of its output)

    # residual = scaling * colsum - w . Vc      (bcores.py:77 / :145); element S = its sum
    def residual(self, colsum, scaling, Vc, w_dev):
        S = colsum.shape[0]
        if self.resid is None or self.resid.shape[0] != S+1:
            self.resid = self.eng.empty(S+1)
        M = 0 if Vc is None else Vc.shape[0]
        nv.call('bc_core_resid', self.ctx, ptr(colsum), float(scaling), ptr(Vc), M, S, S if Vc is None else int(Vc.stride(0)),
                ptr(w_dev), ptr(self.resid), stream_ptr())
        return self.resid

    def grad(self, Vc, resid, out):
        nv.call('bc_core_grad', self.ctx, ptr(Vc), Vc.shape[0], Vc.shape[1], int(Vc.stride(0)), ptr(resid), ptr(out), stream_ptr())
        return out

    def core_max(self, Vc, resid, skip, out):
        nv.call('bc_core_maxcorr', self.ctx, ptr(Vc), Vc.shape[0], Vc.shape[1], int(Vc.stride(0)), ptr(resid), int(skip), ptr(out),
                stream_ptr())


    # group-wise selection (bcores.py:91-123): `gv` = one row per candidate group (sum of the group's centred rows)
    def group_colsum(self, gv):
        G, S = gv.shape
        dd = self.eng.empty(2*(S+1))
        nv.call('bc_dense_colsum', self.ctx, ptr(gv), G, S, int(gv.stride(0)), ptr(dd), stream_ptr())
        out = self.eng.empty(S)
        nv.call('bc_colsum_combine', self.ctx, ptr(dd), 1, S, ptr(out), stream_ptr())
        return out

    def group_best(self, gv, resid):
        G, S = gv.shape
        nv.call('bc_dense_score', self.ctx, nv.SCORE_CORR, ptr(gv), G, S, int(gv.stride(0)), None, ptr(resid), None, 0,
                ptr(self.out4), None, stream_ptr())
        v = self.out4[:2].cpu().numpy()
        return float(v[0]), int(v[1:2].view(np.int64)[0])


class _FusedTangent(_Tangent):
    """likelihood = bound DevicePotential: everything stays on the device, nothing of size n x S exists."""

    def __init__(self, owner, fused, rows):
        super().__init__(owner)
        self.fp = fused
        self.rows = rows
        self.parts = None
        self.colsum_buf = None
        self._theta_buf = None
        self._theta_mode = None        # None (unchecked) / 'replicated' / 'broadcast'
        self._theta_calls = 0
        self._check_next = False

    def begin(self, w, p, beta):
        prj = self.o.ll_projector
        prj.update(w, p)                                   # host sampler: consumes np.random exactly like the reference
        if isinstance(prj.samples, torch.Tensor):          # device-side sampler: the samples are already in HBM
            th = prj.samples.to(device=self.eng.device, dtype=torch.float64)
            if self._theta_buf is None or tuple(self._theta_buf.shape) != tuple(th.shape):
                self._theta_buf = self.eng.empty(*th.shape)
            self._theta_buf.copy_(th)
        else:
            th = np.ascontiguousarray(np.atleast_2d(np.asarray(prj.samples, dtype=np.float64)))
            if self._theta_buf is None or tuple(self._theta_buf.shape) != th.shape:
                self._theta_buf = self.eng.empty(*th.shape)
            self._theta_buf.copy_(torch.from_numpy(th))
        self._replicate_samples()
        self.fp.configure(beta)
        self.fp.set_samples(self._theta_buf)
        self.S = self.fp.S

    def _replicate_samples(self):
        """Every rank runs the user's sampler on the same weights and the same numpy stream, so the samples are normally
        bit-identical already and a per-step 8 S D-byte broadcast would only add latency to the optimiser loop.  That is
        CHECKED, not assumed: the ranks compare a 64-bit checksum of the sample bits (one 8-byte all-gather) on the first
        call and then at every selection; any mismatch switches this object to broadcasting rank 0's samples for good."""
        if self.comm.world == 1:
            return
        if self._theta_mode == 'broadcast':
            self.comm.broadcast(self._theta_buf, 0)
            return
        self._theta_calls += 1
        if self._theta_mode is None or self._check_next:
            self._check_next = False
            h = self._theta_buf.view(torch.int64).sum().reshape(1)
            allh = self.comm.allgather(h).cpu()
            if bool((allh != allh[0]).any()):
                import logging
                logging.getLogger(__name__).warning('posterior samples differ between ranks (unseeded sampler?): broadcasting rank 0\'s every step')
                self._theta_mode = 'broadcast'
                self.comm.broadcast(self._theta_buf, 0)
            else:
                self._theta_mode = 'replicated'

    def check_samples_at_next_call(self):
        self._check_next = True

    def _sub_local(self, sub_idcs):
        if sub_idcs is None:
            return None, None
        pos, loc = local_subsample(sub_idcs, self.rows.row0, self.rows.n_local)
        return pos, self.eng.upload(loc, dtype=torch.int64)

    def colsum(self, sub_idcs):
        """column sum of the centred projection of the (sub-sampled) data rows, all ranks combined"""
        Sld = self.fp.Sld
        if self.parts is None or self.parts.shape[0] != 2*Sld:
            self.parts = self.eng.empty(2*Sld)
            self.colsum_buf = self.eng.empty(self.S)
        self._pos, self._loc = self._sub_local(sub_idcs)
        self.fp.colsum_parts(self.rows, self._loc, out=self.parts)
        allparts = self.comm.allgather(self.parts)
        return self.fp.combine(allparts, self.comm.world, out=self.colsum_buf)

    def core_rows(self, pts_rows):
        V, _, _ = self.fp.materialise(pts_rows)
        return V

    def group_rows(self, groups, gi):
        """(len(gi), S) device matrix: row j = column sum of the centred projection of the rows of group gi[j]
        (bcores.py:50,60) -- one gathered fused pass per group, nothing of size rows x S is formed; all ranks combined"""
        Sld, S = self.fp.Sld, self.S
        if not hasattr(self, '_group_loc'):
            self._group_loc = {}
        parts = self.eng.empty(len(gi), 2*Sld)
        for j, g in enumerate(gi):
            g = int(g)
            if g not in self._group_loc:
                _, loc = local_subsample(np.asarray(groups[g], dtype=np.int64), self.rows.row0, self.rows.n_local)
                self._group_loc[g] = self.eng.upload(loc, dtype=torch.int64)
            self.fp.colsum_parts(self.rows, self._group_loc[g], out=parts[j])
        allparts = self.comm.allgather(parts)                       # (world, G', 2 Sld)
        gv = self.eng.empty(len(gi), S)
        for j in range(len(gi)):
            self.fp.combine(allparts[:, j].contiguous(), self.comm.world, out=gv[j])
        return gv

    def best_row(self, sub_idcs, resid):
        """(score, position) of np.argmax(corrs) over the rows of the last colsum() call"""
        n_here = self.rows.n_local if sub_idcs is None else int(self._loc.numel())
        out = self.out4
        if n_here > 0:
            self.fp.score(self.rows, self._loc, resid, self.rows.row0 if sub_idcs is None else 0, out)
            v = out[:2].cpu().numpy()
            score = float(v[0])
            p = int(v[1:2].view(np.int64)[0])
            if sub_idcs is not None:
                p = int(self._pos[p])                      # position in the global subsample list
        else:
            score, p = 0.0, -1
        if self.comm.world == 1:
            return score, p
        mine = torch.tensor([score, float(p)], dtype=torch.float64, device=self.eng.device)
        allc = self.comm.allgather(mine).cpu().numpy()
        return merge_best((allc[r, 0], int(allc[r, 1])) for r in range(allc.shape[0]))


class _DenseTangent(_Tangent):
    """opaque Python likelihood callbacks: the projector's centred (n, S) host matrix is uploaded and reduced
    with the materialised-matrix kernels (no sharding: the callback sees whole host arrays)."""

    def __init__(self, owner, project):
        super().__init__(owner)
        self.project = project          # (pts) -> centred (n, S) host array, through the projector
        if self.comm.world > 1:
            raise NotImplementedError('row sharding needs a DevicePotential likelihood (opaque callbacks see host arrays)')

    def begin(self, w, p, beta):
        self.o.ll_projector.update(w, p)
        self.beta = beta

    def colsum(self, sub_idcs):
        data = self.o.data
        vecs = self.project(data if sub_idcs is None else data[sub_idcs])
        self.V = self.eng.upload(_as_f64_2d(vecs))
        n, S = self.V.shape
        self.S = S
        dd = self.eng.empty(2*(S+1))
        nv.call('bc_dense_colsum', self.ctx, ptr(self.V), n, S, int(self.V.stride(0)), ptr(dd), stream_ptr())
        out = self.eng.empty(S)
        nv.call('bc_colsum_combine', self.ctx, ptr(dd), 1, S, ptr(out), stream_ptr())
        return out

    def core_rows(self, pts_host):
        return self.eng.upload(_as_f64_2d(self.project(pts_host)))

    def group_rows(self, groups, gi):
        data = self.o.data
        first = self.eng.upload(_as_f64_2d(self.project(data[groups[int(gi[0])], :])))
        S = first.shape[1]
        self.S = S
        gv = self.eng.empty(len(gi), S)
        dd = self.eng.empty(2*(S+1))
        for j, g in enumerate(gi):
            V = first if j == 0 else self.eng.upload(_as_f64_2d(self.project(data[groups[int(g)], :])))
            nv.call('bc_dense_colsum', self.ctx, ptr(V), V.shape[0], S, int(V.stride(0)), ptr(dd), stream_ptr())
            nv.call('bc_colsum_combine', self.ctx, ptr(dd), 1, S, ptr(gv[j]), stream_ptr())
        return gv

    def best_row(self, sub_idcs, resid):
        n, S = self.V.shape
        nv.call('bc_dense_score', self.ctx, nv.SCORE_CORR, ptr(self.V), n, S, int(self.V.stride(0)), None, ptr(resid), None, 0,
                ptr(self.out4), None, stream_ptr())
        v = self.out4[:2].cpu().numpy()
        return float(v[0]), int(v[1:2].view(np.int64)[0])


class GreedyVICoreset(Coreset):
    """shared implementation; subclasses set `_uses_beta` and the projector method names"""
    _uses_beta = False

    def _init_greedy(self, data, ll_projector, n_subsample_select, n_subsample_opt, opt_itrs, step_sched, groups, initialized, kw):
        self.ll_projector = ll_projector
        if isinstance(data, DeviceRows):
            self.rows = data
            self.data = data                     # rows fetched on demand (data[f] -> DeviceRows.__getitem__)
            n_total, ncols = data.n_total, data.ncols
        else:
            self.rows = None
            self.data = data
            n_total, ncols = data.shape[0], data.shape[1]
        self._n_total, self._ncols = n_total, ncols
        self.n_subsample_select = None if n_subsample_select is None else min(n_total, n_subsample_select)
        self.n_subsample_opt = None if n_subsample_opt is None else min(n_total, n_subsample_opt)
        self.step_sched = step_sched
        self.opt_itrs = opt_itrs
        self.groups = groups
        self.selected_groups = []
        self._groups_cover = None
        if groups is not None:
            flat = np.array([r for g in groups for r in g], dtype=np.int64)
            # when the groups list every row exactly once, in order, a pass over "all grouped rows" is a full-block pass
            self._groups_cover = flat.shape[0] == n_total and bool(np.all(flat == np.arange(n_total)))
            self._groups_flat = flat
        Coreset.__init__(self, **kw)
        self.initialized = int(initialized)*len(self.wts)
        self._tangent = None

    # ---- tangent-space back-end ----
    def _host_project(self, pts):
        raise NotImplementedError

    def _beta(self):
        return None

    def _get_tangent(self):
        if self._tangent is None:
            fused = self.ll_projector.fused(self._ncols) if hasattr(self.ll_projector, 'fused') else None
            if fused is not None:
                eng = Engine.get()
                if self.rows is None:
                    comm = Comm.current()
                    if comm.world > 1:
                        from .._shard import partition_rows
                        r0, nl = partition_rows(self._n_total, comm.world, comm.rank)
                        self.rows = DeviceRows(eng, self.data[r0:r0+nl], row0=r0, n_total=self._n_total)
                    else:
                        self.rows = DeviceRows(eng, self.data)
                self._tangent = _FusedTangent(self, fused, self.rows)
            else:
                if self.rows is not None and not isinstance(self.data, np.ndarray):
                    raise TypeError('DeviceRows data need a DevicePotential likelihood')
                self._tangent = _DenseTangent(self, self._host_project)
        return self._tangent

    def _core_operand(self, t):
        """coreset points in the form the tangent's core_rows() takes"""
        if isinstance(t, _FusedTangent):
            return DeviceRows(t.eng, self.pts.reshape(-1, self._ncols))
        return self.pts

    def _build(self, itrs, sz):
        if self.groups is None and self.size()+itrs > sz:      # bcores.py:28-30: group mode has no size guard
            raise ValueError('%s._build(): # itrs + current size cannot exceed total desired size sz. # itr = %s cur sz: %s '
                             'desired sz: %s' % (self.alg_name, itrs, self.size(), sz))
        for i in range(itrs):
            self._select()
            self._optimize()

    # bcores.py:91-123 / sparsevi.py:93-126
    def _select_group(self):
        t = self._get_tangent()
        t.begin(self.wts, self.pts, self._beta())
        G = len(self.groups)
        if self.n_subsample_select is None:
            gi, scaling = list(range(G)), 1.
        else:
            gi = rng.randint(G, self.n_subsample_select)                       # bcores.py:57
            scaling = G/self.n_subsample_select
        gv = t.group_rows(self.groups, gi)
        colsum = t.group_colsum(gv)
        if self.pts.size > 0:
            Vc = t.core_rows(self._core_operand(t))
            w_dev = t.eng.upload(self.wts)
            resid = t.residual(colsum, scaling, Vc, w_dev)
            M = Vc.shape[0]
            if M > self.initialized:
                t.core_max(Vc, resid, self.initialized, t.core_out)
        else:
            Vc = None
            resid = t.residual(colsum, scaling, None, None)
        best, pos = t.group_best(gv, resid)
        if Vc is None:
            take = True
        elif Vc.shape[0] > self.initialized:
            take = best > float(t.core_out.cpu().numpy()[0])           # NaN on either side -> False
        else:
            take = best > -np.inf                                       # bcores.py:107-108 (False for NaN)
        if take and pos >= 0:
            f = int(pos) if self.n_subsample_select is None else int(gi[pos])
            if f not in self.selected_groups:
                self.selected_groups.append(f)
                rows_f = np.asarray(self.groups[f], dtype=np.int64)
                new = np.vstack([np.asarray(self.data[int(r)], dtype=np.float64).reshape(1, -1) for r in rows_f]) \
                    if self.rows is not None and not isinstance(self.data, np.ndarray) else np.asarray(self.data[rows_f, :], dtype=np.float64)
                self.wts = np.append(self.wts, np.zeros(new.shape[0]))
                self.idcs = np.append(self.idcs, rows_f).astype(np.int64)
                self.pts = np.vstack((self.pts.reshape(-1, self._ncols), new))

    # bcores.py:75-90 / sparsevi.py:73-92
    def _select(self):
        if self.groups is not None:
            return self._select_group()
        t = self._get_tangent()
        M = self.wts.shape[0]
        if hasattr(t, 'check_samples_at_next_call'):
            t.check_samples_at_next_call()
        t.begin(self.wts, self.pts, self._beta())
        if self.n_subsample_select is None:
            sub_idcs, scaling = None, 1.
        else:
            sub_idcs = rng.randint(self._n_total, self.n_subsample_select)      # bcores.py:53
            scaling = self._n_total/self.n_subsample_select
        colsum = t.colsum(sub_idcs)
        if self.pts.size > 0:
            Vc = t.core_rows(self._core_operand(t))
            w_dev = t.eng.upload(self.wts)
            resid = t.residual(colsum, scaling, Vc, w_dev)
            t.core_max(Vc, resid, 0, t.core_out)
        else:
            Vc = None
            resid = t.residual(colsum, scaling, None, None)
        best, pos = t.best_row(sub_idcs, resid)
        if Vc is not None:
            core_best = float(t.core_out.cpu().numpy()[0])
            take = best > core_best                          # NaN on either side -> False, like `corrs.max() > corecorrs.max()`
        else:
            take = True
        self._last_select = dict(best=best, pos=pos, take=bool(take))
        if take and pos >= 0:
            f = int(sub_idcs[pos]) if sub_idcs is not None else int(pos)
            if f not in self.idcs:
                row = np.asarray(self.data[f], dtype=np.float64).reshape(1, -1)
                self.wts = np.append(self.wts, 0.)
                self.idcs = np.append(self.idcs, np.int64(f)).astype(np.int64)
                self.pts = np.vstack((self.pts.reshape(-1, self._ncols), row))

    # bcores.py:141-150 / sparsevi.py:129-136
    def _optimize(self):
        t = self._get_tangent()
        M = self.wts.shape[0]
        if M == 0:
            # the reference still calls the sampler (and draws the subsample) opt_itrs times
            for i in range(self.opt_itrs):
                self.ll_projector.update(self.wts, self.pts)
                if self.n_subsample_opt is not None:
                    rng.randint(self._n_total, self.n_subsample_opt)
            return
        core = self._core_operand(t)
        g = t.eng.empty(M)
        beta = self._beta()
        scaling = 1. if self.n_subsample_opt is None else self._n_total/self.n_subsample_opt
        if isinstance(t, _FusedTangent) and FUSED_STEP_CALL:
            self.wts = self._optimize_fused_steps(t, core, beta, scaling)
            return

        def grd(w_host, w_dev):
            t.begin(w_host, self.pts, beta)
            if self.n_subsample_opt is not None:
                sub_idcs = rng.randint(self._n_total, self.n_subsample_opt)
            elif self.groups is not None and not self._groups_cover:
                sub_idcs = self._groups_flat        # bcores.py:46-51: the data term is the sum over the grouped rows
            else:
                sub_idcs = None
            colsum = t.colsum(sub_idcs)
            Vc = t.core_rows(core)
            resid = t.residual(colsum, scaling, Vc, w_dev)
            return t.grad(Vc, resid, g)
        grd.wants_device_iterate = True
        self.wts = nn_opt(self.wts, grd, opt_itrs=self.opt_itrs, step_sched=self.step_sched)

    def _optimize_fused_steps(self, t, core, beta, scaling):
        """the optimiser loop with ONE library call per step (bc_greedy_opt_step: the same kernels in the same order as the
        call-by-call loop, launched back to back from C; a row-sharded job calls its two halves around the exchange of the
        column-sum parts).

        Sampler protocol.  The reference's sampler is a host callback `sampler(S, wts, pts)` and stays one: the weights come
        back to the host every step for it.  A sampler that also offers `device_step(S, w_dev, core)` (this package's
        Laplace / conjugate samplers) forms the posterior and the samples from the device-resident iterate instead; the loop
        then has NO host synchronisation at all -- the host queues step after step (and draws the normals a call ahead,
        util/rng.py), the weights are read back once at the end."""
        from .. import _fused
        eng, fp, rows, comm = t.eng, t.fp, t.rows, t.comm
        M = self.wts.shape[0]
        b1, b2, eps = 0.9, 0.999, 1e-8                     # util/opt.py:36 defaults (nn_opt is called with them)
        x = eng.upload(np.asarray(self.wts, dtype=np.float64))
        m1, m2, g = eng.zeros(M), eng.zeros(M), eng.empty(M)
        xh = np.asarray(self.wts, dtype=np.float64).copy()
        prj = self.ll_projector
        smp = getattr(prj, 'sampler', None)
        on_device = DEVICE_SAMPLER_LOOP and hasattr(smp, 'device_step') and smp.supports_device_step() and not getattr(prj, 'encoder', None)
        sub_mode = self.n_subsample_opt is not None
        if sub_mode:
            n_pass, fixed_sub = int(self.n_subsample_opt), None
        elif self.groups is not None and not self._groups_cover:
            _, loc = local_subsample(self._groups_flat, rows.row0, rows.n_local)
            fixed_sub = eng.upload(np.asarray(loc, dtype=np.int64), dtype=torch.int64)
            n_pass = int(fixed_sub.numel())
        else:
            n_pass, fixed_sub = rows.n_local, None
        a = nv.StepArgs()
        bufs = None
        idx_dev = eng.empty(max(n_pass, 1), dtype=torch.int64) if sub_mode else fixed_sub
        idx_pin = [torch.empty(max(n_pass, 1), dtype=torch.int64).pin_memory() for _ in range(2)] if sub_mode else None
        idx_ev = [None, None]
        world = comm.world
        th = None

        def setup(th):
            """(first step, or the sample count changed) workspaces and every pointer of the step that does not change"""
            S = fp.S
            q = fp._q_operands(rows, None)                 # row image of the whole block (built once), exponents / digits applied
            if q is None and rows.n_local > 0 and _fused.ROUTE == 'q' and fp.D <= nv.lib().bc_q_max_features():
                raise nv.NativeError('row image unavailable')
            b = dict(S=S, Vc=eng.empty(M, S), parts=eng.empty(2*fp.Sld), colsum=eng.empty(S), resid=eng.empty(S+1), q=q)
            a.S, a.ldt = S, int(th.stride(0))
            if q is not None:
                a.d_image, a.d_rowscale, a.d_rowaux_q = q[0].data_ptr(), q[1].data_ptr(), (q[2].data_ptr() if q[2] is not None else None)
                if idx_dev is not None:
                    gs = fp.gather_scratch(n_pass)
                    a.d_gimage, a.d_growscale, a.d_growaux = gs[0].data_ptr(), gs[1].data_ptr(), gs[2].data_ptr()
            else:
                a.d_image = None
                ra = fp._rowaux(rows) if rows.n_local else None
                a.d_X, a.ldx, a.d_rowaux = (rows.t.data_ptr() if rows.n_local else None), rows.ld, (ra.data_ptr() if ra is not None else None)
            cra = fp._rowaux(core)
            a.d_pts, a.ldp, a.M, a.d_pts_rowaux = core.t.data_ptr(), core.ld, M, (cra.data_ptr() if cra is not None else None)
            a.d_Vc, a.ldv = b['Vc'].data_ptr(), S
            a.d_parts, a.d_colsum, a.d_resid, a.d_grad = b['parts'].data_ptr(), b['colsum'].data_ptr(), b['resid'].data_ptr(), g.data_ptr()
            a.d_w, a.d_m1, a.d_m2, a.b1, a.b2, a.eps, a.d_nn_mask = x.data_ptr(), m1.data_ptr(), m2.data_ptr(), b1, b2, eps, None
            a.scaling, a.n = float(scaling), n_pass
            a.d_rows = idx_dev.data_ptr() if idx_dev is not None else None
            return b

        # ---- replayed loop: the whole device side of a step as a CUDA graph (single rank, sampler with the graph protocol) ----
        use_graph = (STEP_GRAPH and on_device and world == 1 and _fused.PASS_TIMERS is None and hasattr(smp, 'graph_enqueue')
                     and self.opt_itrs >= STEP_GRAPH_MIN_ITRS and (not sub_mode or (rows.row0 == 0 and rows.n_local == self._n_total)))
        if use_graph:
            return self._optimize_graph_steps(t, core, beta, smp, prj, a, setup, x, idx_dev, idx_pin, sub_mode)

        for i in range(self.opt_itrs):
            sev = None
            if STEP_EVENTS is not None and world > 1:
                sev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                sev[0].record()
                hst = [time.perf_counter()]
            if on_device:
                th = smp.device_step(prj.projection_dimension, x, core)
                if sev:
                    sev.append(torch.cuda.Event(enable_timing=True))
                    sev[4].record()                            # the sampler's kernels end here, sample preparation follows
                    hst.append(time.perf_counter())
            else:
                prj.update(xh, self.pts)                       # host sampler: consumes np.random exactly like the reference
                th = prj.samples
                if isinstance(th, torch.Tensor):
                    th = th.to(device=eng.device, dtype=torch.float64)
                    if not th.is_contiguous():
                        th = th.contiguous()
                else:
                    th = eng.upload(np.ascontiguousarray(np.atleast_2d(np.asarray(th, dtype=np.float64))))
            if world > 1 and (i == 0 or not on_device):
                t._theta_buf = th
                t._replicate_samples()                         # checked, not assumed: see _FusedTangent._replicate_samples
            fp.configure(beta)
            fp.note_samples(th)
            if bufs is None or bufs['S'] != fp.S:
                bufs = setup(th)
            a.d_theta = th.data_ptr()
            if sub_mode:
                sub = rng.randint(self._n_total, self.n_subsample_opt)                    # bcores.py:53, after the sampler call
                if world > 1 or rows.row0 != 0 or rows.n_local != self._n_total:
                    _, sub = local_subsample(sub, rows.row0, rows.n_local)                # the sub-sampled rows this rank owns
                k = i & 1
                if idx_ev[k] is not None:
                    idx_ev[k].synchronize()                    # the upload that last used this pinned buffer has left it
                pin = idx_pin[k]
                pin.numpy()[:len(sub)] = sub
                idx_dev[:len(sub)].copy_(pin[:len(sub)], non_blocking=True)
                idx_ev[k] = torch.cuda.Event()
                idx_ev[k].record()
                a.n = len(sub)
            a.lr, a.c1, a.c2 = float(self.step_sched(i)), 1.-b1**(i+1), 1.-b2**(i+1)
            if _fused.PASS_TIMERS is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                e1.record()                                    # creates the CUDA events; the library re-records them around the pass
                a.ev_pass_begin, a.ev_pass_end = e0.cuda_event, e1.cuda_event
                _fused.PASS_TIMERS.append(('colsum', int(a.n), e0, e1))
            else:
                a.ev_pass_begin = a.ev_pass_end = None
            if world == 1:
                a.phase = 0
                nv.call('bc_greedy_opt_step', t.ctx, ctypes.byref(a), stream_ptr())
            else:
                a.phase = 1
                nv.call('bc_greedy_opt_step', t.ctx, ctypes.byref(a), stream_ptr())
                if sev:
                    sev[1].record()
                    hst.append(time.perf_counter())
                allparts = comm.allgather(bufs['parts'])       # (world, 2 Sld): one double-double part per rank, rank order
                if sev:
                    sev[2].record()
                    hst.append(time.perf_counter())
                a.phase, a.nparts, a.d_parts_all = 2, world, allparts.data_ptr()
                nv.call('bc_greedy_opt_step', t.ctx, ctypes.byref(a), stream_ptr())
                if sev:
                    sev[3].record()
                    hst.append(time.perf_counter())
                    STEP_EVENTS.append((sev, _fused.PASS_TIMERS[-1] if _fused.PASS_TIMERS else None, hst))
            if not on_device:
                xh = x.cpu().numpy()                           # one D2H + sync per step: the next sampler call needs the weights
        if on_device:
            xh = x.cpu().numpy()
            prj.samples = th
        return xh

    def _optimize_graph_steps(self, t, core, beta, smp, prj, a, setup, x, idx_dev, idx_pin, sub_mode):
        """The host-free optimiser loop with the device side of a step -- upload of the step's normals and sub-sample indices,
        the sampler's kernels, sample preparation, the data pass, the coreset rows, residual / gradient / ADAM -- captured once
        per staging-buffer parity into a CUDA graph and REPLAYED: one graph launch per step instead of a dozen kernel launches
        and two copies.  What changes from step to step is kept out of the launch parameters: the normals and indices arrive
        in fixed pinned buffers, the step size and ADAM's bias corrections (util/opt.py:45-52) are read from a device-resident
        schedule indexed by a device-side step counter.  The first two steps run eagerly (lazy allocations, one-time
        set-up), so the captured stream work is pure launches.  Same kernels, same order, same bits as the eager loop
        (tests/test_gpu_parity.py::test_graph_replayed_optimiser_loop_changes_no_bit)."""
        eng, fp = t.eng, t.fp
        S = prj.projection_dimension
        b1, b2 = 0.9, 0.999
        n = self.opt_itrs
        sched = np.array([[float(self.step_sched(i)), 1.-b1**(i+1), 1.-b2**(i+1)] for i in range(n)], dtype=np.float64)
        sched_dev = eng.upload(sched.ravel())
        counter = torch.zeros(1, dtype=torch.int32, device=eng.device)
        a.d_sched, a.d_step_counter = sched_dev.data_ptr(), counter.data_ptr()
        a.ev_pass_begin = a.ev_pass_end = None
        a.phase = 0
        idx_ev = [None, None]
        graphs, slots, nodes = [None, None], [None, None], [0, 0]
        state = {'bufs': None, 'th': [None, None]}
        lib = nv.lib()

        def enqueue(k):
            """everything a step puts on the stream; only fixed buffers are named"""
            th = smp.graph_enqueue(S, x, core, k)
            if sub_mode:
                idx_dev.copy_(idx_pin[k], non_blocking=True)
            fp.configure(beta)
            fp.note_samples(th)
            if state['bufs'] is None:
                state['bufs'] = setup(th)
            a.d_theta = th.data_ptr()
            nv.call('bc_greedy_opt_step', t.ctx, ctypes.byref(a), stream_ptr())
            state['th'][k] = th

        k_last = None
        for i in range(n):
            k = smp.graph_normals(S)                               # host: the step's normals are in pinned staging buffer k
            if sub_mode:
                sub = rng.randint(self._n_total, self.n_subsample_opt)     # bcores.py:53, after the sampler's draw
                if idx_ev[k] is not None:
                    idx_ev[k].synchronize()                        # the upload that last read this pinned buffer has run
                idx_pin[k].numpy()[:] = sub
            if i < 2:
                enqueue(k)
            else:
                if graphs[k] is None:
                    l0 = lib.bc_launch_count()
                    gr = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gr, capture_error_mode='thread_local'):
                        enqueue(k)
                    graphs[k], slots[k], nodes[k] = gr, lib.bc_sample_slot(t.ctx), int(lib.bc_launch_count() - l0)
                    lib.bc_add_launch_count(-nodes[k])             # the capture itself ran nothing
                graphs[k].replay()
                lib.bc_add_launch_count(nodes[k])
                k_last = k
            smp.graph_launched(k)
            if sub_mode:
                idx_ev[k] = torch.cuda.Event()
                idx_ev[k].record()
        if k_last is not None:
            nv.call('bc_set_sample_slot', t.ctx, int(slots[k_last]))   # host view of the scratch-slot parity = the device's
        xh = x.cpu().numpy()
        prj.samples = state['th'][k].clone()      # (a copy: the replayed buffers belong to the graphs' memory pool)
        a.d_sched = a.d_step_counter = None
        return xh

    def error(self):
        return 0.       # the reference has no KL estimate either (bcores.py:152-153)
