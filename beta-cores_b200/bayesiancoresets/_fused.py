"""FusedProjection: one potential + the current posterior samples, bound to a bc_ctx.

Wraps the stage-1/2 entry points of libbetacores.so for the coreset classes:
    colsum_parts()  -> bc_project_colsum        (column sum of the centred n x S projection, never formed)
    score()         -> bc_project_score         (residual correlation + arg-max, never formed)
    materialise()   -> bc_project_materialise   (centred rows written out: coreset points, Hilbert)
"""
import ctypes
import os
import numpy as np
import torch

from . import _native as nv
from ._device import DeviceRows, ptr, stream_ptr


# Optional instrumentation (bench.py): when a list is installed here, every data-row pass appends
# (kind, rows, start_event, end_event) recorded on the launching stream around the C-ABI call.
PASS_TIMERS = None

# Contraction route of the full-block passes: 'q' = tcgen05 int8 split (csrc/bc_project_q.cu, feature count <= 128),
# 'dmma' = FP64 mma.sync (csrc/bc_project.cu, any shape; always used for gathered sub-samples and materialisation).
ROUTE = os.environ.get('BC_CONTRACTION', 'q')

# Precision tier of the tensor-core route: leading int8 digits of the 7-digit operand images a pass contracts
# (bc_set_contraction_digits: 7 = 28 digit pairs, the accuracy of an fp64 dgemm; 6 = 21 pairs, operands as if rounded
# to 46 bits; 5 = 15 pairs, 38 bits).  tests/test_gpu_parity.py::test_precision_tiers pins what each tier keeps exact.
DIGITS = int(os.environ.get('BC_Q_DIGITS', '6'))


def set_contraction_digits(n):
    global DIGITS
    if n not in (5, 6, 7):
        raise ValueError('contraction digits must be 5, 6 or 7')
    DIGITS = int(n)


def _timed(kind, n, fn):
    if PASS_TIMERS is None:
        return fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    PASS_TIMERS.append((kind, n, e0, e1))


class FusedProjection(object):
    def __init__(self, engine, potential, ncols, ctx_name='main'):
        if not potential.is_bound():
            raise TypeError('potential %s has unbound model constants %s' % (potential.__name__, potential.needs))
        self.eng = engine
        self.pot = potential
        self.ncols = ncols
        self.D = potential.feature_dim(ncols)
        self.ctx_name = ctx_name
        self.ctx = engine.ctx(ctx_name)
        self.siginv = engine.upload(potential.bound['Siginv']) if potential.model == 'gaussian' else None
        self.S = None
        self.Sld = None
        self._theta = None
        self._gather = None        # scratch (image, rowscale, rowaux, capacity in rows) of gathered passes

    # ---- configuration ----
    # The bc_ctx workspace behind `ctx_name` is shared by every FusedProjection of the engine (one potential and one
    # prepared sample set live in it at a time), so what is currently applied is remembered ON THE ENGINE, per workspace:
    # two coreset objects built alternately (a SparseVICoreset and a BetaCoreset, say) each find the other's potential
    # in place and re-apply their own.
    def _applied(self):
        return self.eng.ctx_state.setdefault(self.ctx_name, {'potential': None, 'samples_of': None})

    def configure(self, beta=None):
        st = self._applied()
        key = (id(self), beta)      # per projection object: the workspace also points at THIS object's Siginv copy
        if st['potential'] == key:
            if st['samples_of'] is not self:
                self.S = None      # another projection's samples are in place
            return
        p = nv.params8(self.pot.params(self.D, beta))
        nv.call('bc_set_potential', self.ctx, self.pot.model_id, self.pot.kind_id, self.D, p, ptr(self.siginv))
        st['potential'] = key
        st['potential_owner'] = self       # keeps id(self) from being recycled while the key is in place
        st['samples_of'] = None
        self.S = None          # bc_set_potential invalidates the prepared samples

    def set_samples(self, theta):
        """theta: host (S, D) ndarray or a device tensor; runs the sample-preparation kernels."""
        if isinstance(theta, torch.Tensor):
            t = theta
        else:
            theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
            t = self.eng.upload(theta)
        if t.shape[1] != self.D:
            raise ValueError('samples have %d columns, potential expects %d' % (t.shape[1], self.D))
        st = self._applied()
        if st['potential'] is None or st['potential'][0] != id(self):
            raise nv.NativeError('set_samples() must follow configure()')
        self._theta = t
        self.S = int(t.shape[0])
        self.Sld = nv.lib().bc_colsum_ld(self.S)
        nv.call('bc_set_samples', self.ctx, ptr(t), self.S, int(t.stride(0)), stream_ptr())
        st['samples_of'] = self

    def note_samples(self, theta):
        """bookkeeping of set_samples() for a caller that hands `theta` to the library itself (bc_greedy_opt_step runs
        bc_set_samples as its first stage)"""
        st = self._applied()
        if st['potential'] is None or st['potential'][0] != id(self):
            raise nv.NativeError('set_samples() must follow configure()')
        self._theta = theta
        self.S = int(theta.shape[0])
        self.Sld = nv.lib().bc_colsum_ld(self.S)
        st['samples_of'] = self

    def gather_scratch(self, n):
        """scratch image / scales / aux for a gathered pass of n rows on the tensor-core route"""
        if self._gather is None or self._gather[3] < n:
            cap = max(2*n, 1024)
            nb = nv.c_i64()
            nv.call('bc_q_image_bytes', cap, ctypes.byref(nb))
            self._gather = (torch.empty(nb.value, dtype=torch.uint8, device=self.eng.device), self.eng.empty(cap), self.eng.empty(cap), cap)
        return self._gather

    # ---- helpers ----
    def _rowaux(self, rows):
        if self.pot.model != 'gaussian':
            return None
        key = id(self.siginv)
        if rows.rowaux is None or rows.rowaux_key != key:
            out = self.eng.empty(max(rows.n_local, 1))
            nv.call('bc_rowquad', self.ctx, ptr(rows.t), rows.n_local, rows.ld, ptr(out), stream_ptr())
            rows.rowaux, rows.rowaux_key = out, key
        return rows.rowaux

    def _q_operands(self, rows, sub):
        """(image, rowscale, rowaux) when the tensor-core route applies to this pass, else None.  `sub` (a gather list of local
        row numbers): the listed rows of the dataset's image are re-packed into a compact scratch image first (bc_q_gather_rows)."""
        if ROUTE != 'q' or self.D > nv.lib().bc_q_max_features():
            return None
        aux_col = self.D if self.pot.model == 'neurlin' else None
        # in a sharded job the first call exchanges the feature exponents: every rank comes through here, rows or not
        img, rs, aux, fexp = rows.quantised(self.ctx, self.D, aux_col)
        if rows.n_local == 0 or (sub is not None and int(sub.numel()) == 0):
            return None
        if self.eng.fexp_applied.get(self.ctx_name) is not fexp:     # the ctx's sample image must match this row image
            nv.call('bc_set_feature_exponents', self.ctx, ptr(fexp), self.D, stream_ptr())
            self.eng.fexp_applied[self.ctx_name] = fexp
        if self._applied().get('digits') != DIGITS:
            nv.call('bc_set_contraction_digits', self.ctx, DIGITS)
            self._applied()['digits'] = DIGITS
        ra = self._rowaux(rows) if self.pot.model == 'gaussian' else aux
        if sub is None:
            return img, rs, ra
        n = int(sub.numel())
        gimg, grs, gaux, _ = self.gather_scratch(n)
        nv.call('bc_q_gather_rows', self.ctx, ptr(img), ptr(rs), ptr(ra), ptr(sub), n, ptr(gimg), ptr(grs), ptr(gaux) if ra is not None else None,
                stream_ptr())
        return gimg, grs, (gaux if ra is not None else None)

    def _check(self, rows):
        if self.S is None or self._applied()['samples_of'] is not self:
            raise nv.NativeError('set_samples() must follow configure() (another projection has used this workspace since)')
        if rows.ncols != self.ncols:
            raise ValueError('data rows have %d columns, projector was built for %d' % (rows.ncols, self.ncols))

    # ---- passes ----
    def colsum_parts(self, rows, sub=None, out=None):
        """(2*Sld,) tensor: hi plane, lo plane; element S = sum of row means.  `sub`: optional
        int64 device tensor of local row numbers (gather; duplicates allowed)."""
        self._check(rows)
        if out is None:
            out = self.eng.empty(2 * self.Sld)
        n = rows.n_local if sub is None else int(sub.numel())
        q = self._q_operands(rows, sub)
        if q is not None:
            _timed('colsum', n, lambda: nv.call('bc_project_colsum_q', self.ctx, ptr(q[0]), ptr(q[1]), n, ptr(q[2]), ptr(out),
                                                stream_ptr()))
            return out
        ra = self._rowaux(rows)
        _timed('colsum', n, lambda: nv.call('bc_project_colsum', self.ctx, ptr(rows.t), rows.ld, ptr(sub), n, ptr(ra), ptr(out),
                                            stream_ptr()))
        return out

    def combine(self, parts, nparts, out=None):
        if out is None:
            out = self.eng.empty(self.S)
        nv.call('bc_colsum_combine', self.ctx, ptr(parts), nparts, self.S, ptr(out), stream_ptr())
        return out

    def score(self, rows, sub, resid, idx_offset, out_best, scores=None):
        """out_best (>=2 doubles): best score, int64 bits of (position + idx_offset)."""
        self._check(rows)
        n = rows.n_local if sub is None else int(sub.numel())
        q = self._q_operands(rows, sub)
        if q is not None:
            _timed('score', n, lambda: nv.call('bc_project_score_q', self.ctx, ptr(q[0]), ptr(q[1]), n, ptr(q[2]), ptr(resid),
                                               int(idx_offset), ptr(out_best), ptr(scores), stream_ptr()))
            return
        ra = self._rowaux(rows)
        _timed('score', n, lambda: nv.call('bc_project_score', self.ctx, ptr(rows.t), rows.ld, ptr(sub), n, ptr(ra), ptr(resid),
                                           int(idx_offset), ptr(out_best), ptr(scores), stream_ptr()))

    def materialise(self, rows, sub=None, want_norms=False, want_colsum=False, raw=False):
        self._check(rows)
        n = rows.n_local if sub is None else int(sub.numel())
        V = self.eng.empty(max(n, 1), self.S)
        norms = self.eng.empty(max(n, 1)) if want_norms else None
        dd = self.eng.empty(2 * self.Sld) if want_colsum else None
        nv.call('bc_project_materialise', self.ctx, ptr(rows.t), rows.ld, ptr(sub), n, ptr(self._rowaux(rows)), ptr(V), self.S,
                ptr(norms), ptr(dd), 1 if raw else 0, stream_ptr())
        return V[:n], (norms[:n] if want_norms else None), dd
