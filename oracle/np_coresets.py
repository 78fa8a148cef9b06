"""Oracle: greedy coreset construction (beta-Cores / SparseVI / Hilbert).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates
bayesiancoresets/coreset/{coreset,bcores,sparsevi,hilbert,projector}.py and
bayesiancoresets/util/opt.py:36-54 for the individual-point mode with
learn_beta=False (the only mode the reference drivers run).
RNG call order (sampler, then np.random.randint for the subsample) is the
reference's, so seeded runs reproduce its index sequence.
"""
import numpy as np
from . import np_snnls


def centred(F):
    """projector.py:26 / :55: subtract each row's mean over the S samples (in place)."""
    F -= F.mean(axis=1)[:, np.newaxis]
    return F


def adam_nonneg(x0, grad, itrs, sched, b1=0.9, b2=0.999, eps=1e-8):
    """util/opt.py:36-54 nn_opt: ADAM with projection onto x >= 0."""
    x = x0.copy()
    m1 = np.zeros(x.shape[0])
    m2 = np.zeros(x.shape[0])
    for i in range(itrs):
        g = grad(x)
        m1 = b1*m1 + (1.-b1)*g
        m2 = b2*m2 + (1.-b2)*g**2
        upd = sched(i)*m1/(1.-b1**(i+1))/(eps + np.sqrt(m2/(1.-b2**(i+1))))
        x -= upd
        x = np.maximum(x, 0.)
    return x


class GreedyVI(object):
    """BetaCoreset (potential = beta-likelihood, coreset/bcores.py) and
    SparseVICoreset (potential = log-likelihood, coreset/sparsevi.py) share this
    skeleton; `potential(pts, samples)` returns the un-centred (n, S) matrix.

    sampler(S, wts, pts) -> (S, D) is the user's host callback (projector.py:36-37,65-66).
    """

    def __init__(self, data, sampler, S, potential, n_sub_select=None, n_sub_opt=None,
                 opt_itrs=100, sched=lambda i: 1./(1.+i), groups=None, initialized=False):
        self.data = data
        self.sampler = sampler
        self.S = S
        self.potential = potential
        N = data.shape[0]
        self.n_sub_select = None if n_sub_select is None else min(N, n_sub_select)   # bcores.py:14
        self.n_sub_opt = None if n_sub_opt is None else min(N, n_sub_opt)            # bcores.py:15
        self.opt_itrs = opt_itrs
        self.sched = sched
        self.wts = np.array([])
        self.idcs = np.array([], dtype=np.int64)
        self.pts = np.array([])
        self.groups = groups                                   # bcores.py:22 / sparsevi.py:20: list of row-index lists
        self.selected_groups = []
        self.initialized = int(initialized)*len(self.wts)      # bcores.py:25
        self.samples = sampler(S, np.array([]), np.array([]))   # projector.py:18,46 (ctor draws once)
        self.log = []   # oracle extra: per-select diagnostics

    def size(self):
        return (self.wts > 0).sum()                      # coreset.py:22-23

    def get(self):
        keep = self.wts > 0                              # coreset.py:25-26
        return self.wts[keep], self.pts[keep, :], self.idcs[keep]

    def _group_rows(self, gi):
        """bcores.py:50,60: one row per listed group = the sum of the group's centred rows"""
        return np.array([np.sum(centred(self.potential(self.data[self.groups[i], :], self.samples)), axis=0) for i in gi])

    def _tangent(self, n_sub, w, p, select=False):
        """bcores.py:37-72 / sparsevi.py:35-70."""
        self.samples = self.sampler(self.S, w, p)        # update() first: consumes np.random
        self._gi = None
        if n_sub is None and self.groups is None:
            sub = None
            vecs = centred(self.potential(self.data, self.samples))
            scale = 1.
        elif n_sub is None:                                              # bcores.py:46-51
            self._gi = list(range(len(self.groups)))
            sub = [r for i in self._gi for r in self.groups[i]]
            vecs = self._group_rows(self._gi)
            scale = 1.
        elif self.groups is not None and select:                         # bcores.py:56-61
            self._gi = np.random.randint(len(self.groups), size=n_sub)
            sub = [r for i in self._gi for r in self.groups[i]]
            vecs = self._group_rows(self._gi)
            scale = len(self.groups)/n_sub
        else:
            sub = np.random.randint(self.data.shape[0], size=n_sub)     # bcores.py:53
            vecs = centred(self.potential(self.data[sub], self.samples))
            scale = self.data.shape[0]/n_sub
        if self.pts.size > 0:
            core = centred(self.potential(p, self.samples))
        else:
            core = np.zeros((0, vecs.shape[1]))
        # NB the all-zero-row filter at bcores.py:67-68 / sparsevi.py:63-64 only runs when
        # select=True, and the individual-point _select (bcores.py:76, sparsevi.py:74) never
        # passes it: rows whose centred vector is exactly 0 stay in and score 0/0 = NaN.
        return vecs, scale, sub, core

    def select_group(self):
        """bcores.py:91-123 / sparsevi.py:93-126: add a whole group of rows."""
        gv, scale, sub, core = self._tangent(self.n_sub_select, self.wts, self.pts, select=True)
        if self.n_sub_select is None:
            resid = gv.sum(axis=0) - self.wts.dot(core)
        else:
            resid = scale*gv.sum(axis=0) - self.wts.dot(core)
        with np.errstate(invalid='ignore', divide='ignore'):
            corrs = gv.dot(resid) / np.sqrt((gv**2).sum(axis=1)) / gv.shape[1]
            ccorrs = np.fabs(core.dot(resid) / np.sqrt((core**2).sum(axis=1))) / core.shape[1]
        maxcc = ccorrs[self.initialized:].max() if ccorrs.shape[0] > self.initialized else -np.inf
        if ccorrs.size == 0 or corrs.max() > maxcc:
            f = np.argmax(corrs) if self.n_sub_select is None else self._gi[np.argmax(corrs)]
            if f not in self.selected_groups:
                self.selected_groups.append(f)
                new = self.data[self.groups[f], :]
                self.wts = np.append(self.wts, np.zeros(new.shape[0]))
                self.idcs = np.append(self.idcs, np.asarray(self.groups[f], dtype=np.int64))
                self.pts = np.vstack((self.pts.reshape(-1, self.data.shape[1]), new))

    def select(self):
        """bcores.py:75-90 / sparsevi.py:73-92."""
        if self.groups is not None:
            return self.select_group()
        vecs, scale, sub, core = self._tangent(self.n_sub_select, self.wts, self.pts)
        resid = scale*vecs.sum(axis=0) - self.wts.dot(core)
        with np.errstate(invalid='ignore', divide='ignore'):
            corrs = vecs.dot(resid) / np.sqrt((vecs**2).sum(axis=1)) / vecs.shape[1]
            ccorrs = np.fabs(core.dot(resid) / np.sqrt((core**2).sum(axis=1))) / core.shape[1]
        # np.max / np.argmax propagate NaN (first NaN wins); `nan > x` is False (bcores.py:80)
        rec = {'best': float(corrs.max()), 'pos': int(np.argmax(corrs)),
               'core_best': float(ccorrs.max()) if ccorrs.size else None, 'added': False}
        if ccorrs.size == 0 or corrs.max() > ccorrs.max():
            f = sub[np.argmax(corrs)] if sub is not None else np.argmax(corrs)
            rec['f'] = int(f)
            if f not in self.idcs:
                self.wts = np.append(self.wts, 0.)
                self.idcs = np.append(self.idcs, f)
                self.pts = np.vstack((self.pts.reshape(-1, self.data.shape[1]), self.data[f][np.newaxis, :]))
                rec['added'] = True
        self.log.append(rec)

    def optimise(self):
        """bcores.py:141-150 / sparsevi.py:129-136."""
        def grad(w):
            vecs, scale, _, core = self._tangent(self.n_sub_opt, w, self.pts)
            resid = scale*vecs.sum(axis=0) - w.dot(core)
            return -core.dot(resid) / core.shape[1]
        self.wts = adam_nonneg(self.wts, grad, self.opt_itrs, self.sched)

    def build(self, itrs, sz):
        """coreset.py:33-45 + bcores.py:27-35."""
        if sz < self.size():
            raise ValueError('cannot shrink')
        if self.groups is None and self.size()+itrs > sz:      # bcores.py:28-30: no size guard in group mode
            raise ValueError('itrs + size > sz')
        for _ in range(itrs):
            self.select()
            self.optimise()


class GreedyVILearnBeta(GreedyVI):
    """BetaCoreset with learn_beta=True.  PARITY UNPINNED: the reference's own branch (coreset/bcores.py:127-140) calls
    `self._get_projection_ii`, which no file of the reference defines, so it raises AttributeError on the first
    optimisation and there is no reference output to pin against.  This class restates what that branch spells out
    around the missing call:
        grd(x):  w, beta = x[:-1], x[-1]                                               bcores.py:129-130
                 vecs, sum_scaling, sub_idcs, corevecs, betagrads = projection at (w, pts, beta)       :131
                     with betagrads = the centred beta-gradient of the coreset points, what
                     BetaBlackBoxProjector.project_f(pts, beta, grad=True) returns second     projector.py:56-61
                 resid = sum_scaling vecs.sum(0) - w.corevecs                                           :132
                 wgrad = -corevecs.resid / S;  betagrad = -1e-5 w.(betagrads.resid) / S              :133-134
        x0 = [wts, beta];  xf = partial_nn_opt(x0, grd, all coordinates clamped at 0);  wts, beta = xf[:-1], xf[-1]   :137-140
    potential_beta(pts, samples, beta) and beta_gradient(pts, samples, beta) return un-centred (n, S) matrices."""

    def __init__(self, data, sampler, S, potential_beta, beta_gradient, beta, **kw):
        self.beta = beta
        self.potential_beta = potential_beta
        self.beta_gradient = beta_gradient
        GreedyVI.__init__(self, data, sampler, S, lambda p, th: self.potential_beta(p, th, self.beta), **kw)

    def optimise(self):
        saved = self.beta

        def grad(x):
            w, self.beta = x[:-1], x[-1]           # the projection below evaluates the potential at the iterate's beta
            vecs, scale, _, core = self._tangent(self.n_sub_opt, w, self.pts)
            bg = centred(self.beta_gradient(self.pts, self.samples, self.beta)) if self.pts.size > 0 else np.zeros((0, vecs.shape[1]))
            resid = scale*vecs.sum(axis=0) - w.dot(core)
            wgrad = -core.dot(resid)/core.shape[1]
            bgrad = -1e-5*w.dot(bg.dot(resid))/core.shape[1]
            return np.hstack((wgrad, bgrad))
        x0 = np.hstack((self.wts, np.asarray([saved])))
        xf = adam_partial_nonneg(x0, grad, np.arange(x0.shape[0]), self.opt_itrs, self.sched)
        self.wts, self.beta = xf[:-1], xf[-1]          # a numpy float64, as in the reference (:140)


def adam_partial_nonneg(x0, grad, nn_idcs, itrs, sched, b1=0.9, b2=0.999, eps=1e-8):
    """util/opt.py:56-77 partial_nn_opt: ADAM, projection onto x >= 0 on the coordinates nn_idcs only."""
    x = x0.copy()
    m1 = np.zeros(x.shape[0])
    m2 = np.zeros(x.shape[0])
    for i in range(itrs):
        g = grad(x)
        m1 = b1*m1 + (1.-b1)*g
        m2 = b2*m2 + (1.-b2)*g**2
        upd = sched(i)*m1/(1.-b1**(i+1))/(eps + np.sqrt(m2/(1.-b2**(i+1))))
        x -= upd
        x[nn_idcs] = np.maximum(x[nn_idcs], 0.)
    return x


class BatchPSVI(object):
    """coreset/bpsvi.py:6-65: pseudo-coreset -- sz points drawn from the data, then weights AND point locations optimised
    jointly; potential = log-likelihood, grad_potential(pts, samples) -> (M, S, D) its gradient in the point."""

    def __init__(self, data, sampler, S, potential, grad_potential, opt_itrs, n_sub_opt=None,
                 sched=lambda m: (lambda i: 1./(1.+i))):
        self.data = data
        self.sampler = sampler
        self.S = S
        self.potential = potential
        self.grad_potential = grad_potential
        self.opt_itrs = opt_itrs
        self.n_sub_opt = None if n_sub_opt is None else min(data.shape[0], n_sub_opt)     # bpsvi.py:11
        self.sched = sched
        self.wts = np.array([])
        self.idcs = np.array([], dtype=np.int64)
        self.pts = np.array([])
        self.samples = sampler(S, np.array([]), np.array([]))   # projector ctor draw (projector.py:18)

    def size(self):
        return (self.wts > 0).sum()

    def get(self):
        keep = self.wts > 0
        return self.wts[keep], self.pts[keep, :], self.idcs[keep]

    def _tangent(self, w, p):
        """bpsvi.py:26-42"""
        self.samples = self.sampler(self.S, w, p)
        if self.n_sub_opt is None:
            vecs = centred(self.potential(self.data, self.samples))
            scale = 1.
        else:
            sub = np.random.randint(self.data.shape[0], size=self.n_sub_opt)
            vecs = centred(self.potential(self.data[sub], self.samples))
            scale = self.data.shape[0]/self.n_sub_opt
        core = centred(self.potential(p, self.samples))
        pg = self.grad_potential(p, self.samples)
        pg -= pg.mean(axis=2)[:, :, np.newaxis]                 # projector.py:31: centred over the LAST axis (D), as shipped
        return vecs, scale, core, pg

    def build(self, itrs, sz):
        """coreset.py:33-45 + bpsvi.py:17-24 (itrs is ignored by the reference)"""
        if sz < self.size():
            raise ValueError('cannot shrink')
        init = np.random.choice(self.data.shape[0], size=sz, replace=False)
        self.pts = self.data[init]
        self.wts = self.data.shape[0]/sz*np.ones(sz)
        self.idcs = init
        self.optimise()

    def optimise(self):
        """bpsvi.py:44-62"""
        sz = self.wts.shape[0]
        d = self.pts.shape[1]

        def grad(x):
            w = x[:sz]
            p = x[sz:].reshape((sz, d))
            vecs, scale, core, pg = self._tangent(w, p)
            resid = scale*vecs.sum(axis=0) - w.dot(core)
            wgrad = -core.dot(resid) / core.shape[1]
            ugrad = -(w[:, np.newaxis, np.newaxis]*pg*resid[np.newaxis, :, np.newaxis]).sum(axis=1)/core.shape[1]
            return np.hstack((wgrad, ugrad.reshape(sz*d)))
        x0 = np.hstack((self.wts, self.pts.reshape(sz*d)))
        xf = adam_partial_nonneg(x0, grad, np.arange(sz), self.opt_itrs, self.sched(sz))
        self.wts = xf[:sz]
        self.pts = xf[sz:].reshape((sz, d))


class Hilbert(object):
    """coreset/hilbert.py:7-43: project once, drop zero-norm rows, delegate to a solver."""

    def __init__(self, data, sampler, S, potential, n_sub=None, solver='giga'):
        self.data = data
        samples = sampler(S, np.array([]), np.array([]))       # projector ctor draw
        if n_sub is None:
            self.sub = None
            vecs = centred(potential(data, samples))
        else:
            n_sub = min(data.shape[0], n_sub)
            self.sub = np.random.randint(data.shape[0], size=n_sub)
            vecs = centred(potential(data[self.sub], samples))
        vecs = vecs[np.sqrt((vecs**2).sum(axis=1)) > 0., :]     # hilbert.py:16
        self.vecs = vecs
        self.solver = np_snnls.SOLVERS[solver](vecs.T, vecs.sum(axis=0))
        self.wts = np.array([])
        self.idcs = np.array([], dtype=np.int64)
        self.pts = np.array([])

    def _sync(self):
        w = self.solver.weights()                               # hilbert.py:30-33
        self.wts = w[w > 0]
        self.idcs = self.sub[w > 0] if self.sub is not None else np.where(w > 0)[0]
        self.pts = self.data[self.idcs]

    def build(self, itrs, sz):
        if self.solver.size()+itrs > sz:
            raise ValueError('itrs + size > sz')
        self.solver.run(itrs)
        self._sync()

    def optimise(self):
        self.solver.polish()
        self._sync()

    def error(self):
        return self.solver.error()
