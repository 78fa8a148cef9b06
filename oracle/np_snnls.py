"""Oracle: sparse non-negative least squares solvers on a fixed (S, N) matrix A.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates
bayesiancoresets/snnls/{snnls,giga,frankwolfe,orthopursuit}.py.  The solver
state is a plain dict-like object; `step_*` functions are the reference's
`_select` + `_reweight`, `run` is `SparseNNLS.build`.
"""
import numpy as np
from scipy.optimize import nnls as _scipy_nnls

TOL = 1e-12   # bayesiancoresets/util/__init__.py:4


class PrecisionLoss(Exception):
    """Stands for bayesiancoresets/util/errors.py:1 NumericalPrecisionError."""


class Solver(object):
    """State + driver loop shared by the three solvers (snnls/snnls.py:8-106)."""
    kind = None

    def __init__(self, A, b):
        self.A = A
        self.b = b
        self.w = np.zeros(A.shape[1])              # snnls.py:15
        self.hit_limit = False
        self.monotone = True                       # snnls.py:9 check_error_monotone default
        self.trace = []                            # selected f per successful select (oracle extra)
        norms = np.sqrt((A**2).sum(axis=0))        # giga.py:10 / frankwolfe.py:10 / orthopursuit.py:12
        if np.any(norms == 0):
            raise ValueError('A must not have any 0 columns')
        self.norms = norms
        self.An = A / norms
        self._setup()

    def _setup(self):
        pass

    # snnls.py:22-29
    def size(self):
        return (self.w > 0).sum()

    def weights(self):
        return self.w.copy()

    def error(self):
        return np.sqrt(((self.A.dot(self.w) - self.b)**2).sum())

    def reset(self):
        self.w = np.zeros(self.A.shape[1])
        self.hit_limit = False

    def run(self, itrs):
        """snnls.py:31-78: monotone check, one retry, numeric-limit latch."""
        if self.hit_limit or self.A.size == 0:
            return
        retried = False
        for _ in range(itrs):
            try:
                nonempty = self.size() > 0
                if self.monotone and nonempty:
                    e0 = self.error()
                    w0 = self.w.copy()
                f = self.select()
                self.trace.append(int(f))
                self.reweight(f)
                if self.monotone and nonempty:
                    e1 = self.error()
                    if e1 > e0:
                        self.w = w0
                        raise PrecisionLoss('error not monotone')
                    retried = False
            except PrecisionLoss:
                if retried:
                    self.hit_limit = True
                    break
                retried = True
            if self.hit_limit:
                break

    def polish(self):
        """snnls.py:82-97: scipy Lawson-Hanson on the active columns; revert if worse."""
        e0 = self.error()
        w0 = self.w.copy()
        act = self.w > 0
        sol = _scipy_nnls(self.A[:, act], self.b, maxiter=100*self.A.shape[1])
        self.w[act] = sol[0]
        if self.error() > e0*(1.+TOL):
            self.w = w0
            self.hit_limit = True


class Giga(Solver):
    kind = 'giga'

    def _setup(self):
        # giga.py:15-18
        self.bnorm = np.sqrt(((self.b)**2).sum())
        if self.bnorm == 0.:
            raise PrecisionLoss('norm of b must be > 0')
        self.bn = self.b / self.bnorm

    def select(self):
        # giga.py:20-38
        xw = self.A.dot(self.w)
        nw = np.sqrt(((xw)**2).sum())
        nw = 1. if nw == 0. else nw
        xw /= nw
        cdir = self.bn - self.bn.dot(xw)*xw
        cn = np.sqrt((cdir**2).sum())
        if cn < TOL:
            raise PrecisionLoss('cdirnrm < TOL')
        cdir /= cn
        sc = self.An.T.dot(np.hstack((cdir[:, np.newaxis], xw[:, np.newaxis])))
        ok = np.logical_and(sc[:, 1] > -1.+1e-14, 1.-sc[:, 1]**2 > 0.)
        sc[ok, 1] = np.sqrt(1.-sc[ok, 1]**2)
        sc[np.logical_not(ok), 1] = np.inf
        return (sc[:, 0]/sc[:, 1]).argmax()

    def reweight(self, f):
        # giga.py:40-64
        xw = self.A.dot(self.w)
        nw = np.sqrt((xw**2).sum())
        nw = 1. if nw == 0. else nw
        xf = self.A[:, f]
        nf = np.sqrt((xf**2).sum())
        gA = self.bn.dot((xf/nf)) - self.bn.dot((xw/nw)) * (xw/nw).dot((xf/nf))
        gB = self.bn.dot((xw/nw)) - self.bn.dot((xf/nf)) * (xw/nw).dot((xf/nf))
        if gA <= 0. or gB < 0:
            raise PrecisionLoss
        a = gB/(gA+gB)/nw
        b = gA/(gA+gB)/nf
        x = a*xw + b*xf
        nx = np.sqrt((x**2).sum())
        scale = self.bnorm/nx*(x/nx).dot(self.bn)
        alpha = a*scale
        beta = b*scale
        self.w = alpha*self.w
        self.w[f] = max(0., self.w[f]+beta)


class FrankWolfe(Solver):
    kind = 'fw'

    def select(self):
        # frankwolfe.py:15-17
        r = self.b - self.A.dot(self.w)
        return (self.An.T.dot(r)).argmax()

    def reweight(self, f):
        # frankwolfe.py:19-40
        if self.size() == 0:
            alpha = 0.
            beta = self.norms.sum() / self.norms[f]
        else:
            nsum = self.norms.sum()
            nf = self.norms[f]
            xw = self.A.dot(self.w)
            xf = self.A[:, f]
            num = (nsum/nf*xf - xw).dot(self.b-xw)
            den = ((nsum/nf*xf-xw)**2).sum()
            if num < 0. or den == 0. or num > den:
                raise PrecisionLoss('precision loss in gammanum/gammadenom')
            alpha = 1. - num/den
            beta = nsum/nf*num/den
        self.w = alpha*self.w
        self.w[f] = max(0., self.w[f]+beta)


class OrthoPursuit(Solver):
    kind = 'omp'

    def select(self):
        # orthopursuit.py:17-35
        r = self.b - self.A.dot(self.w)
        dots = self.An.T.dot(r)
        if self.size() == 0:
            return dots.argmax()
        fpos = dots.argmax()
        pos = dots[fpos]
        act = self.w > 0
        fneg = (-dots[act]).argmax()
        neg = (-dots[act])[fneg]
        if pos >= neg:
            return fpos
        return np.arange(self.w.shape[0])[act][fneg]

    def reweight(self, f):
        # orthopursuit.py:37-42
        self.w[f] = 1.
        act = self.w > 0
        sol = _scipy_nnls(self.A[:, act], self.b, maxiter=100*self.A.shape[1])
        self.w[act] = sol[0]


SOLVERS = {'giga': Giga, 'fw': FrankWolfe, 'omp': OrthoPursuit}
