"""Seeded synthetic problems shared by the golden generator, the oracle tests and
the GPU parity tests.  Pure numpy; imports neither the reference nor the product.

Samplers are *user-side* host callbacks (sampler(S, wts, pts) -> (S, D)) exactly
as in the reference drivers (examples/zellner_gaussian/main.py:87-92,
examples/zellner_logreg/main.py:139-144): conjugate posterior / Laplace
approximation of the weighted coreset posterior, then mu + randn(S, D) L^T on
the global legacy numpy RNG.
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import np_models as om   # host sampler maths (restated reference functions)


# ------------------------------------------------------------------ G1 inputs --
def model_function_inputs():
    r = np.random.RandomState(11)
    N, D, S = 37, 5, 13
    d = 6
    A = r.randn(d, d)
    Sig = A.dot(A.T) + d*np.eye(d)
    p = dict(
        Z=r.randn(N, D)*2.0, Th=r.randn(S, D), beta=np.float64(0.3),
        Xg=r.randn(N, d)*3.0, Thg=r.randn(S, d), Siginv=np.linalg.inv(Sig), logdetSig=np.float64(np.linalg.slogdet(Sig)[1]),
        Zn=np.hstack((r.randn(N, D), r.randn(N, 1)*2.0)), Thn=r.randn(S, D), sigsq=np.float64(0.7))
    # large-margin rows exercise the saturating branches (m >= 100, e^m overflow)
    p['Z'][0] *= 200.0
    p['Z'][1] *= -200.0
    p['Z'][2] = 0.0
    return p


# ------------------------------------------------------------------ G2 inputs --
def snnls_matrix():
    """SURVEY 8c fingerprint problem (reference tests/test_snnls seed, :8)."""
    np.random.seed(324)
    return np.random.randn(1000, 100)


def snnls_small_cases():
    """Data generators of reference tests/test_snnls/test_deterministic.py:18-35."""
    r = np.random.RandomState(324)
    cases = []
    for n, d in ((10, 3), (100, 10), (10, 10)):
        cases.append(('gauss_%d_%d_' % (n, d), r.randn(n, d)))
        b = (r.rand(n, d) > 0.5).astype(np.float64)
        b[b.sum(axis=1) == 0, 0] = 1.
        cases.append(('bin_%d_%d_' % (n, d), b))
        g = r.randn(n, d)
        cases.append(('colinear_%d_%d_' % (n, d), np.vstack((g, g.sum(axis=0)[np.newaxis, :]))))
        ax = np.zeros((n, d))
        ax[np.arange(n), r.randint(d, size=n)] = 1. + r.rand(n)
        cases.append(('axis_%d_%d_' % (n, d), ax))
    return cases


# ------------------------------------------------------------- coreset builds --
def make_logistic(N, D, seed, zero_rows=()):
    """SURVEY 8d generator: X ~ N(0, I), theta* = 1/sqrt(D), 10% label flips, Z = y X."""
    def make():
        r = np.random.RandomState(seed)
        X = r.randn(N, D)
        th = np.ones(D)/np.sqrt(D)
        y = np.where(r.rand(N) < 1./(1.+np.exp(-X.dot(th))), 1., -1.)
        y[r.rand(N) < 0.1] *= -1.
        Z = y[:, np.newaxis]*X
        for i in zero_rows:
            Z[i] = 0.
        mu0 = np.zeros(D)

        def sampler(S, w, pts):
            if pts.shape[0] == 0:
                w = np.zeros(1)
                pts = np.zeros((1, D))
            mu, L, _ = om.lr_laplace(w, pts, mu0)
            return mu + np.random.randn(S, mu.shape[0]).dot(L.T)
        return dict(model='lr', data=Z, sampler=sampler, params={},
                    ref_betalik=lambda lr, ga, nl: lr.beta_likelihood,
                    ref_loglik=lambda lr, ga, nl: lr.log_likelihood,
                    oracle_betalik=lambda beta: (lambda pts, th: om.lr_betalik(pts, th, beta)),
                    oracle_loglik=lambda: om.lr_loglik,
                    ref_gradll=lambda lr, ga, nl: lr.grad_z_log_likelihood,
                    oracle_gradll=lambda: om.lr_grad_z_loglik)
    return make


def make_gaussian(N, d, seed):
    """examples/zellner_gaussian/main.py:33-54 scaled down: inliers + three outlier clusters."""
    def make():
        r = np.random.RandomState(seed)
        Sig = 50.*np.eye(d) + 5.*np.ones((d, d))
        L = np.linalg.cholesky(Sig)
        X = r.randn(N, d).dot(L.T)
        Xo = np.concatenate((200. + r.randn(N//50, d).dot(L.T)*np.sqrt(.5),
                             150. + r.randn(N//50, d).dot(L.T)*np.sqrt(.1),
                             r.randn(N//10, d).dot(L.T)*np.sqrt(10.)))
        data = np.concatenate((X, Xo))
        Siginv = np.linalg.inv(Sig)
        logdetSig = np.linalg.slogdet(Sig)[1]
        mu0 = np.zeros(d)
        Sig0inv = np.eye(d)

        def sampler(S, w, pts):
            if pts.shape[0] == 0:
                w = np.zeros(1)
                pts = np.zeros((1, d))
            mu, Lp, _ = om.gauss_weighted_post(mu0, Sig0inv, Siginv, pts, w)
            return mu + np.random.randn(S, mu.shape[0]).dot(Lp.T)
        return dict(model='gauss', data=data, sampler=sampler, params=dict(Siginv=Siginv, logdetSig=logdetSig), prior=dict(mu0=mu0, Sig0inv=Sig0inv),
                    ref_betalik=lambda lr, ga, nl: (lambda x, th, beta: ga.gaussian_beta_likelihood(x, th, beta, Siginv, logdetSig)),
                    ref_loglik=lambda lr, ga, nl: (lambda x, th: ga.gaussian_loglikelihood(x, th, Siginv, logdetSig)),
                    ref_gradll=lambda lr, ga, nl: (lambda x, th: ga.gaussian_grad_x_loglikelihood(x, th, Siginv)),
                    oracle_gradll=lambda: (lambda x, th: om.gauss_grad_x_loglik(x, th, Siginv)),
                    oracle_betalik=lambda beta: (lambda pts, th: om.gauss_betalik(pts, th, beta, Siginv, logdetSig)),
                    oracle_loglik=lambda: (lambda pts, th: om.gauss_loglik(pts, th, Siginv, logdetSig)))
    return make


def make_neurlin(N, D, seed):
    """SURVEY 8d C4 scaled down: random-init relu features, y = Phi w* + noise, 10% rows' y ~ N(10, .5^2)."""
    def make():
        r = np.random.RandomState(seed)
        X = r.randn(N, 6)
        W1 = r.randn(6, 16)/np.sqrt(6.)
        W2 = r.randn(16, D)/np.sqrt(16.)
        Phi = np.maximum(np.maximum(X.dot(W1), 0.).dot(W2), 0.)
        wstar = r.randn(D)
        sigsq = 0.5
        y = Phi.dot(wstar) + np.sqrt(sigsq)*r.randn(N)
        bad = r.rand(N) < 0.1
        y[bad] = 10. + .5*r.randn(bad.sum())
        Z = np.hstack((Phi, y[:, np.newaxis]))
        mu0 = np.zeros(D)
        Sig0inv = np.eye(D)

        def sampler(S, w, pts):
            if pts.shape[0] == 0:
                w = np.zeros(1)
                pts = np.zeros((1, D+1))
            mu, Lp, _ = om.nl_weighted_post(mu0, Sig0inv, sigsq, pts, w)
            return mu + np.random.randn(S, mu.shape[0]).dot(Lp.T)
        return dict(model='nl', data=Z, sampler=sampler, params=dict(sigsq=sigsq), prior=dict(mu0=mu0, Sig0inv=Sig0inv),
                    ref_betalik=lambda lr, ga, nl: (lambda z, th, beta: nl.neurlinr_beta_likelihood(z, th, beta, sigsq)),
                    ref_loglik=lambda lr, ga, nl: (lambda z, th: nl.neurlinr_loglikelihood(z, th, sigsq)),
                    oracle_betalik=lambda beta: (lambda pts, th: om.nl_betalik(pts, th, beta, sigsq)),
                    oracle_loglik=lambda: (lambda pts, th: om.nl_loglik(pts, th, sigsq)))
    return make


def make_c1_zellner_gaussian():
    """BASELINE config 1 -- examples/zellner_gaussian/main.py with its run.sh arguments `BCORES 0`, restated line by line up
    to the point where the driver starts its build loop: the driver seeds numpy's GLOBAL stream once (main.py:14) and
    everything -- data (:43,:51-53), the perturbed proposal (:79-85), the constructor draws of the four projectors
    (:75,:87,:94,:95) -- consumes it in order.  The case runner constructs the last projector (prj_bw) itself, so the
    stream is left positioned right before that draw; `seed` is None so that nothing re-seeds it."""
    def make():
        np.random.seed(0)                                  # main.py:13-14, tr = 0
        N, d = 5000, 100
        mu0 = np.zeros(d)
        Sig0 = np.eye(d)
        Sig = 500*np.eye(d)
        th = np.zeros(d)
        Sig0inv = np.linalg.inv(Sig0)
        Siginv = np.linalg.inv(Sig)
        logdetSig = np.linalg.slogdet(Sig)[1]
        X = np.random.multivariate_normal(th, Sig, N)      # :43
        mup, LSigp, _ = om.gauss_weighted_post(mu0, Sig0inv, Siginv, X, np.ones(X.shape[0]))
        Sigp = LSigp.dot(LSigp.T)
        X1 = np.random.multivariate_normal(th+200, 0.5*Sig, int(N/50.))   # :51
        X2 = np.random.multivariate_normal(th+150, 0.1*Sig, int(N/50.))   # :52
        X3 = np.random.multivariate_normal(th, 10*Sig, int(N/10.))        # :53
        data = np.concatenate((X, X1, X2, X3))
        np.random.randn(200, d)                            # prj_optimal's constructor draw (:75; projector.py:18)
        U = np.random.rand()                               # :78
        muhat = U*mup + (1.-U)*mu0
        muhat += 0.75*np.sqrt((muhat**2).sum())*np.random.randn(muhat.shape[0])    # :82
        np.random.randn()                                  # :83
        np.random.randn(200, d)                            # prj_realistic's constructor draw (:87)
        np.random.randn(200, d)                            # prj_w's constructor draw (:94): sampler_w(200, [], [])

        def sampler(S, w, pts):                            # sampler_w, main.py:88-92
            if pts.shape[0] == 0:
                w = np.zeros(1)
                pts = np.zeros((1, d))
            mu, Lp, _ = om.gauss_weighted_post(mu0, Sig0inv, Siginv, pts, w)
            return mu + np.random.randn(S, mu.shape[0]).dot(Lp.T)
        return dict(model='gauss', data=data, sampler=sampler, params=dict(Siginv=Siginv, logdetSig=logdetSig), prior=dict(mu0=mu0, Sig0inv=Sig0inv),
                    ref_betalik=lambda lr, ga, nl: (lambda x, th, beta: ga.gaussian_beta_likelihood(x, th, beta, Siginv, logdetSig)),
                    ref_loglik=lambda lr, ga, nl: (lambda x, th: ga.gaussian_loglikelihood(x, th, Siginv, logdetSig)),
                    ref_gradll=lambda lr, ga, nl: (lambda x, th: ga.gaussian_grad_x_loglikelihood(x, th, Siginv)),
                    oracle_gradll=lambda: (lambda x, th: om.gauss_grad_x_loglik(x, th, Siginv)),
                    oracle_betalik=lambda beta: (lambda pts, th: om.gauss_betalik(pts, th, beta, Siginv, logdetSig)),
                    oracle_loglik=lambda: (lambda pts, th: om.gauss_loglik(pts, th, Siginv, logdetSig)))
    return make


# first selected rows of the UNMODIFIED driver `python3 main.py BCORES 0` (619 s on one CPU thread)
C1_DRIVER_FIRST_ROWS = [3650, 3950, 435, 3281, 1108, 933, 1640, 4214, 4372, 1164, 30, 3726]


def reseed(case):
    if case['seed'] is not None:
        np.random.seed(case['seed'])


def _sched(i0):
    return lambda i: i0/(1.+i)


def contiguous_groups(n, size):
    return [list(range(i, min(i+size, n))) for i in range(0, n, size)]


def ragged_groups(n, seed):
    """a random partition of the rows into groups of 5..20 rows (row order shuffled)"""
    r = np.random.RandomState(seed)
    perm = r.permutation(n)
    out, i = [], 0
    while i < n:
        k = int(r.randint(5, 21))
        out.append([int(v) for v in perm[i:i+k]])
        i += k
    return out


def build_size(case, m):
    """`sz` argument of build(1, sz) at step m: group mode adds whole groups, so the reference's only guard is
    sz >= current size (coreset.py:38-39)"""
    return 10**9 if case.get('groups') is not None else m


def coreset_cases(heavy=True):
    c = []
    base = dict(n_sel=None, n_opt=None, beta=0.1, solver=None, groups=None)
    c.append(dict(base, name='lr_beta_small', alg='beta', make=make_logistic(2000, 5, 3), seed=1, S=50, opt_itrs=20, M=6, sched=_sched(1.)))
    c.append(dict(base, name='lr_svi_small', alg='svi', make=make_logistic(2000, 5, 3), seed=1, S=50, opt_itrs=20, M=6, sched=_sched(1.)))
    c.append(dict(base, name='lr_beta_zero_rows', alg='beta', make=make_logistic(1500, 4, 5, zero_rows=(0, 7, 450, 1499)), seed=2, S=64, opt_itrs=10, M=5, sched=_sched(1.)))
    # zero rows met only now and then (sub-sampled selection): the NaN-first arg-max / NaN gate rules (bcores.py:78-81) fire on
    # some selections while the coreset keeps growing, so the weights that follow a NaN event are pinned too
    c.append(dict(base, name='lr_beta_zero_rows_sub', alg='beta', make=make_logistic(1500, 4, 5, zero_rows=tuple(range(0, 1500, 100))), seed=3, S=64,
                  opt_itrs=10, M=10, sched=_sched(1.), n_sel=150))
    c.append(dict(base, name='lr_beta_sub', alg='beta', make=make_logistic(3000, 6, 7), seed=4, S=40, opt_itrs=25, M=6, sched=_sched(1.), n_sel=300, n_opt=100))
    c.append(dict(base, name='gauss_beta_sub', alg='beta', make=make_gaussian(500, 10, 0), seed=5, S=40, opt_itrs=30, M=8, sched=_sched(1.), n_sel=100, n_opt=40, beta=0.01))
    c.append(dict(base, name='gauss_svi_full', alg='svi', make=make_gaussian(400, 7, 2), seed=6, S=32, opt_itrs=15, M=5, sched=_sched(.1)))
    c.append(dict(base, name='nl_beta_small', alg='beta', make=make_neurlin(1000, 8, 9), seed=7, S=32, opt_itrs=20, M=5, sched=_sched(.1), beta=0.2))
    c.append(dict(base, name='nl_svi_small', alg='svi', make=make_neurlin(1000, 8, 9), seed=7, S=32, opt_itrs=20, M=5, sched=_sched(.1)))
    c.append(dict(base, name='lr_hilbert_giga', alg='hilbert', make=make_logistic(500, 5, 13), seed=8, S=50, opt_itrs=0, M=10, sched=None, solver='GIGA'))
    c.append(dict(base, name='lr_hilbert_fw_sub', alg='hilbert', make=make_logistic(800, 5, 14), seed=9, S=50, opt_itrs=0, M=10, sched=None, solver='FrankWolfe', n_sel=200))
    c.append(dict(base, name='gauss_hilbert_omp', alg='hilbert', make=make_gaussian(300, 6, 4), seed=10, S=48, opt_itrs=0, M=7, sched=None, solver='OrthoPursuit'))
    # group-wise selection (bcores.py:46-61,91-123 / sparsevi.py:93-126): whole groups of rows enter the coreset
    c.append(dict(base, name='lr_beta_groups', alg='beta', make=make_logistic(600, 5, 21), seed=11, S=40, opt_itrs=15, M=4, sched=_sched(1.),
                  groups=contiguous_groups(600, 15)))
    c.append(dict(base, name='lr_beta_groups_sub', alg='beta', make=make_logistic(900, 4, 22), seed=12, S=32, opt_itrs=12, M=4, sched=_sched(1.),
                  groups=contiguous_groups(900, 10), n_sel=25, n_opt=150))
    c.append(dict(base, name='nl_svi_groups', alg='svi', make=make_neurlin(480, 6, 23), seed=13, S=32, opt_itrs=12, M=3, sched=_sched(.1),
                  groups=ragged_groups(480, 17)))
    # pseudo-coresets (bpsvi.py): weights and point locations optimised jointly; M = number of pseudo-points, one build
    c.append(dict(base, name='lr_bpsvi', alg='bpsvi', make=make_logistic(800, 5, 31), seed=14, S=48, opt_itrs=25, M=6, sched=_sched(.5)))
    c.append(dict(base, name='gauss_bpsvi_sub', alg='bpsvi', make=make_gaussian(400, 6, 8), seed=15, S=40, opt_itrs=20, M=5, sched=_sched(.2),
                  n_opt=120))
    if heavy:
        # BASELINE config 1: the reference's own driver configuration (N = 5700, d = 100, S = 200, 1000 ADAM steps per point,
        # sub-samples 1000 / 200, step 0.1/(1+i), beta = 0.1), first 12 of its 200 build iterations
        c.append(dict(base, name='c1_zellner_gaussian', alg='beta', make=make_c1_zellner_gaussian(), seed=None, S=200, opt_itrs=1000, M=12,
                      sched=_sched(.1), n_sel=1000, n_opt=200, beta=0.1, heavy=True))
        # SURVEY 8c "logistic mini" fingerprint shape (N=10000, D=10, S=100, opt_itrs=50, M=10)
        c.append(dict(base, name='lr_beta_mini', alg='beta', make=make_logistic(10000, 10, 0), seed=1, S=100, opt_itrs=50, M=10, sched=_sched(1.), heavy=True))
        # BASELINE config 3 at its stated size (examples/zellner_logreg: N = 100K, D = 20, S = 100, beta = 0.1, full data)
        c.append(dict(base, name='c3_logreg_100k', alg='beta', make=make_logistic(100000, 20, 41), seed=21, S=100, opt_itrs=20, M=5,
                      sched=_sched(1.), heavy=True))
        # BASELINE config 4 at N = 100K of its 1M (examples/zellner_neural_linear: D = 64 random-init relu features, S = 256,
        # conjugate sampler of the weighted posterior, 10 % of the targets replaced by outliers)
        c.append(dict(base, name='c4_neurlin_100k', alg='beta', make=make_neurlin(100000, 64, 42), seed=22, S=256, opt_itrs=20, M=5,
                      sched=_sched(.1), beta=0.2, heavy=True))
        # the north-star SHAPE (D = 128, S = 1024, beta = 0.1, full data) at a row count the reference finishes in seconds:
        # four build steps on the tensor-core route's home ground
        c.append(dict(base, name='c5_northstar_16k', alg='beta', make=make_logistic(16384, 128, 43), seed=23, S=1024, opt_itrs=5, M=4,
                      sched=_sched(1.), heavy=True))
    return c
