"""CPU checks of the projection epilogue's scalar mathematics.

The kernels evaluate the potentials with restricted-domain polynomials (beta-cores_b200/csrc/bc_fastmath.cuh) and an
algebraic rewrite of the logistic beta-likelihood (bc_models.cuh).  The same headers compile as plain C++; here they are
built with g++ and compared with (a) mpmath at 100 digits and (b) the numpy oracle, on grids that include the tails.
No GPU needed; the polynomial fit itself is the host routine inside libbetacores.so (bc_fit_pow_poly).
"""
import ctypes
import os
import subprocess

import mpmath as mp
import numpy as np
import pytest

from oracle import np_models as om

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'native', 'fastmath_host.cpp')
OUT = os.path.join(HERE, 'native', '_build', 'libfm_host.so')
mp.mp.dps = 60


@pytest.fixture(scope='module')
def fm():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.run(['g++', '-O2', '-std=c++17', '-fPIC', '-shared', '-mfma', '-ffp-contract=off', SRC, '-o', OUT], check=True)
    L = ctypes.CDLL(OUT)
    d, i, dp = ctypes.c_double, ctypes.c_int, ctypes.POINTER(ctypes.c_double)
    L.fm_exp.argtypes, L.fm_exp.restype = [d], d
    L.fm_log1p_unit.argtypes, L.fm_log1p_unit.restype = [d], d
    L.fm_logistic.argtypes, L.fm_logistic.restype = [i, i, d, d, dp], d
    L.fm_gaussian.argtypes, L.fm_gaussian.restype = [i, d, d, d, dp], d
    L.fm_neurlin.argtypes, L.fm_neurlin.restype = [i, d, d, dp], d
    L.fm_logistic_v4.argtypes, L.fm_logistic_v4.restype = [i, i, dp, d, dp, dp], None
    return L


def _fit(beta, deg):
    from bayesiancoresets import _native as nv
    q = (ctypes.c_double*(deg+1))()
    err = ctypes.c_double()
    nv.call('bc_fit_pow_poly', float(beta), deg, q, ctypes.byref(err))
    full = (ctypes.c_double*25)()
    for k in range(deg+1):
        full[24-deg+k] = q[k]
    return full, err.value


def test_exp_polynomial_is_one_ulp(fm):
    r = np.random.RandomState(0)
    xs = np.concatenate([r.uniform(-700, 700, 4000), r.uniform(-2, 2, 4000), [-700., 700., 0., -0.0, 1e-300, -745., -1e6, 800.]])
    worst = 0.
    for x in xs:
        xc = min(max(x, -700.), 700.)
        want = mp.exp(mp.mpf(xc))
        got = mp.mpf(fm.fm_exp(float(x)))
        worst = max(worst, float(abs(got-want)/want))
    assert worst < 2.3e-16, worst
    assert np.isnan(fm.fm_exp(float('nan')))


def test_log1p_polynomial(fm):
    ts = np.concatenate([np.linspace(0., 1., 3001), 10.**np.linspace(-300, 0, 301)])
    worst = max(abs(mp.mpf(fm.fm_log1p_unit(float(t))) - mp.log1p(mp.mpf(float(t)))) for t in ts)
    assert worst < 3e-16, worst


@pytest.mark.parametrize('beta,deg', [(0.01, 20), (0.1, 20), (0.4, 20), (0.5, 24), (0.9, 24), (1.5, 24)])
def test_pow_polynomial_fit(beta, deg):
    q, err = _fit(beta, deg)
    assert err < 2.5e-17
    c = [mp.mpf(q[24-deg+k]) for k in range(deg+1)]
    worst = 0
    for t in np.linspace(0., 1., 2001):
        x = 2*mp.mpf(float(t)) - 1
        p = mp.mpf(0)
        for ck in c:
            p = p*x + ck
        worst = max(worst, abs(p - (1+mp.mpf(float(t)))**(-mp.mpf(beta))))
    assert worst < 1e-16, worst


def _lr_beta_exact(m, beta):
    m, b = mp.mpf(float(m)), mp.mpf(beta)
    em = mp.exp(m)
    return -(((b+1)/b)*(1+em)**(-b) - ((1+em)**(-b-1) + (1+mp.exp(-m))**(-b-1)))     # model_lr.py:85


@pytest.mark.parametrize('beta,poly', [(0.1, 20), (0.4, 20), (0.9, 24), (0.1, 0), (3.0, 0)])
def test_logistic_beta_likelihood_rewrite(fm, beta, poly):
    q, _ = _fit(beta, poly) if poly else ((ctypes.c_double*25)(), 0)
    r = np.random.RandomState(1)
    ms = np.concatenate([r.normal(0, 12, 3000), np.linspace(-60, 60, 1201), [0., -0.0, 700., -700., 745., -745., 5000., -5000.]])
    worst = 0.
    for m in ms:
        want = _lr_beta_exact(m, beta)
        got = mp.mpf(fm.fm_logistic(1, poly, float(-m), beta, q))
        worst = max(worst, float(abs(got-want)/max(1, abs(want))))
    # a few ulp of the largest intermediate term ((beta+1)/beta ~ 11 at beta = 0.1)
    assert worst < 4e-16*max(1., (beta+1)/beta), worst
    # numpy oracle (the reference's own expression, evaluated in double) agrees to its own rounding level
    Z = -ms[np.abs(ms) < 600][:, None]
    ref = om.lr_betalik(Z, np.ones((1, 1)), beta)[:, 0]
    got = np.array([fm.fm_logistic(1, poly, float(z), beta, q) for z in Z[:, 0]])
    assert np.allclose(got, ref, rtol=0, atol=2e-14*max(1., (beta+1)/beta))
    assert np.isnan(fm.fm_logistic(1, poly, float('nan'), beta, q))
    assert fm.fm_logistic(1, poly, float('inf'), beta, q) == pytest.approx(1.-(beta+1)/beta, abs=1e-15)    # m = -inf
    assert fm.fm_logistic(1, poly, float('-inf'), beta, q) == pytest.approx(1., abs=1e-15)                   # m = +inf


def test_logistic_loglik_rewrite(fm):
    ms = np.concatenate([np.random.RandomState(2).normal(0, 15, 3000), np.linspace(-50, 120, 851), [99.9, 100., 100.1, 700., -700., 1e4]])
    worst = 0.
    for m in ms:
        mm = mp.mpf(float(m))
        want = -mp.log1p(mp.exp(mm))
        got = mp.mpf(fm.fm_logistic(0, 0, float(-m), 0.1, None))
        worst = max(worst, float(abs(got-want)/max(1, abs(want))))
    assert worst < 4e-16, worst
    ref = om.lr_loglik(-ms[:, None], np.ones((1, 1)))[:, 0]
    got = np.array([fm.fm_logistic(0, 0, float(-m), 0.1, None) for m in ms])
    assert np.allclose(got, ref, rtol=1e-15, atol=1e-15)


@pytest.mark.parametrize('kind,beta,poly', [(1, 0.1, 20), (1, 0.9, 24), (1, 0.1, 0), (1, 3.0, 0), (0, 0.1, 0)])
def test_logistic_vector_forms(fm, kind, beta, poly):
    """evalv<4> -- the stage-interleaved form k_project_q runs (single clamp, integer sign test, no NaN selects) -- against
    60-digit arithmetic"""
    q = _fit(beta, poly)[0] if poly > 0 else (ctypes.c_double*25)()
    r = np.random.RandomState(4)
    ms = np.concatenate([r.normal(0, 12, 3000), np.linspace(-60, 60, 1201), [0., -0.0, 700., -700., 745., -745., 5000., -5000.]])
    if kind == 0:
        ms = np.concatenate([ms, [99.9, 100., 100.1, 1e4]])
    ms = ms[:4*(len(ms)//4)]
    worst = 0.
    out = (ctypes.c_double*4)()
    for k in range(0, len(ms), 4):
        c4 = (ctypes.c_double*4)(*[float(-m) for m in ms[k:k+4]])
        fm.fm_logistic_v4(kind, poly, c4, beta, q, out)
        for m, got in zip(ms[k:k+4], out):
            want = _lr_beta_exact(m, beta) if kind == 1 else -mp.log1p(mp.exp(mp.mpf(float(m))))
            worst = max(worst, float(abs(mp.mpf(got)-want)/max(1, abs(want))))
    assert worst < 5e-16*(max(1., (beta+1)/beta) if kind == 1 else 1.), worst
