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


@pytest.mark.parametrize('kind,beta,poly', [(1, 0.01, 20), (1, 0.1, 20), (1, 0.9, 24), (1, 0.1, 0), (1, 3.0, 0), (0, 0.1, 0)])
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


# ---- lane-table forms (the tensor-core kernel's default for the logistic beta-likelihood) ----
def _fit_tab(beta):
    from bayesiancoresets import _native as nv
    w, rs, us, err = (ctypes.c_double*7)(), (ctypes.c_double*32)(), (ctypes.c_double*32)(), ctypes.c_double()
    nv.call('bc_fit_pow_tab', float(beta), w, rs, us, ctypes.byref(err))
    return w, rs, us, err.value


@pytest.fixture(scope='module')
def fmt(fm):
    d, i, dp = ctypes.c_double, ctypes.c_int, ctypes.POINTER(ctypes.c_double)
    fm.fm_exp_tab.argtypes, fm.fm_exp_tab.restype = [d, i], d
    fm.fm_logistic_v4_tab.argtypes, fm.fm_logistic_v4_tab.restype = [dp, d, dp, dp, dp, dp], None
    return fm


@pytest.mark.parametrize('lo', [0, 1])
def test_table_exp_is_one_ulp(fmt, lo):
    """exp_tab_v: 32-entry 2^(j/32) table + degree-6 expm1 polynomial.  The two-step reduction is within an ulp everywhere; the
    one-step form adds K * 1.7e-18 (K = 32 x / ln2), i.e. stays within an ulp for |x| < 1 and within 1e-17 ABSOLUTE always"""
    r = np.random.RandomState(0)
    xs = np.concatenate([r.uniform(-700, 0, 4000), r.uniform(-2, 0, 4000), -10.**np.linspace(-300, 2.8, 400), [-700., 0., -0.0]])
    worst_rel, worst_abs = 0., 0.
    for x in xs:
        want = mp.exp(mp.mpf(float(x)))
        got = mp.mpf(fmt.fm_exp_tab(float(x), lo))
        rel = float(abs(got-want)/want)
        if lo or abs(x) < 1:
            worst_rel = max(worst_rel, rel)
        worst_abs = max(worst_abs, float(abs(got-want)))
    assert worst_rel < 2.3e-16, worst_rel
    assert worst_abs < (1.2e-16 if lo else 1.2e-16), worst_abs
    # the degree-5 form the logistic potentials use (weighted minimax polynomial, one FP64 instruction less): 1.3 ulp
    fmt.fm_exp_tab5.argtypes, fmt.fm_exp_tab5.restype = [ctypes.c_double, ctypes.c_int], ctypes.c_double
    worst5 = 0.
    for x in xs:
        if lo or abs(x) < 1:
            want = mp.exp(mp.mpf(float(x)))
            worst5 = max(worst5, float(abs(mp.mpf(fmt.fm_exp_tab5(float(x), lo))-want)/want))
    assert worst5 < 3.2e-16, worst5


@pytest.mark.parametrize('beta', [0.01, 0.1, 0.4, 0.5, 0.9])
def test_pow_table_fit(beta):
    w, rs, us, err = _fit_tab(beta)
    assert err < 5e-17, err           # the acceptance bound of bc_set_potential
    mb = mp.mpf(beta)
    for j in range(32):
        sj = 1 + (mp.mpf(j) + mp.mpf(1)/2)/32
        assert abs(mp.mpf(rs[j]) - 1/sj) <= abs(1/sj)*mp.mpf(2)**-53
        assert abs(mp.mpf(us[j]) - sj**(-mb)) <= sj**(-mb)*mp.mpf(2)**-53
    worst = 0
    for x in np.linspace(-1/65., 1/65., 1001):
        ww = mp.mpf(float(x))
        q = mp.mpf(0)
        for ck in w:
            q = q*ww + mp.mpf(ck)
        worst = max(worst, abs(1 + ww*q - (1+ww)**(-mb)))
    assert worst < 5e-17, worst


def test_pow_table_is_refused_where_it_is_not_accurate():
    assert _fit_tab(3.0)[3] > 5e-17          # bc_set_potential then keeps the polynomial / exp-log forms


@pytest.mark.parametrize('beta', [0.01, 0.1, 0.5, 0.9])
def test_logistic_table_form(fmt, beta):
    """LogisticF<KIND_BETALIK, kPowTab>::evalv<4> (integer |c| and clamp, table exponentials, cubic reciprocal, table power)
    against 60-digit arithmetic: the same bound as the polynomial forms"""
    w, rs, us, err = _fit_tab(beta)
    r = np.random.RandomState(4)
    ms = np.concatenate([r.normal(0, 12, 3000), np.linspace(-60, 60, 1201), r.uniform(-1, 1, 800)*1e-3,
                         [0., -0.0, 700., -700., 745., -745., 5000., -5000., 1e-300, -1e-300]])
    ms = ms[:4*(len(ms)//4)]
    worst = 0.
    out = (ctypes.c_double*4)()
    for k in range(0, len(ms), 4):
        c4 = (ctypes.c_double*4)(*[float(-m) for m in ms[k:k+4]])
        fmt.fm_logistic_v4_tab(c4, beta, w, rs, us, out)
        for m, got in zip(ms[k:k+4], out):
            want = _lr_beta_exact(m, beta)
            worst = max(worst, float(abs(mp.mpf(got)-want)/max(1, abs(want))))
    assert worst < 5e-16*max(1., (beta+1)/beta), worst
    # a constant row must give one double: equal arguments, equal results, whatever the lane
    c4 = (ctypes.c_double*4)(0.3, 0.3, 0.3, 0.3)
    fmt.fm_logistic_v4_tab(c4, beta, w, rs, us, out)
    assert len(set(out)) == 1


def test_gaussian_and_neural_linear_vector_forms(fmt):
    """evalv<4> of the Gaussian (gaussian.py:34-62) and neural-linear (model_neurlinr.py:102-110) potentials -- the forms the
    tensor-core kernel runs, exponential from the lane table -- against 60-digit arithmetic and against the scalar forms"""
    d, i, dp = ctypes.c_double, ctypes.c_int, ctypes.POINTER(ctypes.c_double)
    fmt.fm_gaussian_v.argtypes, fmt.fm_gaussian_v.restype = [i, d, d, d, dp], d
    fmt.fm_neurlin_v.argtypes, fmt.fm_neurlin_v.restype = [i, d, d, dp], d
    r = np.random.RandomState(7)
    beta, dd, s2 = 0.3, 20, 1.7
    pg = (ctypes.c_double*8)(-3.2, 1/beta, -.5*beta, (1+beta)**(-.5*dd-1), 0.7, 1/beta**2, 1/(2*beta), (1+beta)**(-.5*dd-1)*np.log(1+beta))
    pn = (ctypes.c_double*8)(-.5*np.log(2*np.pi*s2), 1/(2*s2), (2*np.pi*s2)**(-beta/2), -(beta+1)/beta, -beta/(2*s2), 1/np.sqrt(1+beta), 0, 0)
    worst = 0.
    for _ in range(1500):
        c, ra, ca = r.normal(0, 5), abs(r.normal(0, 30)), abs(r.normal(0, 30))
        q = mp.mpf(ra) + mp.mpf(ca) - 2*mp.mpf(c)
        e = mp.exp(mp.mpf(pg[2])*q) if mp.mpf(pg[2])*q > -700 else mp.exp(-700)
        want = {0: mp.mpf(pg[0]) - q/2, 1: mp.mpf(pg[1])*e - mp.mpf(pg[3]),
                2: mp.mpf(pg[4])*(mp.mpf(pg[1])*e - mp.mpf(pg[3])) - mp.mpf(pg[5])*e - mp.mpf(pg[6])*q*e - mp.mpf(pg[7])}
        for kind in (0, 1, 2):
            got = fmt.fm_gaussian_v(kind, c, ra, ca, pg)
            scale = max(1, abs(want[kind]), float(abs(q))*float(e)*pg[6] if kind == 2 else 0)
            worst = max(worst, float(abs(mp.mpf(got) - want[kind])/scale))
            assert abs(got - fmt.fm_gaussian(kind, c, ra, ca, pg)) <= 4e-15*float(scale)
        y = r.normal(0, 3)
        r2 = mp.mpf(y)**2 - 2*mp.mpf(c)*mp.mpf(y) + mp.mpf(c)**2
        wl = mp.mpf(pn[0]) - mp.mpf(pn[1])*r2
        wb = mp.mpf(pn[2])*(mp.mpf(pn[3])*mp.exp(mp.mpf(pn[4])*r2) + mp.mpf(pn[5]))
        worst = max(worst, float(abs(mp.mpf(fmt.fm_neurlin_v(0, c, y, pn)) - wl)/max(1, abs(wl))))
        worst = max(worst, float(abs(mp.mpf(fmt.fm_neurlin_v(1, c, y, pn)) - wb)/max(1, abs(wb))))
        assert fmt.fm_neurlin_v(1, c, y, pn) == pytest.approx(fmt.fm_neurlin(1, c, y, pn), rel=1e-14, abs=1e-15)
    assert worst < 2e-15, worst


def test_logistic_loglik_table_form(fmt):
    """LogisticF<KIND_LOGLIK, kPowTab>::evalv<4>: -(max(m,0) + log1p(e^-|m|)) with the table exponential and the interval
    table of log(1+t), against 60-digit arithmetic (the bound of the polynomial form) and against the numpy oracle"""
    dp = ctypes.POINTER(ctypes.c_double)
    fmt.fm_logistic_loglik_v4_tab.argtypes, fmt.fm_logistic_loglik_v4_tab.restype = [dp, dp], None
    ms = np.concatenate([np.random.RandomState(2).normal(0, 15, 3000), np.linspace(-50, 120, 851), np.linspace(-1e-3, 1e-3, 201),
                         [0., -0.0, 99.9, 100., 100.1, 700., -700., 745., -745., 1e4, -1e4, 1e-300]])
    ms = ms[:4*(len(ms)//4)]
    out = (ctypes.c_double*4)()
    worst, got_all = 0., []
    for k in range(0, len(ms), 4):
        fmt.fm_logistic_loglik_v4_tab((ctypes.c_double*4)(*[float(-m) for m in ms[k:k+4]]), out)
        for m, got in zip(ms[k:k+4], out):
            want = -mp.log1p(mp.exp(mp.mpf(float(m))))
            worst = max(worst, float(abs(mp.mpf(got)-want)/max(1, abs(want))))
            got_all.append(got)
    assert worst < 4e-16, worst
    ref = om.lr_loglik(-ms[:, None], np.ones((1, 1)))[:, 0]
    assert np.allclose(got_all, ref, rtol=1e-15, atol=1e-15)
