"""CPU checks of the error-free integer split behind the tensor-core route (csrc/bc_project_q.cu): the digit
decomposition and the truncated diagonal recombination are restated here in exact integer arithmetic, next to the
operand-image geometry compiled from the kernel's own header.  No GPU needed."""
import ctypes
import math
import os
import subprocess
from fractions import Fraction

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'native', 'qsplit_host.cpp')
OUT = os.path.join(HERE, 'native', '_build', 'libqsplit_host.so')
K = 7     # digits


@pytest.fixture(scope='module')
def geo():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.run(['g++', '-O2', '-std=c++17', '-fPIC', '-shared', '-I/usr/local/cuda/include', SRC, '-o', OUT], check=True)
    L = ctypes.CDLL(OUT)
    L.q_off.argtypes, L.q_off.restype = [ctypes.c_uint, ctypes.c_uint], ctypes.c_uint
    L.np_sum_const.argtypes, L.np_sum_const.restype = [ctypes.c_double, ctypes.c_int], ctypes.c_double
    L.np_score_const.argtypes, L.np_score_const.restype = [ctypes.c_double, ctypes.c_int, ctypes.c_double], ctypes.c_double
    return L


def split(vec):
    """k_quantise: one power-of-two exponent per vector, 7 signed base-256 digits per entry (top digit within +-64)"""
    amax = float(np.abs(vec).max())
    e = (math.frexp(amax)[1] - 1) + 3 if amax > 0 else 0          # ilogb(amax) + 3
    digs = np.zeros((K, len(vec)), dtype=np.int64)
    for i, x in enumerate(vec):
        Q = int(round(Fraction(float(x))*Fraction(2)**(56-e)))   # round-half-even on an exact rational = __double2ll_rn
        q = Q
        for s in range(K-1, 0, -1):
            d = ((q + 128) % 256) - 128                           # sign-extended low byte
            digs[s, i] = d
            q = (q - d) >> 8
        digs[0, i] = q
        assert sum(int(digs[s, i])*256**(K-1-s) for s in range(K)) == Q
    return e, digs


def test_digits_are_int8_and_exact():
    r = np.random.RandomState(0)
    for trial in range(40):
        v = r.randn(128)*math.exp(6*r.randn())
        if trial % 5 == 0:
            v[r.randint(128, size=20)] = 0.
        e, d = split(v)
        assert np.abs(d[0]).max() <= 64 and d[1:].min() >= -128 and d[1:].max() <= 127
        # the digits reproduce the entries to 2^-56 of the scale 2^e (at most 2^-53 of the largest entry)
        back = sum(d[s].astype(object)*Fraction(256)**(-(s+1)) for s in range(K))
        err = max(abs(Fraction(float(x)) - b*Fraction(2)**e) for x, b in zip(v, back))
        assert err <= Fraction(2)**(e-57)


def test_truncated_diagonals_reproduce_the_fp64_contraction():
    """sum_{d<=6} 256^(6-d) D_d, D_d = sum_{i+j=d} a_i . b_j in int32 range, times 2^(e_x + e_th - 64): within 2e-14 of the
    exact dot product relative to max|x| max|theta| (numpy's own dgemm is at ~5e-15)"""
    r = np.random.RandomState(1)
    worst = 0.
    for trial in range(30):
        D = int(r.choice([1, 5, 20, 64, 128]))
        x = r.randn(D)*math.exp(3*r.randn())
        th = r.randn(D)*math.exp(r.randn())
        ex, a = split(x)
        et, b = split(th)
        H = 0
        for d in range(K):
            Dd = sum(int(np.dot(a[i], b[d-i])) for i in range(d+1))
            assert abs(Dd) < 2**24                      # what the kernel keeps in an int32 TMEM accumulator
            H += Dd*256**(K-1-d)
        hi = sum(sum(int(np.dot(a[i], b[d-i])) for i in range(d+1))*256**(3-d) for d in range(4))
        assert abs(hi) < 2**51                          # exact_ll2d's range (q_combine's upper half)
        got = Fraction(H)*Fraction(2)**(ex+et-64)
        exact = sum(Fraction(float(u))*Fraction(float(v)) for u, v in zip(x, th))
        worst = max(worst, float(abs(got-exact))/(np.abs(x).max()*np.abs(th).max()))
    assert worst < 2e-14, worst


def test_swizzled_image_geometry(geo):
    """q_swizzle_off is a bijection of a digit plane onto itself, keeps 16-byte chunks intact inside their 128-byte row
    and XORs the chunk index with (row mod 8) -- the SWIZZLE_128B pattern the UMMA descriptors declare"""
    rows, kk = geo.q_tile_rows(), geo.q_k()
    assert (geo.q_slices(), rows, geo.q_chunk(), kk) == (7, 128, 32, 128)
    seen = set()
    for rr in range(rows):
        for c in range(kk):
            o = geo.q_off(rr, c)
            assert o // 128 == rr and o % 16 == c % 16
            assert (o % 128) // 16 == (c // 16) ^ (rr % 8)
            seen.add(o)
    assert len(seen) == rows*kk and max(seen) == rows*kk - 1


def test_constant_rows_are_centred_like_numpy(geo):
    """a row of S identical values: the reference's `x -= x.mean(axis=1)` leaves 0 or one ulp depending on how numpy's
    pairwise summation rounds -- which decides between a NaN and a finite correlation (bcores.py:78).  The kernels use
    bc::np_sum_const; it must agree with numpy bit for bit for every S."""
    r = np.random.RandomState(7)
    for trial in range(4000):
        S = int(r.choice([1, 2, 3, 7, 8, 9, 31, 32, 33, 40, 64, 100, 127, 128, 129, 136, 200, 255, 256, 257, 500, 1000, 1024, 1025, 2000, 4096, 5000]))
        x = float(r.randn()*10.**r.uniform(-8, 8))
        A = np.full((3, S), x)
        assert geo.np_sum_const(x, S) == A.sum(axis=1)[1], (x, S)
        V = A - A.mean(axis=1)[:, np.newaxis]
        res = r.randn(S)
        with np.errstate(all='ignore'):
            want = (V.dot(res)/np.sqrt((V**2).sum(axis=1))/S)[1]
        got = geo.np_score_const(x, S, float(res.sum()))
        assert (np.isnan(want) and np.isnan(got)) or np.isclose(got, want, rtol=1e-9, atol=0), (x, S, got, want)


def test_recombination_forms_round_alike():
    """q_combine_n (csrc/bc_umma.cuh): for up to 6 digits the diagonals are summed in ONE int64 and converted once; the 7-digit
    split sums two halves and joins them with one FMA.  Both are a single rounding of the same integer -- restated here in
    exact arithmetic over the diagonals' full range (|D_d| <= (d+1) 128 * 128 * 128) -- and the int64 never overflows."""
    r = np.random.RandomState(11)
    for NS in (5, 6):
        for _ in range(4000):
            big = r.rand() < 0.3
            d = [int(r.randint(-(k+1)*2**21, (k+1)*2**21+1)) if not big else int(((k+1)*2**21)*r.choice([-1, 1])) for k in range(NS)]
            exact = sum(dk*256**(NS-1-k) for k, dk in enumerate(d))
            assert abs(exact) < 2**62
            # the kernel's integer steps: four trailing digits by multiply-add, the leading ones into the high word
            t = d[NS-1] + d[NS-2]*256 + d[NS-3]*65536 + d[NS-4]*16777216
            hi = (t >> 32) + d[NS-5] + ((d[NS-6] << 8) if NS >= 6 else 0)
            t = ((hi & 0xffffffff) << 32) | (t & 0xffffffff)
            if t >= 2**63:
                t -= 2**64
            assert t == exact
            one = float(np.float64(t))                                  # I2F.F64.S64, round to nearest
            h = sum(dk*256**(NS-4-k) for k, dk in enumerate(d[:NS-3]))   # the two-half form of the same digits
            lo = d[NS-3]*65536 + d[NS-2]*256 + d[NS-1]
            two = float(Fraction(h)*16777216 + Fraction(lo))            # fma(h, 2^24, lo): one rounding of the exact sum
            assert one == two == float(Fraction(exact))
