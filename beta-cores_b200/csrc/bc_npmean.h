// beta-cores B200: how numpy rounds the mean of a row of identical values.  Plain C++ (also compiled by the CPU tests).
#pragma once
#include <math.h>
#if defined(__CUDACC__)
#define BC_NP_HD __host__ __device__ inline
#else
#define BC_NP_HD inline
#endif

namespace bc {

// ------------------------------------------- numpy's mean of a constant row --
// The reference centres a row with `bls -= bls.mean(axis=1)[:, None]` (projector.py:26,55).  For a row whose S values are
// all the same double x -- a zero data row, or a potential that has underflowed to its constant term for every sample:
// hundreds of the outlier rows of the Gaussian example -- the outcome depends on how numpy ROUNDS the sum of S copies
// of x: mean == x gives a zero row (0/0 = NaN correlation, which np.argmax then picks), mean != x gives a row of S equal
// values of one ulp, whose correlation is +-sum(resid)/S^1.5.  Which of the two happens decides the reference's
// selection, so the kernels reproduce it: np_sum_const is numpy's pairwise summation (8 interleaved accumulators up to
// 128 elements, halving above; numpy/core/src/umath/loops_utils.h.src) specialised to equal addends.
BC_NP_HD double np_sum_leaf(double x, int n) {   // n <= 128: numpy's unrolled-by-8 loop
  if (n < 8) {
    double r = 0.0;
    for (int i = 0; i < n; ++i) r += x;
    return r;
  }
  double rr = x;
  for (int i = 8; i < n - (n % 8); i += 8) rr += x;
  double res = ((rr + rr) + (rr + rr)) + ((rr + rr) + (rr + rr));
  for (int i = 0; i < n % 8; ++i) res += x;
  return res;
}
// sum(n) = sum(n2) + sum(n - n2), n2 = n/2 rounded down to a multiple of 8, down to leaves of <= 128 elements: the tree is
// walked in post-order with an explicit stack (a recursive device function would be an ABI call, which also makes ptxas
// drop the kernel's setmaxnreg register hand-over).  Depth <= log2(n / 64).
BC_NP_HD double np_sum_const(double x, int n) {
  if (n <= 128) return np_sum_leaf(x, n);
  int size[32];       // pending right-hand sizes
  double left[32];    // value of the left sibling once it is known
  bool have[32];
  int top = 0;
  int cur = n;
  double val = 0.0;
  while (true) {
    while (cur > 128) {            // descend along left children
      int n2 = cur / 2;
      n2 -= n2 % 8;
      size[top] = cur - n2;
      have[top] = false;
      ++top;
      cur = n2;
    }
    val = np_sum_leaf(x, cur);
    // climb: a finished left child starts its right sibling, a finished right child is added to the stored left value
    while (top > 0 && have[top - 1]) {
      val = left[top - 1] + val;
      --top;
    }
    if (top == 0) return val;
    left[top - 1] = val;
    have[top - 1] = true;
    cur = size[top - 1];
  }
}
// what the reference's centring leaves in every column of a row that is constantly x
BC_NP_HD double np_centred_const(double x, int S) { return x - np_sum_const(x, S) / (double)S; }
// its correlation with the residual, bcores.py:78: (v * sum r) / sqrt(sum of S copies of v^2) / S  (NaN when v == 0)
BC_NP_HD double np_score_const(double x, int S, double rsum) {
  const double v = np_centred_const(x, S);
  return (v * rsum) / sqrt(np_sum_const(v * v, S)) / (double)S;
}

}  // namespace bc
