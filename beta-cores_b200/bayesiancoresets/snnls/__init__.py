"""Sparse non-negative least-squares solvers over a materialised (S, N) matrix: greedy geodesic ascent, Frank-Wolfe and
orthogonal pursuit score every datapoint with one device pass per iteration (csrc/bc_dense.cu); the two sampling
baselines are host RNG bookkeeping.  Public names as in the reference's `bayesiancoresets.snnls`."""
from .snnls import SparseNNLS
from .giga import GIGA
from .frankwolfe import FrankWolfe
from .orthopursuit import OrthoPursuit
from .sampling import ImportanceSampling, UniformSampling

__all__ = ['SparseNNLS', 'GIGA', 'FrankWolfe', 'OrthoPursuit', 'ImportanceSampling', 'UniformSampling']
