"""Coreset construction classes (the public names of the reference's `bayesiancoresets.coreset` package).

Greedy variational classes share `_greedy.py` (fused device passes); `HilbertCoreset` materialises its projection once
and hands it to the `snnls` solvers; the projectors hold the user's sampler / likelihood callbacks.
"""
from .projector import Projector, BlackBoxProjector, BetaBlackBoxProjector
from .bcores import BetaCoreset
from .sparsevi import SparseVICoreset
from .bpsvi import BatchPSVICoreset, DiffPrivBatchPSVICoreset   # the latter: named by the reference, never shipped by it
from .hilbert import HilbertCoreset
from .sampling import UniformSamplingCoreset

__all__ = ['Projector', 'BlackBoxProjector', 'BetaBlackBoxProjector', 'BetaCoreset', 'SparseVICoreset', 'BatchPSVICoreset',
           'DiffPrivBatchPSVICoreset', 'HilbertCoreset', 'UniformSamplingCoreset']
