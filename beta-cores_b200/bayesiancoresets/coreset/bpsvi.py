"""Pseudo-coreset constructions are outside the accelerated path (they need
d/dx kernels for the (n, S, D) gradient tensors of bpsvi.py:39,54).  The names are exported so that
`import bayesiancoresets` keeps the reference's surface (bayesiancoresets/__init__.py:1; the
reference itself ships no dpbpsvi.py), and say so when used."""
from .coreset import Coreset


class BatchPSVICoreset(Coreset):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError('BatchPSVICoreset is not part of the B200 hot path (pseudo-point gradients, SURVEY 8f.3)')


class DiffPrivBatchPSVICoreset(Coreset):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError('DiffPrivBatchPSVICoreset: the reference imports coreset/dpbpsvi.py but does not ship it')
