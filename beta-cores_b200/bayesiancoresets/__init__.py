"""bayesiancoresets -- B200-native drop-in for the coreset-construction hot path of
dionman/beta-cores (same public names as bayesiancoresets/__init__.py:1 of the reference).

The arithmetic of the path (projection, scoring / arg-max, weight update) runs in
libbetacores.so (hand-written sm_100a CUDA behind the C ABI in include/betacores.h).  There is no
CPU fallback: constructing a coreset or solver without the library or without a CUDA device raises.
"""
from .coreset import (HilbertCoreset, SparseVICoreset, BatchPSVICoreset, DiffPrivBatchPSVICoreset, BetaCoreset,
                      UniformSamplingCoreset, BlackBoxProjector, Projector, BetaBlackBoxProjector)
from . import snnls
from . import util
from .potentials import DevicePotential
from ._device import DeviceRows, Engine
