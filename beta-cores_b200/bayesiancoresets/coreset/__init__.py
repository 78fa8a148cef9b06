from .hilbert import HilbertCoreset
from .sampling import UniformSamplingCoreset
from .sparsevi import SparseVICoreset
from .projector import BlackBoxProjector, Projector, BetaBlackBoxProjector
from .bpsvi import BatchPSVICoreset, DiffPrivBatchPSVICoreset
from .bcores import BetaCoreset
