"""HilbertCoreset: project once, then sparse non-negative least squares in the S-dimensional tangent space
(drop-in for bayesiancoresets/coreset/hilbert.py:7-43).

With a bound DevicePotential log-likelihood the (n, S) matrix is written by the materialise kernel straight
into HBM, datapoint-major, together with its row norms and column sum, and the solver is attached to it
without a host round trip.  Opaque callbacks go through project() (host matrix) like the reference.
"""
import numpy as np
import torch

from .. import _native as nv
from .._device import Engine, DeviceRows, ptr, stream_ptr
from .._shard import Comm, partition_rows
from ..util import rng
from ..snnls.giga import GIGA
from ..snnls.snnls import SparseNNLS
from .coreset import Coreset


class HilbertCoreset(Coreset):
    def __init__(self, data, ll_projector, n_subsample=None, snnls=GIGA, **kw):
        if n_subsample is None:
            sub_idcs = None
        else:
            n_subsample = min(data.shape[0], n_subsample)
            sub_idcs = rng.randint(data.shape[0], n_subsample)          # hilbert.py:14
        fused = ll_projector.fused(data.shape[1]) if hasattr(ll_projector, 'fused') else None
        device_solver = isinstance(snnls, type) and issubclass(snnls, SparseNNLS)
        if fused is not None and device_solver:
            eng = Engine.get()
            comm = Comm.current()
            sub = data if sub_idcs is None else data[sub_idcs]
            # datapoints sharded over the ranks in contiguous blocks; every rank holds the host rows, as the
            # reference's caller does, and projects its own block
            r0, nl = partition_rows(sub.shape[0], comm.world, comm.rank) if comm.world > 1 else (0, sub.shape[0])
            rows = DeviceRows(eng, sub[r0:r0+nl])
            fused.configure(None)
            fused.set_samples(ll_projector.samples)
            if nl:
                V, norms, _ = fused.materialise(rows, want_norms=True)
                nh = norms.cpu().numpy()
            else:       # more ranks than datapoints: this rank holds none
                V, norms, nh = eng.empty(0, fused.S), eng.empty(1), np.zeros(0)
            keep = nh > 0.                                                           # hilbert.py:16
            nk = int(keep.sum())
            if not keep.all():
                idx = eng.upload(np.nonzero(keep)[0].astype(np.int64), dtype=torch.int64)
                Vk = eng.empty(max(nk, 1), V.shape[1])
                nv.call('bc_dense_gather', eng.ctx(), ptr(V), int(V.stride(0)), V.shape[1], ptr(idx), nk, ptr(Vk),
                        V.shape[1], stream_ptr())
                V = Vk[:nk]
                norms = eng.upload(nh[keep]) if nk else eng.empty(1)
            S = V.shape[1]
            dd = eng.empty(2*(S+1))
            b = eng.empty(S)
            if nk:
                nv.call('bc_dense_colsum', eng.ctx(), ptr(V), nk, S, int(V.stride(0)), ptr(dd), stream_ptr())
            else:
                dd.zero_()
            alldd = comm.allgather(dd)                                              # one double-double part per rank, added in rank order
            nv.call('bc_colsum_combine', eng.ctx(), ptr(alldd), comm.world, S, ptr(b), stream_ptr())
            # the zero-norm filter shifts the indices (hilbert.py:16 then :32): positions count the KEPT rows of all lower ranks
            kept = [int(v) for v in comm.allgather(torch.tensor([nk], dtype=torch.int64, device=eng.device)).cpu().numpy()[:, 0]]
            row0 = sum(kept[:comm.rank])
            self.snnls = snnls.from_device(V, norms, b.cpu().numpy(), row0=row0, n_total=sum(kept))   # hilbert.py:17: snnls(vecs.T, vecs.sum(0))
        else:
            vecs = ll_projector.project(data if sub_idcs is None else data[sub_idcs])
            vecs = vecs[np.sqrt((vecs**2).sum(axis=1)) > 0., :]
            self.snnls = snnls(vecs.T, vecs.sum(axis=0))
        self.sub_idcs = sub_idcs
        self.data = data
        super().__init__(**kw)

    def reset(self):
        self.snnls.reset()
        super().reset()

    def _sync_from_solver(self):
        w = self.snnls.weights()
        self.wts = w[w > 0]
        self.idcs = self.sub_idcs[w > 0] if self.sub_idcs is not None else np.where(w > 0)[0]
        self.pts = self.data[self.idcs]

    def _build(self, itrs, sz):
        if self.snnls.size()+itrs > sz:
            raise ValueError('%s._build(): # itrs + current size cannot exceed total desired size sz. # itr = %s cur sz: %s '
                             'desired sz: %s' % (self.alg_name, itrs, self.snnls.size(), sz))
        self.snnls.build(itrs)
        self._sync_from_solver()

    def _optimize(self):
        self.snnls.optimize()
        self._sync_from_solver()

    def error(self):
        return self.snnls.error()
