"""Row sharding across ranks -- host-side logic only, no CUDA in this file.

Rows are split in contiguous blocks; per optimiser step the ranks exchange one (2 x (S+1))
double-double column-sum part each, per selection additionally one (score, position) pair.
Nothing of size N crosses a link.  The merge rules here reproduce numpy's arg-max semantics
(first NaN wins, then the larger value, then the lower position) independently of rank order.
"""
import math
import numpy as np


def partition_rows(n_total, world, rank):
    """contiguous block [row0, row0 + n_local) of rank `rank`; blocks differ by at most one row"""
    base, rem = divmod(int(n_total), int(world))
    row0 = rank*base + min(rank, rem)
    return row0, base + (1 if rank < rem else 0)


def owner_of(row, n_total, world):
    base, rem = divmod(int(n_total), int(world))
    cut = rem*(base+1)
    if row < cut:
        return int(row // (base+1))
    return int(rem + (row-cut) // base)


def local_subsample(sub_idcs, row0, n_local):
    """positions (in the global subsample list) and local row numbers of the subsampled rows this
    rank owns, in list order (bcores.py:53 draws WITH replacement: duplicates are kept)."""
    sub_idcs = np.asarray(sub_idcs, dtype=np.int64)
    pos = np.nonzero((sub_idcs >= row0) & (sub_idcs < row0+n_local))[0].astype(np.int64)
    return pos, sub_idcs[pos]-row0


def better(va, ia, vb, ib):
    """is candidate (va, ia) preferred over (vb, ib) under np.argmax semantics?  i < 0 = empty."""
    if ib < 0:
        return ia >= 0
    if ia < 0:
        return False
    na, nb = math.isnan(va), math.isnan(vb)
    if na != nb:
        return na
    if na:
        return ia < ib
    if va != vb:
        return va > vb
    return ia < ib


def merge_best(cands):
    """cands: iterable of (score, position); returns the np.argmax winner (score, position)"""
    bv, bi = 0.0, -1
    for v, i in cands:
        v, i = float(v), int(i)
        if better(v, i, bv, bi):
            bv, bi = v, i
    return bv, bi


def nan_max(vals):
    """np.max semantics over host floats (NaN propagates); -inf for an empty list"""
    out = -math.inf
    for v in vals:
        v = float(v)
        if math.isnan(v) or math.isnan(out):
            out = math.nan
        elif v > out:
            out = v
    return out


class Comm(object):
    """Thin view of torch.distributed (or of a single process)."""

    def __init__(self, dist=None):
        self.dist = dist
        self.world = dist.get_world_size() if dist is not None else 1
        self.rank = dist.get_rank() if dist is not None else 0

    @classmethod
    def current(cls):
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                return cls(dist)
        except ImportError:
            pass
        return cls(None)

    def _staged(self, t):
        """gloo moves host memory only: CUDA tensors are staged through the host (two ranks sharing one GPU in the tests;
        NCCL, the production backend, takes device tensors directly)"""
        return t.is_cuda and self.dist.get_backend() == 'gloo'

    def allgather(self, t):
        """(world, *t.shape) tensor of every rank's `t` (same device/dtype), rank-ordered"""
        if self.world == 1:
            return t.unsqueeze(0)
        import torch
        if self._staged(t):
            return self.allgather(t.cpu()).to(t.device)
        out = torch.empty((self.world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        if t.is_cuda:
            self.dist.all_gather_into_tensor(out, t.contiguous())
        else:
            parts = [torch.empty_like(t) for _ in range(self.world)]
            self.dist.all_gather(parts, t.contiguous())
            for r, p in enumerate(parts):
                out[r].copy_(p)
        return out

    def allgather_var(self, t):
        """list of every rank's 1-d tensor `t`, rank-ordered; the lengths may differ"""
        if self.world == 1:
            return [t]
        import torch
        n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
        sizes = [int(v) for v in self.allgather(n).cpu().numpy()[:, 0]]
        pad = torch.zeros(max(max(sizes), 1), dtype=t.dtype, device=t.device)
        pad[:t.numel()] = t
        allp = self.allgather(pad)
        return [allp[r, :sizes[r]] for r in range(self.world)]

    def broadcast(self, t, src):
        if self.world > 1:
            if self._staged(t):
                h = t.cpu()
                self.dist.broadcast(h, src=src)
                t.copy_(h)
            else:
                self.dist.broadcast(t, src=src)
        return t
