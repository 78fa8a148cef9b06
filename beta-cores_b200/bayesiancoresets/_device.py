"""Device-side plumbing shared by the projectors, coreset classes and solvers.

PyTorch supplies device memory, the current stream and (optionally) a torch.distributed process
group; the arithmetic is libbetacores.so's.  One `Engine` per (process, device): it owns the
bc_ctx workspaces and small result buffers.
"""
import ctypes
import numpy as np
import torch

from . import _native as nv


def _require_cuda():
    if not torch.cuda.is_available():
        raise nv.NativeError('no CUDA device: the beta-cores B200 path has no CPU fallback')


def stream_ptr():
    """raw handle of torch's current stream on the current device.  torch.cuda.current_stream() builds a Stream object
    through several Python layers (16 us per call, measured: it was 0.13 ms of a 0.87 ms optimiser step of the Gaussian
    example); the private accessor below returns the same handle in a fraction of a microsecond."""
    try:
        return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))
    except AttributeError:      # pragma: no cover  (a torch build without the private accessors)
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def dist_group():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


class Engine(object):
    _instances = {}

    @classmethod
    def get(cls, device=None):
        _require_cuda()
        if device is None:
            device = torch.cuda.current_device()
        device = torch.device('cuda', device) if isinstance(device, int) else torch.device(device)
        key = device.index if device.index is not None else torch.cuda.current_device()
        if key not in cls._instances:
            cls._instances[key] = cls(key)
        return cls._instances[key]

    def __init__(self, index):
        self.index = index
        self.device = torch.device('cuda', index)
        self._ctx = {}
        self._stage = []           # pinned staging buffers of _staged_upload
        self.fexp_applied = {}     # ctx name -> feature-exponent tensor last handed to bc_set_feature_exponents
        self.ctx_state = {}        # ctx name -> which potential / whose samples are applied to that workspace (_fused.py)
        self.sms = None

    def ctx(self, name='main'):
        """A bc_ctx workspace; `name` separates workspaces used on concurrent streams."""
        if name not in self._ctx:
            h = ctypes.c_void_p()
            with torch.cuda.device(self.device):
                nv.call('bc_create', self.index, ctypes.byref(h))
            self._ctx[name] = h
            if self.sms is None:
                self.sms = nv.lib().bc_sm_count(h)
        return self._ctx[name]

    def empty(self, *shape, dtype=torch.float64):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    def zeros(self, *shape, dtype=torch.float64):
        return torch.zeros(*shape, dtype=dtype, device=self.device)

    def upload(self, arr, dtype=torch.float64):
        a = np.ascontiguousarray(arr)
        src = torch.from_numpy(a)
        if (a.nbytes >= self.STAGED_UPLOAD_MIN_BYTES and a.ndim == 2 and src.dtype == dtype and not src.is_pinned()):
            return self._staged_upload(src)
        return src.to(self.device, dtype=dtype, non_blocking=False)

    # A large PAGEABLE host matrix (what a numpy user hands to the coreset classes) goes up at ~11 GB/s through the
    # driver's own bounce buffers -- 0.9 s for the 10 GB of the north-star rows -- against 55 GB/s from pinned memory.
    # Row blocks are copied into pinned staging buffers by a few host threads (the copies release the GIL) and sent from
    # there asynchronously on a side stream, so the host copies and the DMA overlap.
    STAGED_UPLOAD_MIN_BYTES = 256 << 20
    STAGED_UPLOAD_BLOCK_BYTES = 64 << 20
    STAGED_UPLOAD_THREADS = 4

    def _staged_upload(self, src):
        from concurrent.futures import ThreadPoolExecutor
        n, row_bytes = src.shape[0], src.shape[1]*src.element_size()
        rows_per = max(1, self.STAGED_UPLOAD_BLOCK_BYTES // row_bytes)
        dst = torch.empty(src.shape, dtype=src.dtype, device=self.device)
        nbuf = self.STAGED_UPLOAD_THREADS
        if len(self._stage) < nbuf:                   # pinned staging buffers are kept: page-locking costs more than the copy
            self._stage += [torch.empty(self.STAGED_UPLOAD_BLOCK_BYTES, dtype=torch.uint8).pin_memory()
                            for _ in range(nbuf - len(self._stage))]
        stage = [self._stage[k][:rows_per*row_bytes].view(src.dtype).view(rows_per, src.shape[1]) for k in range(nbuf)]
        side = torch.cuda.Stream(device=self.device)
        # `dst` may be a recycled block whose previous owner still has kernels queued on the current stream
        side.wait_stream(torch.cuda.current_stream(self.device))
        dst.record_stream(side)
        blocks = [(s, min(s + rows_per, n)) for s in range(0, n, rows_per)]

        def send(k):
            # worker k owns staging buffer k: blocks k, k + nbuf, ...; an event guards the buffer's reuse
            ev = None
            for s, e in blocks[k::nbuf]:
                if ev is not None:
                    ev.synchronize()
                stage[k][:e-s].copy_(src[s:e])
                with torch.cuda.stream(side):
                    dst[s:e].copy_(stage[k][:e-s], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(side)
        with torch.cuda.device(self.device):
            with ThreadPoolExecutor(max_workers=nbuf) as ex:
                list(ex.map(send, range(nbuf)))
            side.synchronize()
        torch.cuda.current_stream(self.device).wait_stream(side)
        return dst


def padded_ld(ncols):
    """Device leading dimension of a data matrix: multiple of 4 doubles (32 bytes) so that every
    row start is 16-byte aligned and the DMMA k-loop never reads past a row's allocation."""
    return ((ncols + 3) // 4) * 4


class DeviceRows(object):
    """A block of data rows resident in HBM for the whole build (zero-padded to `ld` columns).

    In a torch.distributed job every rank holds a contiguous shard [row0, row0 + n_local) of the
    N rows; positions reported by the kernels are global row numbers.
    """

    def __init__(self, engine, host_rows, row0=0, n_total=None):
        host_rows = np.asarray(host_rows, dtype=np.float64)
        if host_rows.ndim != 2:
            host_rows = host_rows.reshape(1, -1) if host_rows.size else host_rows.reshape(0, max(host_rows.shape[-1], 1))
        self.engine = engine
        self.n_local, self.ncols = host_rows.shape
        self.row0 = int(row0)
        self.n_total = int(n_total) if n_total is not None else self.n_local
        self.is_shard = n_total is not None      # built as one rank's slice of a larger set (even if this rank got all of it)
        self.ld = padded_ld(self.ncols)
        if self.ld == self.ncols:
            self.t = engine.upload(host_rows)
        else:
            self.t = engine.zeros(max(self.n_local, 1), self.ld)
            if self.n_local:
                self.t[:self.n_local, :self.ncols].copy_(torch.from_numpy(np.ascontiguousarray(host_rows)))
        self.rowaux = None          # Gaussian x Siginv x cache
        self.rowaux_key = None
        self._q = {}

    @classmethod
    def from_device(cls, engine, t, row0=0, n_total=None):
        """wrap rows that are ALREADY resident in HBM: a contiguous fp64 CUDA tensor (n_local, ld) whose
        leading dimension satisfies padded_ld (e.g. a shard generated or loaded on the device)"""
        if not (t.is_cuda and t.dtype == torch.float64 and t.dim() == 2 and t.is_contiguous()):
            raise TypeError('from_device needs a contiguous 2-d fp64 CUDA tensor')
        if t.shape[1] != padded_ld(t.shape[1]) or t.data_ptr() % 16:
            raise ValueError('device rows need a leading dimension that is a multiple of 4 doubles and a 16-byte aligned base')
        self = cls.__new__(cls)
        self.engine = engine
        self.n_local, self.ncols = int(t.shape[0]), int(t.shape[1])
        self.row0 = int(row0)
        self.n_total = int(n_total) if n_total is not None else self.n_local
        self.is_shard = n_total is not None
        self.ld = self.ncols
        self.t = t
        self.rowaux = None
        self.rowaux_key = None
        self._q = {}
        return self

    def quantised(self, ctx, D, aux_col=None):
        """(image, rowscale, aux, fexp) for the tensor-core route: the rows split once into 7 int8 digit planes, stored
        in the tensor core's own swizzled tile layout (bc_quantise_rows); aux = column `aux_col` of every row; fexp = the
        per-feature power-of-two exponents the image was built with (bc_feature_exponents; the maximum over the row
        shards of a sharded job, so that the image -- and every result -- does not depend on how the rows are split)."""
        key = (int(D), aux_col)
        if key not in self._q:
            from ._native import call, c_i64
            from ._shard import Comm
            nb = c_i64()
            call('bc_q_image_bytes', self.n_local, ctypes.byref(nb))
            img = torch.empty(max(nb.value, 16), dtype=torch.uint8, device=self.engine.device)
            rs = self.engine.empty(max(self.n_local, 1))
            aux = self.engine.empty(max(self.n_local, 1)) if aux_col is not None else None
            fexp = torch.empty(int(D), dtype=torch.int32, device=self.engine.device)
            call('bc_feature_exponents', ctx, ptr(self.t), self.ld, self.n_local, int(D), ptr(fexp), stream_ptr())
            comm = Comm.current()
            if self.sharded and comm.world > 1:
                fexp = comm.allgather(fexp).amax(dim=0).to(torch.int32)
            fexp = torch.where(fexp < -(1 << 20), torch.zeros_like(fexp), fexp).contiguous()
            if self.n_local:
                call('bc_quantise_rows', ctx, ptr(self.t), self.ld, self.n_local, int(D), 0 if aux_col is None else int(aux_col),
                     ptr(img), ptr(rs), ptr(aux), ptr(fexp), stream_ptr())
            self._q[key] = (img, rs, aux, fexp)
        return self._q[key]

    @property
    def sharded(self):
        """rows split over the ranks of the current process group.  A job-level property: every rank must answer alike (it
        guards collectives), so it cannot be read off the local row count -- a rank may own all of a tiny set, or none"""
        from ._shard import Comm
        return self.is_shard and Comm.current().world > 1

    @property
    def shape(self):
        return (self.n_total, self.ncols)

    def __getitem__(self, f):
        """host copy of GLOBAL row f (the coreset classes read `data[f]` for the selected point); in a
        sharded job the owning rank broadcasts it."""
        from ._shard import Comm, owner_of
        f = int(f)
        comm = Comm.current()
        buf = self.engine.empty(self.ld)
        own = owner_of(f, self.n_total, comm.world) if (comm.world > 1 and self.sharded) else comm.rank
        if own == comm.rank:
            buf.copy_(self.t[f-self.row0])
        if comm.world > 1 and self.sharded:
            comm.broadcast(buf, own)
        return buf[:self.ncols].cpu().numpy()
