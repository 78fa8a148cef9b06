"""numpy's GLOBAL random stream, drawn one sampler call AHEAD of the caller on a helper thread -- in exactly the caller's order.

Why: every optimiser step of the reference calls the user's sampler, which ends in `np.random.randn(S, D)`
(examples/zellner_gaussian/main.py:87-92, zellner_logreg/main.py:139-144), and the sub-sampled modes then draw
`np.random.randint(N, size=n)` (bcores.py:53).  The draws do not depend on the weights, only their ORDER in the one
global stream matters for reproducing the reference's samples; at S x D = 200 x 100 the legacy generator needs 0.3 ms,
a third of a whole optimiser step of the Gaussian example.  `StreamAhead` learns the sequence of draws of one cycle
(= from one sampler call to the next), then has a helper thread execute the next cycle's draws while the caller is busy
with the current one (numpy releases the GIL while generating).  If the caller's requests stop matching the speculated
sequence (the selection draws a different sub-sample size than the optimiser steps, a build ends, ...) the global stream
is rewound to the state right after the last draw the caller actually consumed, and drawing continues directly: the stream
is consumed exactly as without this class.

The helper does not call numpy's generator: while a speculation is in flight the global stream is CHECKED OUT of numpy
(`np.random.get_state()` -> a native bc_mt_state) and continued by libbetacores' own implementation of the legacy generator
(csrc/bc_hostrng.cpp: the same MT19937 words, the same polar Box-Muller arithmetic, the process's libm -- bit-identical
output, tests/test_host_cpu.py), which is several times faster than numpy's per-call path and splits the log/sqrt part
over a few threads: at 25-30 ns per normal numpy's generator alone took longer than all the kernels of an optimiser step
of the Gaussian example.  Rewinding restores a native snapshot; the state goes back into numpy (`np.random.set_state`)
whenever speculation stops (drain(), a pattern change), so code that draws from np.random between builds sees the stream
exactly where the reference's run would have left it.

Only the samplers built by this package's factories with `prefetch=True` switch it on (`activate()`); the coreset classes
route their own sub-sample draws through `randint()` below so that they take part in the ordering.  Code that draws from
np.random on its own while a speculation is in flight must call `drain()` first (np.random.seed users: the factories'
samplers expose it).
"""
import ctypes
import os
import numpy as np
from concurrent.futures import ThreadPoolExecutor

from .. import _native as nv

_ACTIVE = None
# worker threads of the native generator's log/sqrt part (the word stream itself is sequential)
RNG_THREADS = int(os.environ.get('BC_RNG_THREADS', str(max(1, min(4, (os.cpu_count() or 2)//2)))))


def _checkout():
    """numpy's global stream as a native state (numpy's own state is stale until _checkin)"""
    s = np.random.get_state()
    if s[0] != 'MT19937':
        raise RuntimeError('np.random global state is not the legacy MT19937 stream')
    m = nv.MTState()
    key = np.ascontiguousarray(s[1], dtype=np.uint32)
    ctypes.memmove(m.key, key.ctypes.data, 624*4)
    m.pos, m.has_gauss, m.gauss = int(s[2]), int(s[3]), float(s[4])
    return m


def _checkin(m):
    np.random.set_state(('MT19937', np.frombuffer(m.key, dtype=np.uint32).copy(), int(m.pos), int(m.has_gauss), float(m.gauss)))


def _snap(m):
    return nv.MTState.from_buffer_copy(m)


class _Spec(object):
    __slots__ = ('ops', 'fut')

    def __init__(self, ops, fut):
        self.ops, self.fut = ops, fut


def _do(op):
    """a draw from numpy's own generator (nothing is checked out)"""
    kind, args, post = op
    if kind == 'randn':
        r = np.random.randn(*args)
    else:
        r = np.random.randint(args[0], size=args[1])
    return r if post is None else post(r)


def _do_native(m, op):
    """the same draw from the checked-out native state `m`"""
    kind, args, post = op
    if kind == 'randn':
        into = getattr(post, 'into', None)
        if into is not None:
            # the consumer names the destination (e.g. a pinned staging buffer): the normals are generated in place
            token, buf = into(args)
            nv.call('bc_mt_randn', ctypes.byref(m), buf.ctypes.data, int(buf.size), RNG_THREADS)
            return token
        r = np.empty(args, dtype=np.float64)
        nv.call('bc_mt_randn', ctypes.byref(m), r.ctypes.data, int(r.size), RNG_THREADS)
    else:
        r = np.empty(args[1], dtype=np.int64)
        nv.call('bc_mt_randint', ctypes.byref(m), int(args[0]), r.ctypes.data, int(r.size))
    return r if post is None else post(r)


def _native_ok(ops):
    return all(op[0] == 'randn' or (op[0] == 'randint' and 1 <= op[1][0] <= 2**32) for op in ops)


def _run(m, ops):
    """helper thread: the draws of one cycle, in order, with the stream state after each (for rewinding)"""
    state0 = _snap(m)
    out, states = [], []
    for op in ops:
        out.append(_do_native(m, op))
        states.append(_snap(m))
    return state0, out, states


class StreamAhead(object):
    def __init__(self):
        self._pool = ThreadPoolExecutor(max_workers=1)
        self._cycle = []      # draws requested since the last begin_cycle()
        self._cur = None      # speculation covering the current cycle
        self._nxt = None      # speculation queued behind it for the next cycle
        self._k = 0           # draws of _cur handed out
        self._mt = None       # the global stream while it is checked out of numpy (owned by the helper thread's tasks)
        self.hits = self.rewinds = 0

    # ---- protocol ----
    def begin_cycle(self):
        """the sampler calls this first thing: the draws since the previous call are the pattern speculated from here on"""
        if self._cur is not None and self._k < len(self._cur.ops):
            self._rewind()                     # the cycle ended before the speculated draws were used up
        pattern, self._cycle = self._cycle, []
        self._cur, self._nxt, self._k = self._nxt, None, 0
        if not pattern or len(pattern) > 16 or not _native_ok(pattern):
            if self._cur is not None:
                self._rewind()
            elif self._mt is not None:         # (cannot happen: a checked-out stream always has a speculation in flight)
                _checkin(self._mt)
                self._mt = None
            return
        if self._cur is not None and self._cur.ops != pattern:
            self._rewind()
        if self._cur is None:
            # nothing in flight: check the stream out and queue this cycle's draws now (the caller waits for them once) ...
            if self._mt is None:
                self._mt = _checkout()
            self._cur = _Spec(pattern, self._pool.submit(_run, self._mt, pattern))
        # ... and the next cycle's right behind them: from the next call on the helper is a full cycle ahead
        self._nxt = _Spec(pattern, self._pool.submit(_run, self._mt, pattern))

    def _rewind(self):
        """drop every speculated draw the caller has not consumed: put the global stream back right after the last consumed one"""
        state0, _, states = self._cur.fut.result()
        if self._nxt is not None:
            self._nxt.fut.result()
        _checkin(states[self._k-1] if self._k > 0 else state0)      # speculation stops: numpy owns the stream again
        self._mt = None
        self._cur = self._nxt = None
        self._k = 0
        self.rewinds += 1

    def _request(self, op):
        if len(self._cycle) <= 16:
            self._cycle.append(op)
        if self._cur is not None:
            if self._k < len(self._cur.ops) and self._cur.ops[self._k] == op:
                r = self._cur.fut.result()[1][self._k]
                self._k += 1
                self.hits += 1
                return r
            self._rewind()
        return _do(op)

    # ---- draws ----
    def randn(self, S, D, post=None):
        """np.random.randn(S, D); `post` (optional) runs on the thread that made the draw (e.g. a copy into pinned memory);
        if it has an attribute `into(shape) -> (token, writable float64 ndarray)` the native generator writes the draw
        straight into that array and the call returns `token` (what post(draw) would have returned)"""
        return self._request(('randn', (int(S), int(D)), post))

    def randint(self, high, size):
        return self._request(('randint', (int(high), int(size)), None))

    def peek_next(self):
        """what the FIRST draw of the next cycle will return if the speculation holds (the very object the next cycle's
        request hands out), or None without a speculation in flight.  Waits for the helper to have made the draw.  Nothing
        is consumed: a caller may start moving the data (e.g. upload a pinned staging buffer) a cycle early and must be
        ready for the next cycle to hand out something else."""
        if self._nxt is None or not self._nxt.ops:
            return None
        return self._nxt.fut.result()[1][0]

    def drain(self):
        """stop speculating and leave the global stream exactly where the caller's own draws have left it"""
        if self._cur is not None:
            self._rewind()
        self._cycle = []


def activate():
    """switch the process-wide helper on (idempotent); returns it"""
    global _ACTIVE
    if _ACTIVE is None:
        _ACTIVE = StreamAhead()
    return _ACTIVE


def active():
    return _ACTIVE


def drain():
    if _ACTIVE is not None:
        _ACTIVE.drain()


def randint(high, size):
    """np.random.randint(high, size=size) for the coreset classes' sub-sample draws (bcores.py:53,57; sparsevi.py:50,54;
    hilbert.py:14), in stream order with the samplers' look-ahead when that is switched on"""
    if _ACTIVE is not None:
        return _ACTIVE.randint(high, size)
    return np.random.randint(high, size=size)
