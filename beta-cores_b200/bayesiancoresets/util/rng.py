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

Only the samplers built by this package's factories with `prefetch=True` switch it on (`activate()`); the coreset classes
route their own sub-sample draws through `randint()` below so that they take part in the ordering.  Code that draws from
np.random on its own while a speculation is in flight must call `drain()` first (np.random.seed users: the factories'
samplers expose it).
"""
import numpy as np
from concurrent.futures import ThreadPoolExecutor

_ACTIVE = None


class _Spec(object):
    __slots__ = ('ops', 'fut')

    def __init__(self, ops, fut):
        self.ops, self.fut = ops, fut


def _do(op):
    kind, args, post = op
    if kind == 'randn':
        r = np.random.randn(*args)
    else:
        r = np.random.randint(args[0], size=args[1])
    return r if post is None else post(r)


def _run(ops):
    """helper thread: the draws of one cycle, in order, with the stream state after each (for rewinding)"""
    state0 = np.random.get_state()
    out, states = [], []
    for op in ops:
        out.append(_do(op))
        states.append(np.random.get_state())
    return state0, out, states


class StreamAhead(object):
    def __init__(self):
        self._pool = ThreadPoolExecutor(max_workers=1)
        self._cycle = []      # draws requested since the last begin_cycle()
        self._cur = None      # speculation covering the current cycle
        self._nxt = None      # speculation queued behind it for the next cycle
        self._k = 0           # draws of _cur handed out
        self.hits = self.rewinds = 0

    # ---- protocol ----
    def begin_cycle(self):
        """the sampler calls this first thing: the draws since the previous call are the pattern speculated from here on"""
        if self._cur is not None and self._k < len(self._cur.ops):
            self._rewind()                     # the cycle ended before the speculated draws were used up
        pattern, self._cycle = self._cycle, []
        self._cur, self._nxt, self._k = self._nxt, None, 0
        if not pattern or len(pattern) > 16:
            if self._cur is not None:
                self._rewind()
            return
        if self._cur is None:
            # nothing in flight: queue this cycle's draws now (the caller waits for them once) ...
            self._cur = _Spec(pattern, self._pool.submit(_run, pattern))
        elif self._cur.ops != pattern:
            self._rewind()
            self._cur = _Spec(pattern, self._pool.submit(_run, pattern))
        # ... and the next cycle's right behind them: from the next call on the helper is a full cycle ahead
        self._nxt = _Spec(pattern, self._pool.submit(_run, pattern))

    def _rewind(self):
        """drop every speculated draw the caller has not consumed: put the global stream back right after the last consumed one"""
        state0, _, states = self._cur.fut.result()
        if self._nxt is not None:
            self._nxt.fut.result()
        np.random.set_state(states[self._k-1] if self._k > 0 else state0)
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
        """np.random.randn(S, D); `post` (optional) runs on the thread that made the draw (e.g. a copy into pinned memory)"""
        return self._request(('randn', (int(S), int(D)), post))

    def randint(self, high, size):
        return self._request(('randint', (int(high), int(size)), None))

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
