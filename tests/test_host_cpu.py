"""CPU-only: the C-ABI library loads and exports every symbol include/betacores.h declares; host-side logic
(sharding helpers, arg-max merge semantics, potential constants); world_size-2 gloo exchange of the per-rank
parts.  No compute call is made (no GPU here): the product has no CPU path and says so."""
import ctypes
import os
import re
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'betacores.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(bc_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    path = ge.build()
    L = ctypes.CDLL(path)
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), 'libbetacores.so does not export %s' % n
    L.bc_version.restype = ctypes.c_int
    assert L.bc_version() >= 100
    L.bc_error_string.restype = ctypes.c_char_p
    assert b'argument' in L.bc_error_string(-1)


def test_ctypes_table_covers_header():
    from bayesiancoresets import _native as nv
    bound = set(nv.SIGNATURES) | set(nv.PLAIN)
    assert bound == set(_declared_symbols())


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    import bayesiancoresets as bc
    from bayesiancoresets import _native as nv
    import model_lr
    with pytest.raises(nv.NativeError):
        model_lr.beta_likelihood(np.zeros((3, 2)), np.zeros((4, 2)), 0.1)
    with pytest.raises(nv.NativeError):
        bc.snnls.GIGA(np.ones((3, 5)), np.ones(3))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'beta-cores_b200')
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                s = open(os.path.join(d, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', s, flags=re.M), os.path.join(d, f)


def test_partition_and_owner():
    from bayesiancoresets._shard import partition_rows, owner_of, local_subsample
    for n, w in ((10, 3), (7, 8), (1000003, 8), (16, 4), (5, 1)):
        spans = [partition_rows(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and sum(s[1] for s in spans) == n
        for r in range(1, w):
            assert spans[r][0] == spans[r-1][0]+spans[r-1][1]
        for row in (0, n-1, n//2, n//3):
            own = owner_of(row, n, w)
            assert spans[own][0] <= row < spans[own][0]+spans[own][1]
    sub = np.array([5, 0, 9, 5, 3, 7])
    pos, loc = local_subsample(sub, 4, 4)
    assert list(pos) == [0, 3, 5] and list(loc) == [1, 1, 3]


def test_merge_best_is_numpy_argmax():
    from bayesiancoresets._shard import merge_best, nan_max
    r = np.random.RandomState(0)
    for trial in range(200):
        n = r.randint(1, 30)
        x = np.round(r.randn(n), 1)
        if trial % 3 == 0:
            x[r.randint(n, size=2)] = np.nan
        cuts = sorted(r.randint(0, n+1, size=3))
        chunks = [(0, cuts[0]), (cuts[0], cuts[1]), (cuts[1], cuts[2]), (cuts[2], n)]
        cands = []
        for a, b in chunks:
            if b > a:
                j = int(np.argmax(x[a:b]))
                cands.append((x[a+j], a+j))
            else:
                cands.append((0.0, -1))
        r.shuffle(cands)
        v, i = merge_best(cands)
        assert i == int(np.argmax(x))
        m = nan_max([c[0] for c in cands if c[1] >= 0])
        assert (np.isnan(m) and np.isnan(x.max())) or m == x.max()


def test_potential_constants_follow_reference_expressions():
    from bayesiancoresets.potentials import DevicePotential
    p = DevicePotential('logistic', 'betalik').params(5, 0.3)
    assert p[0] == 0.3 and p[1] == (0.3+1.)/0.3
    g = DevicePotential('gaussian', 'betalik').bind(Siginv=np.eye(3), logdetSig=0.7)
    q = g.params(3, 0.1)
    assert q[1] == 1./0.1 and q[2] == -.5*0.1 and q[3] == (1+0.1)**(-.5*3.-1)
    with pytest.raises(TypeError):
        DevicePotential('gaussian', 'betalik').params(3, None)
    assert not DevicePotential('neurlin', 'loglik').is_bound()
    assert DevicePotential('neurlin', 'loglik').bind(sigsq=2.).is_bound()


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        sys.path.insert(0, os.path.join(ROOT, 'beta-cores_b200'))
        from bayesiancoresets._shard import Comm, partition_rows, local_subsample, merge_best
        comm = Comm.current()
        assert comm.world == world and comm.rank == rank
        # the data path: every rank reduces its own row block; only S-length parts + one pair are exchanged
        N, S = 1001, 16
        V = np.random.RandomState(3).randn(N, S)
        V[417] = np.nan if False else V[417]
        resid = np.random.RandomState(4).randn(S)
        r0, nl = partition_rows(N, world, rank)
        part = torch.from_numpy(V[r0:r0+nl].sum(axis=0))
        allp = comm.allgather(part).numpy()
        colsum = allp.sum(axis=0)
        sc = V[r0:r0+nl].dot(resid)
        j = int(np.argmax(sc))
        mine = torch.tensor([sc[j], float(r0+j)], dtype=torch.float64)
        allc = comm.allgather(mine).numpy()
        best = merge_best((allc[r, 0], int(allc[r, 1])) for r in range(world))
        # subsample ownership
        sub = np.random.RandomState(5).randint(N, size=200)
        pos, loc = local_subsample(sub, r0, nl)
        cnt = comm.allgather(torch.tensor([len(pos)])).numpy().sum()
        t = torch.zeros(3, dtype=torch.float64)
        if rank == 1:
            t += 7.
        comm.broadcast(t, 1)
        # "sharded" guards collectives, so every rank must answer alike -- also when one rank owns all rows of a tiny set
        # and the other none (no device needed for the property itself)
        from bayesiancoresets._device import DeviceRows
        flags = []
        for n_tot in (1, 3, 1001):
            r0t, nlt = partition_rows(n_tot, world, rank)
            rows = DeviceRows.__new__(DeviceRows)
            rows.n_local, rows.n_total, rows.row0, rows.is_shard = nlt, n_tot, r0t, True
            flags.append(bool(rows.sharded))
        whole = DeviceRows.__new__(DeviceRows)
        whole.n_local, whole.n_total, whole.row0, whole.is_shard = 5, 5, 0, False
        flags.append(bool(whole.sharded))
        assert flags == [True, True, True, False], flags
        q.put((rank, colsum, best, int(cnt), t.numpy()))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_exchange_matches_single_process():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    N, S = 1001, 16
    V = np.random.RandomState(3).randn(N, S)
    resid = np.random.RandomState(4).randn(S)
    for rank, colsum, best, cnt, t in res:
        np.testing.assert_allclose(colsum, V.sum(axis=0), rtol=1e-12, atol=1e-12)
        assert best[1] == int(np.argmax(V.dot(resid)))
        assert cnt == 200
        assert (t == 7.).all()


def test_newton_laplace_mode_is_the_bfgs_mode():
    """the bench's host sampler finds the Laplace mode by Newton steps; it is the stationary point scipy's BFGS (the
    reference's optimiser, util/opt.py:10-33) converges towards"""
    import model_lr
    r = np.random.RandomState(3)
    Z = r.randn(6, 12)
    w = np.abs(r.randn(6))*50.
    mu_b, L_b, _ = model_lr.get_laplace(w, Z, np.zeros(12))
    mu_n, L_n, _ = model_lr.get_laplace(w, Z, np.zeros(12), method='newton')
    assert np.abs(model_lr.grad_th_log_joint(Z, mu_n, w)).max() < 1e-10
    np.testing.assert_allclose(mu_n, mu_b, atol=1e-4)
    np.testing.assert_allclose(L_n, L_b, atol=1e-4)


def test_c_abi_argument_errors_need_no_device():
    """every entry point checks its handle / pointers before touching CUDA: misuse comes back as a code (no device here)"""
    import ctypes
    from bayesiancoresets import _native as nv
    L = nv.lib()
    null = None
    assert L.bc_set_samples(null, null, 4, 4, null) == -1
    assert L.bc_project_colsum(null, null, 0, null, 0, null, null, null) == -1
    assert L.bc_project_colsum_q(null, null, null, 0, null, null, null) == -1
    assert L.bc_quantise_rows(null, null, 0, 0, 1, 0, null, null, null, null, null) == -1
    assert L.bc_feature_exponents(null, null, 0, 0, 1, null, null) == -1
    assert L.bc_set_feature_exponents(null, null, 0, null) == -1
    assert L.bc_laplace_logistic(null, null, 0, null, 1, 1, null, null, 1, 1e-9, null, null) == -1
    assert L.bc_sample_affine(null, null, null, null, 1, 1, null, 1, null) == -1
    assert L.bc_destroy(null) == 0
    nb = ctypes.c_int64()
    assert L.bc_q_image_bytes(1000, ctypes.byref(nb)) == 0 and nb.value == 8*128*7*128   # 8 tiles of 128 rows x 7 digit planes x 128 B
    assert L.bc_q_max_features() == 128
    L.bc_error_string.restype = ctypes.c_char_p
    msgs = {L.bc_error_string(c) for c in (0, -1, -2, -3, -4, -5)}
    assert len(msgs) == 6 and all(m for m in msgs)


def test_uniform_sampling_coreset_matches_reference_outputs():
    """bayesiancoresets/coreset/sampling.py:5-47 (host RNG bookkeeping only, no GPU): the constructor takes `groups` /
    `selected_groups`, a warm start through wts/idcs/pts seeds one count per initial point, group mode takes whole groups.
    Expected values: the unmodified reference run on the same seeds in the build container."""
    import bayesiancoresets as bc
    X = np.arange(60.).reshape(20, 3)
    groups = [list(range(i, i+4)) for i in range(0, 20, 4)]
    np.random.seed(5)
    a = bc.UniformSamplingCoreset(X, wts=np.ones(2), idcs=np.array([3, 7]), pts=X[[3, 7]].copy())
    a.build(6, 100)
    assert a.wts.tolist() == [5.0, 2.5, 2.5, 2.5, 2.5, 2.5, 2.5] and a.idcs.tolist() == [3, 7, 14, 15, 6, 16, 9]
    np.testing.assert_array_equal(a.pts, X[a.idcs])
    np.random.seed(6)
    b = bc.UniformSamplingCoreset(X, wts=np.ones(2), idcs=np.array([3, 7]), pts=X[[3, 7]].copy(), groups=groups, selected_groups=None)
    b.build(3, 100)
    assert b.idcs.tolist() == [3, 7, 8, 9, 10, 11, 4, 5, 6, 7, 12, 13, 14, 15] and b.selected_groups == [2, 1, 3]
    np.testing.assert_allclose(b.wts, np.full(14, 20./14.), rtol=1e-15)
    np.testing.assert_array_equal(b.pts, X[b.idcs])
    np.random.seed(7)
    c = bc.UniformSamplingCoreset(X)
    c.build(5, 100)
    assert c.wts.tolist() == [4.0]*5 and c.idcs.tolist() == [15, 4, 3, 19, 7]
    c.reset()
    assert c.size() == 0 and c.cts == [] and c.ct_idcs == []


def test_fused_projection_reapplies_its_potential_after_another_owner(monkeypatch):
    """the bc_ctx workspace is shared by all FusedProjections of an engine.  A.configure -> B.configure ->
    A.configure must issue bc_set_potential three times, and A must not keep using samples B has replaced."""
    from bayesiancoresets import _fused, _native as nv
    from bayesiancoresets.potentials import DevicePotential
    calls = []
    monkeypatch.setattr(nv, 'call', lambda name, *a: calls.append(name))
    monkeypatch.setattr(nv, 'lib', lambda: type('L', (), {'bc_colsum_ld': staticmethod(lambda S: S+1)})())
    monkeypatch.setattr(_fused, 'stream_ptr', lambda: None)

    class Eng(object):
        ctx_state = {}
        fexp_applied = {}

        def ctx(self, name='main'):
            return 1

        def upload(self, a, dtype=None):
            import torch
            return torch.from_numpy(np.ascontiguousarray(a))
    eng = Eng()
    A = _fused.FusedProjection(eng, DevicePotential('logistic', 'loglik'), 4)
    B = _fused.FusedProjection(eng, DevicePotential('logistic', 'betalik'), 4)
    th = np.zeros((3, 4))
    A.configure(None); A.set_samples(th)
    B.configure(0.1); B.set_samples(th)
    assert A.S is not None
    A.configure(None)
    assert calls.count('bc_set_potential') == 3
    assert A.S is None                       # its prepared samples are gone: set_samples() must follow
    A.set_samples(th)
    A.configure(None)                        # nothing changed hands: no further call
    assert calls.count('bc_set_potential') == 3 and A.S == 3
    with pytest.raises(nv.NativeError):
        B.set_samples(th)                    # B never re-configured: the workspace holds A's potential


def test_native_legacy_stream_is_bit_identical_to_numpy():
    """csrc/bc_hostrng.cpp (bc_mt_randn / bc_mt_randint): numpy's legacy RandomState continued natively from its own state --
    same normals bit for bit (odd counts, the cached second value, several worker threads), same bounded integers, and
    the state handed back is the state numpy itself ends in (numpy/random/src/legacy/legacy-distributions.c::legacy_gauss,
    mt19937.c, _bounded_integers.pyx masked rejection)."""
    import ctypes
    from bayesiancoresets import _native as nv
    from bayesiancoresets.util import rng
    np.random.seed(20260101)
    np.random.randn(3)                      # leaves a cached gaussian behind
    for n, threads, high, k in [(1, 1, 5700, 13), (7, 1, 1, 4), (20000, 4, 10_000_000, 200), (3, 2, 2**32, 9), (131073, 3, 1000, 50),
                                (4096, 2, 3, 7), (20001, 4, 5700, 1000), (2, 1, 2**31, 3)]:
        st = np.random.get_state()
        ref = np.random.randn(n)
        ri = np.random.randint(high, size=k)
        after = np.random.get_state()
        np.random.set_state(st)
        m = rng._checkout()
        out = np.empty(n)
        nv.call('bc_mt_randn', ctypes.byref(m), out.ctypes.data, n, threads)
        oi = np.empty(k, dtype=np.int64)
        nv.call('bc_mt_randint', ctypes.byref(m), high, oi.ctypes.data, k)
        rng._checkin(m)
        np.testing.assert_array_equal(out.view(np.int64), ref.view(np.int64))
        np.testing.assert_array_equal(oi, ri)
        s2 = np.random.get_state()
        np.testing.assert_array_equal(s2[1], after[1])
        assert tuple(s2[2:]) == tuple(after[2:])


def test_stream_ahead_peek_consumes_nothing():
    """StreamAhead.peek_next(): the samplers look at the NEXT cycle's speculated draw to start its upload a step early
    (examples/common/model_lr.py::device_step).  Peeking hands out the very object the next cycle's request returns when the
    speculation holds, consumes nothing, and changes no number when the pattern changes and the speculation is rewound."""
    from bayesiancoresets.util.rng import StreamAhead
    S, D, N = 6, 4, 500
    script = [[('randn', S, D)]]*5 + [[('randn', S, D), ('randint', N, 9)]]*4 + [[('randn', S+2, D)]]*3

    def play(ahead, peek):
        out, same = [], 0
        pending = None
        for cycle in script:
            if ahead is not None:
                ahead.begin_cycle()
            for j, op in enumerate(cycle):
                if op[0] == 'randn':
                    r = ahead.randn(op[1], op[2]) if ahead is not None else np.random.randn(op[1], op[2])
                    if j == 0 and pending is not None and pending is r:
                        same += 1
                else:
                    r = ahead.randint(op[1], op[2]) if ahead is not None else np.random.randint(op[1], size=op[2])
                out.append(r)
            pending = ahead.peek_next() if (peek and ahead is not None) else None
        if ahead is not None:
            ahead.drain()
        out.append(np.random.rand(2))
        return out, same
    np.random.seed(77)
    ref, _ = play(None, False)
    np.random.seed(77)
    got, same = play(StreamAhead(), True)
    for x, y in zip(ref, got):
        np.testing.assert_array_equal(x, y)
    assert same >= 6          # in the steady stretches the peeked object IS what the next cycle handed out


def test_stream_ahead_consumes_the_global_stream_like_direct_draws():
    """util/rng.py: draws made one sampler call ahead on a helper thread, with pattern changes (selection vs optimiser
    sub-sample sizes, cycles that end early or run long) and a drain in between, return exactly the numbers direct
    np.random calls return -- and leave the global stream in the same state."""
    from bayesiancoresets.util.rng import StreamAhead
    S, D, N = 7, 5, 1000
    # a script of cycles: each = the draws between two sampler calls
    script = ([[('randn', S, D)]]*3 + [[('randn', S, D), ('randint', N, 20)]]*5 + [[('randn', S, D), ('randint', N, 50)]]
              + [[('randn', S, D), ('randint', N, 20)]]*4 + [[('randn', S, D)]] + [[('randn', S, D), ('randint', N, 20), ('randint', N, 3)]]*3
              + [[('randn', S+1, D)]]*2)

    def play(ahead):
        out = []
        for c, cycle in enumerate(script):
            if ahead is not None:
                ahead.begin_cycle()
            for op in cycle:
                if op[0] == 'randn':
                    out.append(ahead.randn(op[1], op[2]) if ahead is not None else np.random.randn(op[1], op[2]))
                else:
                    out.append(ahead.randint(op[1], op[2]) if ahead is not None else np.random.randint(op[1], size=op[2]))
            if c == 12 and ahead is not None:
                ahead.drain()
                out.append(np.random.rand(3))        # someone else draws directly after a drain
            elif c == 12:
                out.append(np.random.rand(3))
        if ahead is not None:
            ahead.drain()
        out.append(np.random.rand(4))                # the stream ends up in the same place
        return out
    np.random.seed(123)
    ref = play(None)
    np.random.seed(123)
    a = StreamAhead()
    got = play(a)
    assert len(ref) == len(got)
    for x, y in zip(ref, got):
        np.testing.assert_array_equal(x, y)
    assert a.hits > 20 and a.rewinds >= 4          # it did speculate, and it did have to rewind at the pattern changes
