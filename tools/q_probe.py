"""GPU probe of the tensor-core (tcgen05 int8) projection route: contraction accuracy, column sums / arg-max against
the FP64 DMMA route, and a timing at N rows.   python tools/q_probe.py [N]"""
import ctypes, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'beta-cores_b200'), os.path.join(ROOT, 'beta-cores_b200', 'examples', 'common')]
import torch
from bayesiancoresets import _native as nv
from bayesiancoresets._device import Engine, ptr, stream_ptr

eng = Engine.get()
ctx = eng.ctx('probe')
dev = eng.device


def quantise(X, D, aux_col=None):
    n = X.shape[0]
    nb = ctypes.c_int64()
    nv.call('bc_q_image_bytes', n, ctypes.byref(nb))
    img = torch.empty(nb.value, dtype=torch.uint8, device=dev)
    rs = torch.empty(max(n, 1), dtype=torch.float64, device=dev)
    aux = torch.empty(max(n, 1), dtype=torch.float64, device=dev) if aux_col is not None else None
    nv.call('bc_quantise_rows', ctx, ptr(X), int(X.stride(0)), n, D, 0 if aux_col is None else aux_col, ptr(img), ptr(rs), ptr(aux), None, stream_ptr())
    return img, rs, aux


def case(n, D, S, seed, scale=1.0):
    rng = np.random.RandomState(seed)
    ld = ((D + 3)//4)*4
    Xh = np.zeros((n, ld)); Xh[:, :D] = rng.randn(n, D)*scale*np.exp(rng.randn(n, 1))
    Th = rng.randn(S, D)*np.exp(0.5*rng.randn(S, 1))
    X = torch.from_numpy(Xh).to(dev); T = torch.from_numpy(Th).to(dev)
    nv.call('bc_set_potential', ctx, nv.MODEL_LOGISTIC, nv.KIND_BETALIK, D, nv.params8([0.1, 11.0, 0, 0, 0, 0, 0, 0]), None)
    nv.call('bc_set_samples', ctx, ptr(T), S, int(T.stride(0)), stream_ptr())
    img, rs, _ = quantise(X, D)
    V = torch.empty(n, S, dtype=torch.float64, device=dev)
    nv.call('bc_contraction_q', ctx, ptr(img), ptr(rs), n, ptr(V), S, stream_ptr())
    torch.cuda.synchronize()
    ref = (Xh[:, :D].astype(np.longdouble) @ Th.T.astype(np.longdouble))
    got = V.cpu().numpy()
    bound = np.abs(Xh).max(axis=1)[:, None]*np.abs(Th).max(axis=1)[None, :]
    err = np.abs(got - ref).astype(np.float64)/bound
    f64 = np.abs(Xh[:, :D] @ Th.T - ref).astype(np.float64)/bound
    print('contraction n=%d D=%d S=%d: max err/(|x|max|th|max) = %.3e   (numpy dgemm: %.3e)' % (n, D, S, err.max(), f64.max()), flush=True)
    # column sums + arg-max vs the DMMA route
    Sld = nv.lib().bc_colsum_ld(S)
    o1 = torch.empty(2*Sld, dtype=torch.float64, device=dev); o2 = torch.empty_like(o1)
    nv.call('bc_project_colsum', ctx, ptr(X), ld, None, n, None, ptr(o1), stream_ptr())
    nv.call('bc_project_colsum_q', ctx, ptr(img), ptr(rs), n, None, ptr(o2), stream_ptr())
    c1 = torch.empty(S, dtype=torch.float64, device=dev); c2 = torch.empty_like(c1)
    nv.call('bc_colsum_combine', ctx, ptr(o1), 1, S, ptr(c1), stream_ptr())
    nv.call('bc_colsum_combine', ctx, ptr(o2), 1, S, ptr(c2), stream_ptr())
    torch.cuda.synchronize()
    a, b = c1.cpu().numpy(), c2.cpu().numpy()
    print('   colsum: max |dmma - q| / max|dmma| = %.3e' % (np.abs(a-b).max()/np.abs(a).max()), flush=True)
    resid = torch.cat([c1, c1.sum()[None]])
    b1 = torch.zeros(2, dtype=torch.float64, device=dev); b2 = torch.zeros_like(b1)
    s1 = torch.empty(n, dtype=torch.float64, device=dev); s2 = torch.empty_like(s1)
    nv.call('bc_project_score', ctx, ptr(X), ld, None, n, None, ptr(resid), 5, ptr(b1), ptr(s1), stream_ptr())
    nv.call('bc_project_score_q', ctx, ptr(img), ptr(rs), n, None, ptr(resid), 5, ptr(b2), ptr(s2), stream_ptr())
    torch.cuda.synchronize()
    v1, v2 = b1.cpu().numpy(), b2.cpu().numpy()
    sa, sb = s1.cpu().numpy(), s2.cpu().numpy()
    print('   score: best dmma (%.12e, %d)  q (%.12e, %d)  max rel score diff %.3e' % (
        v1[0], v1[1:2].view(np.int64)[0], v2[0], v2[1:2].view(np.int64)[0], np.nanmax(np.abs(sa-sb)/np.abs(sa).max())), flush=True)


for (n, D, S) in [(128, 128, 32), (300, 128, 70), (1000, 20, 100), (5000, 100, 200), (4096, 64, 1024)]:
    case(n, D, S, seed=n)

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
D, S = 128, 1024
X = torch.randn(N, D, dtype=torch.float64, device=dev)
T = torch.randn(S, D, dtype=torch.float64, device=dev)/np.sqrt(D)
nv.call('bc_set_potential', ctx, nv.MODEL_LOGISTIC, nv.KIND_BETALIK, D, nv.params8([0.1, 11.0, 0, 0, 0, 0, 0, 0]), None)
nv.call('bc_set_samples', ctx, ptr(T), S, D, stream_ptr())
t0 = time.perf_counter(); img, rs, _ = quantise(X, D); torch.cuda.synchronize(); print('quantise %d rows: %.2f ms' % (N, 1e3*(time.perf_counter()-t0)))
Sld = nv.lib().bc_colsum_ld(S)
o = torch.empty(2*Sld, dtype=torch.float64, device=dev)
for name, fn in [('dmma', lambda: nv.call('bc_project_colsum', ctx, ptr(X), D, None, N, None, ptr(o), stream_ptr())),
                 ('q   ', lambda: nv.call('bc_project_colsum_q', ctx, ptr(img), ptr(rs), N, None, ptr(o), stream_ptr()))]:
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/5
    print('colsum %s N=%d: %.3f ms/pass  %.2f G evals/s' % (name, N, ms, N*S/ms/1e6), flush=True)
