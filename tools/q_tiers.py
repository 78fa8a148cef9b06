"""k_project_q at N rows x 1024 samples x D=128 (logistic beta-likelihood): per digit tier (5/6/7) the column-sum and score pass
times (CUDA events, mean of 5 after 2 warm-ups) and the deviation of the column sums from the FP64 DMMA route.
    python tools/q_tiers.py [N]        (library chosen by BC_LIB_PATH)"""
import os, sys, ctypes, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'beta-cores_b200')]
import numpy as np, torch
from bayesiancoresets import _native as nv
from bayesiancoresets._device import Engine, ptr, stream_ptr
eng = Engine.get(); ctx = eng.ctx('t'); dev = eng.device
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
D, S = 128, 1024
g = torch.Generator(device=dev); g.manual_seed(3)
X = torch.randn(N, D, generator=g, dtype=torch.float64, device=dev)
T = torch.randn(S, D, generator=g, dtype=torch.float64, device=dev)/np.sqrt(D) + 1./np.sqrt(D)
nv.call('bc_set_potential', ctx, nv.MODEL_LOGISTIC, nv.KIND_BETALIK, D, nv.params8([0.1, 11.0, 0, 0, 0, 0, 0, 0]), None)
nv.call('bc_set_samples', ctx, ptr(T), S, D, stream_ptr())
nb = ctypes.c_int64(); nv.call('bc_q_image_bytes', N, ctypes.byref(nb))
img = torch.empty(nb.value, dtype=torch.uint8, device=dev); rs = torch.empty(N, dtype=torch.float64, device=dev)
nv.call('bc_quantise_rows', ctx, ptr(X), D, N, D, 0, ptr(img), ptr(rs), None, None, stream_ptr())
Sld = S+1
o = torch.empty(2*Sld, dtype=torch.float64, device=dev)
cs = torch.empty(S, dtype=torch.float64, device=dev)
best = torch.zeros(4, dtype=torch.float64, device=dev)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps


nv.call('bc_project_colsum', ctx, ptr(X), D, None, N, None, ptr(o), stream_ptr())
nv.call('bc_colsum_combine', ctx, ptr(o), 1, S, ptr(cs), stream_ptr())
ref = cs.cpu().numpy().copy()
resid = torch.cat([cs, cs.sum()[None]]).contiguous()
nv.call('bc_project_score', ctx, ptr(X), D, None, N, None, ptr(resid), 0, ptr(best), None, stream_ptr())
bref = best.cpu().numpy().copy()
out = {'lib': os.environ.get('BC_LIB_PATH', 'default'), 'rows': N}
for dg in (7, 6, 5):
    nv.call('bc_set_contraction_digits', ctx, dg)
    ms_c = timed(lambda: nv.call('bc_project_colsum_q', ctx, ptr(img), ptr(rs), N, None, ptr(o), stream_ptr()))
    nv.call('bc_colsum_combine', ctx, ptr(o), 1, S, ptr(cs), stream_ptr())
    got = cs.cpu().numpy()
    ms_s = timed(lambda: nv.call('bc_project_score_q', ctx, ptr(img), ptr(rs), N, None, ptr(resid), 0, ptr(best), None, stream_ptr()))
    b = best.cpu().numpy()
    out['digits_%d' % dg] = {'colsum_ms': round(ms_c, 4), 'score_ms': round(ms_s, 4), 'G_evals_per_s': round(N*S/ms_c/1e6, 2),
                             'colsum_vs_dmma_rel': float(np.abs(got-ref).max()/np.abs(ref).max()),
                             'argmax_equal_dmma': int(b[1:2].view(np.int64)[0]) == int(bref[1:2].view(np.int64)[0]),
                             'best_rel_diff': float(abs(b[0]-bref[0])/abs(bref[0]))}
print(json.dumps(out), flush=True)
