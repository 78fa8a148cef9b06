"""hand-off timeline of k_project_q on SM 0 (tracing build of the library only: BC_LIB_PATH=.../lib_trace.so)"""
import os, sys, ctypes, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'beta-cores_b200')]
import numpy as np, torch
from bayesiancoresets import _native as nv
from bayesiancoresets._device import Engine, ptr, stream_ptr
eng = Engine.get(); ctx = eng.ctx('t'); dev = eng.device
N, D, S = 148*128*6, 128, 1024
X = torch.randn(N, D, dtype=torch.float64, device=dev)
T = torch.randn(S, D, dtype=torch.float64, device=dev)/np.sqrt(D)
KIND = nv.KIND_LOGLIK if (len(sys.argv) > 1 and sys.argv[1] == 'loglik') else nv.KIND_BETALIK
nv.call('bc_set_potential', ctx, nv.MODEL_LOGISTIC, KIND, D, nv.params8([0.1, 11.0, 0, 0, 0, 0, 0, 0]), None)
nv.call('bc_set_samples', ctx, ptr(T), S, D, stream_ptr())
nb = ctypes.c_int64(); nv.call('bc_q_image_bytes', N, ctypes.byref(nb))
img = torch.empty(nb.value, dtype=torch.uint8, device=dev); rs = torch.empty(N, dtype=torch.float64, device=dev)
nv.call("bc_quantise_rows", ctx, ptr(X), D, N, D, 0, ptr(img), ptr(rs), None, None, stream_ptr())
o = torch.empty(2*(S+1), dtype=torch.float64, device=dev)
for _ in range(3):
    nv.call('bc_project_colsum_q', ctx, ptr(img), ptr(rs), N, None, ptr(o), stream_ptr())
torch.cuda.synchronize()
buf = (ctypes.c_uint64*(24*512))()
L = ctypes.CDLL(os.environ['BC_LIB_PATH'])
print('rc', L.bc_trace_read(buf))
a = np.frombuffer(buf, dtype=np.uint64).reshape(24, 512).astype(np.int64)
t0 = a[0, 0]
a = a - t0
np.save(os.path.join(ROOT, 'gpurun_out', 'q_trace.npy'), a)
# MMA issuer rows indexed by chunk itb; group rows indexed by the group's own chunk counter (chunk = 2*i + buf)
for c in range(64, 80):
    g = (c & 1)*2   # first group of the pair that owns this chunk
    i = c >> 1
    print('chunk %3d | mma: wait_empty %7d got_empty %7d got_b %7d issued %7d | grp%d: want %7d got %7d release %7d end %7d | grp%d: want %7d got %7d release %7d end %7d'
          % (c, a[0, c], a[1, c], a[2, c], a[3, c], g, a[4+g, i], a[8+g, i], a[12+g, i], a[16+g, i], g+1, a[5+g, i], a[9+g, i], a[13+g, i], a[17+g, i]))

# steady state (chunks 64..160 of CTA 0): mean periods
sel = np.arange(64, 160)
mma_issue = a[3, sel] - a[2, sel]
mma_wait_empty = a[1, sel] - a[0, sel]
print('MMA issuer per chunk: wait for the buffer %.0f cycles, issue+commit %.0f, chunk period %.0f' % (mma_wait_empty.mean(), mma_issue.mean(), np.diff(a[3, sel]).mean()))
for g in range(4):
    i = np.arange(32, 80)
    print('group %d per chunk: wait tmem_full %.0f, full->release %.0f, release->end %.0f, period %.0f' % (
        g, (a[8+g, i]-a[4+g, i]).mean(), (a[12+g, i]-a[8+g, i]).mean(), (a[16+g, i]-a[12+g, i]).mean(), np.diff(a[16+g, i]).mean()))
# completion time of the MMAs of chunk c (= the moment the owning groups get tmem_full) against their issue
for c in range(64, 72):
    g = (c & 1)*2
    print('chunk %d: issued at %d (buffer free at %d), accumulators complete by %d -> %d cycles after issue' % (c, a[3, c], a[1, c], min(a[8+g, c >> 1], a[9+g, c >> 1]), min(a[8+g, c >> 1], a[9+g, c >> 1]) - a[2, c]))
