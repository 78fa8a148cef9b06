"""GPU probe: time of one fused column-sum pass (tensor-core route) for every built-in potential, N rows x S = 1024."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'beta-cores_b200'), os.path.join(ROOT, 'beta-cores_b200', 'examples', 'common')]
import numpy as np, torch
from bayesiancoresets import _fused
from bayesiancoresets._fused import FusedProjection
from bayesiancoresets._device import Engine, DeviceRows
from bayesiancoresets.potentials import DevicePotential
eng = Engine.get(); dev = eng.device
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
S = 1024
r = np.random.RandomState(0)
cases = []
for D in (128, 20):
    cases.append(('logistic', 'betalik', D, {}, 0.1))
cases.append(('logistic', 'loglik', 128, {}, None))
A = r.randn(100, 100); Sig = A.dot(A.T) + 100*np.eye(100)
cases.append(('gaussian', 'betalik', 100, dict(Siginv=np.linalg.inv(Sig), logdetSig=np.linalg.slogdet(Sig)[1]), 0.1))
cases.append(('gaussian', 'loglik', 100, dict(Siginv=np.linalg.inv(Sig), logdetSig=np.linalg.slogdet(Sig)[1]), None))
cases.append(('neurlin', 'betalik', 64, dict(sigsq=0.8), 0.2))
cases.append(('neurlin', 'loglik', 64, dict(sigsq=0.8), None))
for model, kind, D, consts, beta in cases:
    ncols = D + (1 if model == 'neurlin' else 0)
    ld = ((ncols + 3)//4)*4
    X = torch.zeros(N, ld, dtype=torch.float64, device=dev)
    X[:, :ncols] = torch.randn(N, ncols, dtype=torch.float64, device=dev)*(0.3 if model == 'gaussian' else 1.0)
    Th = torch.randn(S, D, dtype=torch.float64, device=dev)/np.sqrt(D)
    pot = DevicePotential(model, kind).bind(**consts)
    fp = FusedProjection(eng, pot, ld if model != 'neurlin' else ncols)
    rows = DeviceRows.from_device(eng, X) if ld == ncols else DeviceRows(eng, X[:, :ncols].cpu().numpy())
    fp = FusedProjection(eng, pot, rows.ncols)
    fp.configure(beta)
    fp.set_samples(Th)
    out = {}
    for route in ('q', 'dmma'):
        _fused.ROUTE = route
        for _ in range(2): fp.colsum_parts(rows)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): fp.colsum_parts(rows)
        e1.record(); torch.cuda.synchronize()
        out[route] = e0.elapsed_time(e1)/5
    print('%-9s %-8s D=%-4d N=%d S=%d: tensor-core route %.3f ms (%.1f G evals/s), DMMA route %.3f ms' % (
        model, kind, D, N, S, out['q'], N*S/out['q']/1e6, out['dmma']), flush=True)
    del X, rows, fp
    torch.cuda.empty_cache()
