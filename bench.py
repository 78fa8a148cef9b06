#!/usr/bin/env python
"""bench.py -- beta-Cores coreset construction on B200: N*S beta-likelihood evaluations per second.

Workload (BASELINE.json north star, SURVEY.md 8d "C5"): synthetic Bayesian logistic regression,
N = 10M rows, D = 128, S = 1024 posterior samples, beta = 0.1, 10 % label-flip outliers, full data
(no sub-sampling), Laplace sampler of the weighted coreset posterior on the host.

One STEP = one `BetaCoreset.build(1, m)` iteration = 1 selection + `opt_itrs` projected-ADAM steps
= (1 + opt_itrs) N x S projections of the data in the reference (bcores.py:27-35, :141-150).  Evaluations are
counted ALGORITHMICALLY, (1 + opt_itrs) * (N + M) * S per step -- the second (recompute) pass of the fused
selection is not counted.

    python bench.py --gpus N --steps K --warmup W          # this repo (N > 1: launched by torch.distributed.run)
    python bench.py --impl reference --steps K --warmup W  # the reference algorithm (numpy, host cores)

legs of the default arm
    value        : data rows resident in HBM before the timed region (DeviceRows.from_device)
    e2e          : the public API from HOST buffers -- the upload of the data rows, every sample upload and every
                   weight read-back inside the timed region
    roofline     : mean CUDA-event duration of the dominant kernel (k_project, column-sum mode) inside the timed region
    cpu_baseline : the numpy oracle (a restatement of the reference pinned bit-for-bit to it) on a bounded sample of
                   the same workload on this box's host cores, plus an index/weight parity check of the CUDA path on
                   that sample (rank 0, N = 1 only)
Rows are sharded over ranks at fixed total N (strong scaling); per step the ranks exchange one 2 x (S+1) double-double
part, per selection one (score, position) pair.  Inputs (10.24 GB) exceed L2 (126 MB): no flush needed.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'beta-cores_b200')
for p in (ROOT, PKG, os.path.join(PKG, 'examples', 'common')):
    if p not in sys.path:
        sys.path.insert(0, p)

SAMPLER = 'newton'   # mode finder of the host Laplace sampler (examples/common/model_lr.py); 'bfgs' = the reference's scipy call
BLK = 65536          # rows per generator block: the dataset does not depend on how many ranks generate it
FP64_DMMA_PEAK_TFLOPS = 36.9   # measured on this pool's B200 (profiles/r01_fp64_peaks.jsonl); MEASURED_PEAKS.json has no fp64 entry


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--n', '--rows', dest='n', type=int, default=10_000_000)
    ap.add_argument('--d', type=int, default=128)
    ap.add_argument('--s', type=int, default=1024)
    ap.add_argument('--beta', type=float, default=0.1)
    ap.add_argument('--opt-itrs', type=int, default=20)
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--cpu-rows', type=int, default=16384, help='rows of the cpu_baseline / parity sample')
    ap.add_argument('--cpu-opt-itrs', type=int, default=5)
    ap.add_argument('--cpu-steps', type=int, default=3, help='build(1, m) steps of the cpu_baseline / parity sample')
    ap.add_argument('--ref-rows', type=int, default=16384, help='rows per step of the --impl reference arm')
    ap.add_argument('--ref-opt-itrs', type=int, default=None, help='optimiser steps per build step of the reference arm (default: --opt-itrs)')
    ap.add_argument('--ref-sampler', default='newton', choices=['newton', 'bfgs'],
                    help="mode finder of the reference arm's Laplace sampler: 'newton' (default: the same algebra as the product arm, faster, "
                         "so the ratio is conservative) or 'bfgs' (the reference's stock scipy call, util/opt.py:10-33)")
    ap.add_argument('--digits', type=int, default=None, choices=[5, 6, 7], help='precision tier of the tensor-core contraction (default: the library default)')
    ap.add_argument('--configs', action='store_true', help='instead of the north-star workload: time BASELINE configs 1-4 (build seconds, CPU port beside)')
    ap.add_argument('--sweep', default=None, metavar='N1,N2,...|default',
                    help="BASELINE config 5, the scale sweep: one value-leg line per N (default list 1M,3M,10M,30M,100M) at this GPU count, "
                         "appended to --sweep-out; points whose row shard + row image do not fit one GPU's HBM are reported as skipped")
    ap.add_argument('--sweep-out', default=None, help='file the sweep lines are appended to (default profiles/r02_sweep_n<gpus>.jsonl)')
    ap.add_argument('--sampler', default='hybrid', choices=['newton', 'hybrid', 'device', 'bfgs'],
                    help="product arm's Laplace sampler: 'hybrid' (default) = mode and Cholesky factor on the host, the S x D x D affine map "
                         "of the normals on the device; 'newton' = all on the host (what the reference arm runs); 'device' = all on the "
                         "device (csrc/bc_sampler.cu); 'bfgs' = the reference's scipy optimiser")
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    return ap.parse_args()


def sched(i):
    return 1./(1.+i)          # examples/zellner_logreg/main.py:118 with i0 = 1


def step_evals(N, M, S, opt_itrs):
    return (1 + opt_itrs)*(N + M)*S


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([d.get('num_threads', 1) for d in threadpool_info() if d.get('user_api') == 'blas'] or [1])
    except Exception:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------- reference arm (CPU) --
_SAMPLERS = []


def new_sampler(D, seed, method=None):
    """the sampler of both arms (the product arm may run its algebra on the device: --sampler device); numpy's global
    stream is re-seeded only once no prefetched draw is in flight"""
    import numpy as np
    import model_lr
    for s in _SAMPLERS:
        s.drain()
    np.random.seed(seed)
    s = model_lr.make_laplace_sampler(D, method=method or SAMPLER, prefetch=True)
    _SAMPLERS.append(s)
    return s


def run_oracle_build(Z, S, beta, opt_itrs, steps, warmup, seed=1, method=None):
    """time `steps` build(1, m) iterations of the numpy restatement of the reference on rows Z"""
    import numpy as np
    import model_lr
    from oracle import np_models as om, np_coresets as oc
    D = Z.shape[1]
    o = oc.GreedyVI(Z, new_sampler(D, seed, method), S, lambda p, t: om.lr_betalik(p, t, beta), opt_itrs=opt_itrs, sched=sched)
    for m in range(1, warmup+1):
        o.build(1, m)
    evals = 0
    t0 = time.perf_counter()
    for m in range(warmup+1, warmup+steps+1):
        M0 = o.wts.shape[0]
        o.build(1, m)
        evals += step_evals(Z.shape[0], M0 + 1, S, opt_itrs)
    dt = time.perf_counter() - t0
    return o, evals, dt


def reference_arm(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import numpy as np
    import model_lr
    ref_itrs = a.ref_opt_itrs if a.ref_opt_itrs is not None else a.opt_itrs
    Z, _, _, _ = model_lr.gen_synthetic_outliers(a.ref_rows, a.d, seed=a.seed)
    _, evals, dt = run_oracle_build(Z, a.s, a.beta, ref_itrs, a.steps, a.warmup, method=a.ref_sampler)
    v = evals/dt
    cores = blas_threads()
    sample = ('numpy restatement of the reference (oracle/, pinned bit-for-bit to /root/reference): %d build(1,m) steps on '
              '%d rows x D=%d x S=%d, opt_itrs=%d, %s Laplace sampler; dgemm on %d BLAS threads, elementwise numpy single-threaded as in '
              'the reference.  The reference cannot hold N=%d (its N x S matrix alone is %.0f GB): evaluations per second on this bounded '
              'sample stand for its rate (SURVEY 8d)' % (a.steps, a.ref_rows, a.d, a.s, ref_itrs, a.ref_sampler, cores, a.n, a.n*a.s*8/1e9))
    cfg = workload_config(a, world=1)
    # what this arm actually ran: a bounded row sample of the workload above (same D, S, beta, opt_itrs per step)
    cfg['ran'] = {'rows': a.ref_rows, 'opt_itrs': ref_itrs, 'sampler': a.ref_sampler, 'extrapolated': True,
                  'note': 'rate measured on %d of the %d rows; a full-size reference build does not fit host memory' % (a.ref_rows, a.n)}
    print(json.dumps({
        'impl': 'reference', 'metric': 'beta_likelihood_evals_per_s', 'value': v, 'unit': 'evals/s', 'n_gpus': a.gpus,
        'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': 1e3*dt/a.steps, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': cfg,
        'cpu_baseline': {'value': v, 'unit': 'evals/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': v, 'unit': 'evals/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'host': {'cpu_count': os.cpu_count()},
    }))


def workload_config(a, world):
    return {'workload': 'logistic regression beta-coreset build, N=%d D=%d S=%d beta=%g, 10%% label flips, full data'
                        % (a.n, a.d, a.s, a.beta),
            'N': a.n, 'D': a.d, 'S': a.s, 'beta': a.beta, 'opt_itrs': a.opt_itrs,
            'step': 'one BetaCoreset.build(1, m): 1 selection + opt_itrs ADAM steps = (1+opt_itrs) N x S projections',
            'sampler': {'hybrid': 'Laplace approximation of the weighted coreset posterior, called every optimiser step.  Selection steps: mode '
                                  '(warm-started damped Newton steps) and D x D Cholesky factor on the host, samples mu + C^-1 R^T formed on the '
                                  'device (k_sample_solve).  Optimiser steps: the sampler\'s device_step protocol -- mode (dual-space Newton), '
                                  'factor and samples all by kernels from the device-resident weights, no host round trip.  The S x D normals '
                                  'come from the global numpy stream, drawn one call ahead on a helper thread, in the reference\'s order.  The '
                                  'reference arm runs the same algebra entirely on the host (same random numbers)',
                        'device': 'Laplace approximation of the weighted coreset posterior on the device (csrc/bc_sampler.cu: warm-started '
                                  'damped Newton mode search, D x D Cholesky factor and inverse, affine map of S x D normals drawn from '
                                  'the global numpy stream one call ahead on a helper thread), called every optimiser step; the reference '
                                  'arm runs the same algebra on the host'}.get(
                            getattr(a, 'sampler', 'newton'),
                            'host Laplace approximation of the weighted coreset posterior (mode by warm-started damped Newton steps, D x D '
                            'Cholesky factor, S x D normal draws from the global numpy stream, drawn one call ahead on a helper thread), '
                            'called every optimiser step; the same callback in both arms'),
            'sharding': 'rows over %d rank(s), fixed total N' % world,
            'l2': 'inputs (%.2f GB of rows) exceed the 126 MB L2; no flush' % (a.n*a.d*8/1e9)}


# ---------------------------------------------------------------------------------- B200 arm --
def gen_rows(torch, row0, n_local, D, seed, device, flip=0.1):
    """rows [row0, row0+n_local) of the synthetic dataset Z = y X, generated block-wise on the device"""
    out = torch.empty(n_local, D, dtype=torch.float64, device=device)
    th = 1./math.sqrt(D)
    b = row0 // BLK
    done = 0
    while done < n_local:
        g = torch.Generator(device=device)
        g.manual_seed(seed*1000003 + b)
        X = torch.randn(BLK, D, generator=g, dtype=torch.float64, device=device)
        p = torch.sigmoid(X.sum(dim=1)*th)
        y = torch.where(torch.rand(BLK, generator=g, dtype=torch.float64, device=device) < p, 1., -1.)
        fl = torch.rand(BLK, generator=g, dtype=torch.float64, device=device) < flip
        y = torch.where(fl, -y, y)
        X.mul_(y[:, None])
        lo = max(row0 + done - b*BLK, 0)
        take = min(BLK - lo, n_local - done)
        out[done:done+take].copy_(X[lo:lo+take])
        done += take
        b += 1
    return out


class ClockSampler(object):
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                       '-lms', '200'], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(',')]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0]))
                mx.append(float(c[1]))
                pw.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            s = sorted(sm)
            out.update(sm_mhz=s[len(s)//2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm), power_w_max=max(pw))
        return out


def b200_arm(a, emit=True):
    import numpy as np
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- the beta-cores B200 path has no CPU fallback (use --impl reference for the CPU arm)')
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group('nccl', device_id=dev)
    import bayesiancoresets as bc
    import model_lr
    from bayesiancoresets import _native as nv, _fused
    from bayesiancoresets._device import Engine, DeviceRows
    from bayesiancoresets._shard import partition_rows

    N, D, S, beta, K, W = a.n, a.d, a.s, a.beta, a.steps, a.warmup
    if a.digits is not None:
        _fused.set_contraction_digits(a.digits)
    digits = _fused.DIGITS
    # the host side of this arm is the sampler's D x D algebra: BLAS thread pools only thrash on it
    blas_default = blas_threads()
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=1, user_api='blas')
    except Exception:
        threadpool_limits = None
    eng = Engine.get()
    row0, n_local = partition_rows(N, world, rank)
    Zdev = gen_rows(torch, row0, n_local, D, a.seed, dev)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler_clock = [0, 0.]

    def timed_sampler(fn):
        """counts the sampler calls and the HOST time spent in them; forwards the device-step protocol (there the host time is
        what it takes to queue the kernels and fetch the pre-drawn normals -- the sampler's algebra runs on the GPU)"""
        def sampler(Sn, w, pts):
            t = time.perf_counter()
            out = fn(Sn, w, pts)
            sampler_clock[0] += 1
            sampler_clock[1] += time.perf_counter() - t
            return out
        if hasattr(fn, 'device_step'):
            def device_step(Sn, w_dev, core):
                t = time.perf_counter()
                out = fn.device_step(Sn, w_dev, core)
                sampler_clock[0] += 1
                sampler_clock[1] += time.perf_counter() - t
                return out
            sampler.device_step = device_step
            sampler.supports_device_step = fn.supports_device_step
        if hasattr(fn, 'drain'):
            sampler.drain = fn.drain
        return sampler

    def make_alg(rows):
        # every rank draws the same stream (seed 1); rank 0's samples are broadcast anyway
        prj = bc.BetaBlackBoxProjector(timed_sampler(new_sampler(D, 1, a.sampler)), S, model_lr.beta_likelihood, model_lr.log_likelihood, None)
        return bc.BetaCoreset(rows, prj, opt_itrs=a.opt_itrs, step_sched=sched, beta=beta, learn_beta=False)

    # ------------------------------------------------------------ value: rows resident in HBM --
    alg = make_alg(DeviceRows.from_device(eng, Zdev, row0=row0, n_total=N))
    for m in range(1, W+1):
        alg.build(1, m)
    clocks = ClockSampler(local) if rank == 0 else None
    _fused.PASS_TIMERS = []
    from bayesiancoresets.coreset import _greedy
    _greedy.STEP_EVENTS = [] if world > 1 else None
    sampler_clock[0], sampler_clock[1] = 0, 0.
    launches0 = nv.lib().bc_launch_count()
    evals = 0
    barrier()
    torch.cuda.profiler.start()       # no-op unless run under `ncu --profile-from-start off`
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for m in range(W+1, W+K+1):
        M0 = alg.wts.shape[0]
        alg.build(1, m)
        evals += step_evals(N, alg.wts.shape[0] if alg.wts.shape[0] > M0 else M0, S, a.opt_itrs)
    e1.record()
    barrier()
    torch.cuda.profiler.stop()
    dt = max_over_ranks(e0.elapsed_time(e1)*1e-3)
    launches = nv.lib().bc_launch_count() - launches0
    sampler_calls, sampler_s = sampler_clock[0], sampler_clock[1]
    timers, _fused.PASS_TIMERS = _fused.PASS_TIMERS, None
    step_events, _greedy.STEP_EVENTS = _greedy.STEP_EVENTS, None
    step_breakdown = None
    if step_events:
        # where an optimiser step of a row-sharded job goes (CUDA events on this rank's stream; mean over the steps, then the
        # mean and the max over ranks): sampler + sample preparation | data pass | its finalize | all-gather of the parts
        # (includes waiting for the slowest rank) | second half (combine, coreset rows, residual / gradient / ADAM) | gap to
        # the next step's first kernel
        acc = {k: 0. for k in ('sampler_and_prepare', 'sampler', 'pass', 'finalize', 'exchange', 'second_half', 'gap')}
        host_dt = sorted(b[2][0] - a_[2][0] for a_, b in zip(step_events[:-1], step_events[1:]))
        import numpy as _np
        hparts = _np.median(_np.array([[h[k+1] - h[k] for k in range(len(h)-1)] for _, _, h in step_events if len(h) == 5]), axis=0)
        for j, (sev, pt, _) in enumerate(step_events):
            pb, pe = (pt[2], pt[3]) if pt is not None else (sev[1], sev[1])
            acc['sampler_and_prepare'] += sev[0].elapsed_time(pb)
            if len(sev) > 4:
                acc['sampler'] = acc.get('sampler', 0.) + sev[0].elapsed_time(sev[4])
            acc['pass'] += pb.elapsed_time(pe)
            acc['finalize'] += pe.elapsed_time(sev[1])
            acc['exchange'] += sev[1].elapsed_time(sev[2])
            acc['second_half'] += sev[2].elapsed_time(sev[3])
            if j + 1 < len(step_events):
                g = sev[3].elapsed_time(step_events[j+1][0][0])
                acc['gap'] += g if g < 5. else 0.          # (a selection step lies between two build steps' loops)
        nst = float(len(step_events))
        mine = torch.tensor([acc[k]/nst for k in sorted(acc)], dtype=torch.float64, device='cuda')
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        allr = torch.stack(allr).cpu().numpy()
        step_breakdown = {'unit': 'ms per optimiser step', 'steps': int(nst),
                          'host_ms_between_steps': {'min': 1e3*host_dt[0], 'median': 1e3*host_dt[len(host_dt)//2],
                                                    'median_ms_in': dict(zip(('sampler.device_step', 'first_half_call', 'allgather', 'second_half_call'), [1e3*float(v) for v in hparts])),
                                                    'note': 'wall clock of the host loop on rank 0: a median near the device step time means '
                                                            'the host is paced by the device (a full launch queue), a small minimum that it queues a step in that time'},
                          'mean_over_ranks': dict(zip(sorted(acc), [float(v) for v in allr.mean(axis=0)])),
                          'max_over_ranks': dict(zip(sorted(acc), [float(v) for v in allr.max(axis=0)]))}
    clk = clocks.stop() if clocks else None
    value = evals/dt
    col_ms = [x[2].elapsed_time(x[3]) for x in timers if x[0] == 'colsum']
    sco_ms = [x[2].elapsed_time(x[3]) for x in timers if x[0] == 'score']
    col_mean = sum(col_ms)/len(col_ms)
    col_mean = max_over_ranks(col_mean)
    sco_mean = max_over_ranks(sum(sco_ms)/len(sco_ms)) if sco_ms else None
    flops = 2.*n_local*S*D
    achieved = flops/(col_mean*1e-3)/1e12
    hbm_peak = 6546.2
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
    except Exception:
        pass
    alg_bytes = 8.*(n_local*D + S*D + S)
    route = _fused.ROUTE if D <= nv.lib().bc_q_max_features() else 'dmma'
    bf16_peak = 1644.4
    try:
        bf16_peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['bf16_tflops']
    except Exception:
        pass
    sm_clock_hz = 1e6*(clk['sm_mhz'] if clk and clk.get('sm_mhz') else 1965.0)
    if route == 'q':
        kname = ('k_project_q<LogisticF<BETALIK>, COLSUM> (tcgen05 int8 Ozaki contraction in TMEM + beta-likelihood + centring + '
                 'column sums, fused)')
        digit_pairs, fp64_inst = digits*(digits+1)//2, 38   # kept digit pairs (diagonals d < digits); FP64-pipe instructions per evaluation (SASS count of the lane-table form, profiles/r02_sass_histogram.txt)
        int8_ops = digit_pairs*2.*n_local*S*128
        # measured on this pool (tools/mma_i8_rate.cu, profiles/r01_mma_i8_rate.txt): 8192 int8 MACs per cycle per SM from N = 128 up
        int8_peak = eng.sms*8192*2.*1965e6/1e12
        extra = {
            'route': 'q', 'contraction_digits': digits,
            'precision': ('operands contracted to their leading %d int8 digits (%d bits below the row / sample maximum; 7 digits = the '
                          'accuracy of an fp64 dgemm, the north star allows 3xTF32 = about 30 bits); every tier keeps the reference\'s '
                          'index sequences on all golden builds (tests/test_gpu_parity.py::test_precision_tiers_keep_reference_selections)'
                          % (digits, 8*digits-2)),
            'int8_tensor': {'ops_per_launch': int8_ops, 'achieved_tops': int8_ops/(col_mean*1e-3)/1e12, 'peak_tops': int8_peak,
                            'frac': int8_ops/(col_mean*1e-3)/1e12/int8_peak,
                            'peak_source': 'tcgen05.mma.kind::i8 rate measured on this pool: 8192 MACs/cycle/SM x SMs x 1965 MHz '
                                           '(profiles/r01_mma_i8_rate.txt; MEASURED_PEAKS.json has no int8 figure, 2 x its bf16_tflops '
                                           '= %.0f would be the estimate)' % (2*bf16_peak)},
            'fp64_pipe': {'instructions_per_eval': fp64_inst,
                          'frac_of_issue_peak': n_local*S*fp64_inst/32./(eng.sms*4*0.5*sm_clock_hz*col_mean*1e-3),
                          'note': 'the potential (two table exponentials, one reciprocal, one table power per evaluation: 38 FP64 instructions in the loop, '
                                  '69 in its polynomial form) runs on the FP64 pipe (0.5 warp-instructions/cycle/sub-partition, measured '
                                  'tools/fp64_ipc.cu)'}}
        # FP64 instructions and tcgen05 MMAs issue through ONE pipe of the SM (ncu: sm__pipe_shared_cycles_active = fp64 + tensor, to
        # the decimal, profiles/r02_ncu_k_project_q_lane_tables.txt): the honest denominator of this kernel is that pipe's time
        warp_evals_per_subpartition = n_local*S/32./(eng.sms*4)
        cycles_per_warp_eval = col_mean*1e-3*sm_clock_hz/warp_evals_per_subpartition
        mma_cycles_per_warp_eval = digit_pairs*128*32*128/8192./32.     # one 128 x 32 chunk = 32 warp-evaluations per sub-partition
        extra['shared_pipe'] = {'cycles_per_warp_evaluation': cycles_per_warp_eval,
                                'fp64_cycles': 2.*fp64_inst, 'int8_mma_cycles': mma_cycles_per_warp_eval,
                                'frac': (2.*fp64_inst + mma_cycles_per_warp_eval)/cycles_per_warp_eval,
                                'note': 'pipe-time roofline computed live from launch_ms: (FP64 instructions x 2 cycles + int8 MMA cycles at '
                                        '8192 MACs/cycle/SM) per warp-evaluation per sub-partition over the measured cycles; ncu measured '
                                        '58.7 % for the same quantity (fp64 38.9 % + tensor 19.8 %)'}
    else:
        kname = 'k_project<LogisticF<BETALIK>, MODE_COLSUM> (FP64 DMMA contraction + beta-likelihood + centring + column sums, fused)'
        extra = {'route': 'dmma'}
    roofline = {
        'kernel': kname,
        'bound': 'tensor', 'achieved': achieved, 'peak': FP64_DMMA_PEAK_TFLOPS, 'unit': 'TFLOP/s', 'frac': achieved/FP64_DMMA_PEAK_TFLOPS,
        'traffic': None,
        'peak_source': 'algorithmic fp64 flops 2*N*S*D against the FP64 tensor (DMMA mma.sync m16n8k4 f64) peak: 36.9 TFLOP/s MEASURED on '
                       'this pool (tools/fp64_peaks.cu, profiles/r01_fp64_peaks.jsonl), 40 TFLOP/s nominal -- the fastest NATIVE fp64 '
                       'contraction rate of the chip; MEASURED_PEAKS.json holds no fp64 figure and tcgen05 has no f64 kind (the q route '
                       'contracts on int8 tensor cores, so it can exceed this figure: see int8_tensor and fp64_pipe for the pipes it runs on)',
        'peak_nominal': 40.0, 'frac_of_nominal': achieved/40.0,
        'launch_ms': col_mean, 'launches_timed': len(col_ms), 'rows_per_launch': n_local,
        'algorithmic_flops_per_launch': flops, 'evals_per_s_kernel': n_local*S/(col_mean*1e-3),
        'hbm': {'algorithmic_bytes_per_launch': alg_bytes, 'achieved_gbs': alg_bytes/(col_mean*1e-3)/1e9, 'peak_gbs': hbm_peak,
                'frac': alg_bytes/(col_mean*1e-3)/1e9/hbm_peak},
        'score_pass_ms': sco_mean,
        'share_of_step': (sum(col_ms) + sum(sco_ms))*1e-3/dt,
    }
    roofline.update(extra)
    if route == 'q':
        qbytes = ((n_local + 127)//128)*digits*128*128 + 8.*n_local   # digit planes read + row scales
        roofline['hbm'].update({'algorithmic_bytes_per_launch': qbytes, 'achieved_gbs': qbytes/(col_mean*1e-3)/1e9,
                                'frac': qbytes/(col_mean*1e-3)/1e9/hbm_peak,
                                'note': 'int8 digit planes of the rows (%d B per feature) + row scales; samples stay in L2' % digits})
        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel in profiles/r01_ncu_k_project_q_v4.txt: 911.4 MB for a
        # 1,000,000-row launch (= the algorithmic 904 MB; the operands are read exactly once); it scales linearly in rows
        per_row = {7: 911.4, 6: 784.0, 5: 911.4*5/7.}[digits]   # 6 digits: 777.5 + 6.5 MB, profiles/r02_ncu_k_project_q_lane_tables.txt
        roofline['traffic'] = per_row*n_local
        roofline['traffic_source'] = ('ncu --set full captures at 1M rows per launch (profiles/r02_ncu_k_project_q_lane_tables.txt: 784.0 MB for the 6-digit tier; '
                                      'profiles/r01_ncu_k_project_q_v4.txt: 911.4 MB for the 7-digit image; = the algorithmic bytes), scaled by rows; not '
                                      'measured in this run (DRAM counters need ncu)')
        roofline['bound_detail'] = ('FP64 pipe of the fused potential epilogue PLUS the tensor time of the int8 digit MMAs: on B200 both issue through '
                                    'the same SM pipe (ncu pipe_shared = fp64 + tensor; measured as a stall of the epilogue during the MMA windows in '
                                    'profiles/r01_q_ablation.txt), so the two add up -- see shared_pipe.frac; HBM is at 2 % of its peak')
    idcs_value = [int(i) for i in alg.idcs]
    if route == 'q' and not getattr(a, 'lean', False):
        # the other precision tiers of the same pass, timed right here on the resident rows (3 passes each after 2 warm-ups)
        tg, tiers = alg._tangent, {}
        for dg in (7, 6, 5):
            _fused.set_contraction_digits(dg)
            for _ in range(2):
                tg.fp.colsum_parts(tg.rows, None, out=tg.parts)
            t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0e.record()
            for _ in range(3):
                tg.fp.colsum_parts(tg.rows, None, out=tg.parts)
            t1e.record()
            torch.cuda.synchronize()
            ms = max_over_ranks(t0e.elapsed_time(t1e)/3.)
            tiers['digits_%d' % dg] = {'launch_ms': ms, 'digit_pairs': dg*(dg+1)//2, 'evals_per_s_kernel': n_local*S/(ms*1e-3),
                                       'fp64_equiv_tflops': flops/(ms*1e-3)/1e12}
        _fused.set_contraction_digits(digits)
        roofline['tiers'] = tiers

    # ---- stage 2 (materialised n x S matrix: snnls / Hilbert scoring) and stage 3 (coreset-side step), timed on their own ----
    stage2 = stage3 = None
    if rank == 0 and world == 1 and not getattr(a, 'lean', False):
        from bayesiancoresets._device import ptr, stream_ptr
        ctx2 = eng.ctx('bench-stage2')
        n2 = min(1_000_000, N)
        V = torch.randn(n2, S, dtype=torch.float64, device=dev)
        nrm = torch.empty(n2, dtype=torch.float64, device=dev)
        uu = torch.randn(2*S, dtype=torch.float64, device=dev)
        o4 = torch.zeros(4, dtype=torch.float64, device=dev)

        def ev_time(fn, reps):
            for _ in range(3):
                fn()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(reps):
                fn()
            a1.record()
            torch.cuda.synchronize()
            return a0.elapsed_time(a1)/reps
        nv.call('bc_dense_rownorms', ctx2, ptr(V), n2, S, S, ptr(nrm), stream_ptr())
        stage2 = {'bytes_per_launch': 8.*n2*S, 'rows': n2, 'S': S, 'peak_gbs': hbm_peak, 'bound': 'hbm', 'kernels': {}}
        for nm, mode in (('k_dense_score<FW> (frankwolfe.py:15-17)', nv.SCORE_FW), ('k_dense_score<GIGA> (giga.py:20-38)', nv.SCORE_GIGA),
                         ('k_dense_score<CORR> (bcores.py:78-81)', nv.SCORE_CORR)):
            ms = ev_time(lambda: nv.call('bc_dense_score', ctx2, mode, ptr(V), n2, S, S, ptr(nrm), ptr(uu), None, 0, ptr(o4), None,
                                         stream_ptr()), 10)
            stage2['kernels'][nm] = {'ms': ms, 'achieved_gbs': 8.*n2*S/ms/1e6, 'frac': 8.*n2*S/ms/1e6/hbm_peak}
        del V, nrm
        # stage 3: residual + gradient + ADAM on an M = 64 coreset (latency-bound: microseconds per optimiser step)
        M3 = 64
        Vc = torch.randn(M3, S, dtype=torch.float64, device=dev)
        cs_ = torch.randn(S, dtype=torch.float64, device=dev)
        w3 = torch.rand(M3, dtype=torch.float64, device=dev)
        m1 = torch.zeros(M3, dtype=torch.float64, device=dev)
        m2 = torch.zeros(M3, dtype=torch.float64, device=dev)
        g3 = torch.zeros(M3, dtype=torch.float64, device=dev)
        rs3 = torch.zeros(S+1, dtype=torch.float64, device=dev)

        def step3():
            nv.call('bc_core_resid', ctx2, ptr(cs_), 1.0, ptr(Vc), M3, S, S, ptr(w3), ptr(rs3), stream_ptr())
            nv.call('bc_core_grad', ctx2, ptr(Vc), M3, S, S, ptr(rs3), ptr(g3), stream_ptr())
            nv.call('bc_adam_step', ctx2, ptr(g3), ptr(w3), ptr(m1), ptr(m2), M3, 0.1, 0.9, 0.999, 0.1, 0.001, 1e-8, None, stream_ptr())
        stage3 = {'M': M3, 'S': S, 'us_per_step': 1e3*ev_time(step3, 50),
                  'kernels': 'k_core_resid + k_core_rows (gradient) + k_adam (bcores.py:145-146, util/opt.py:45-52)',
                  'bound': 'launch latency (3 launches, M x S = 0.5 MB)'}
        torch.cuda.empty_cache()

    # ------------------------------------------------- e2e: public API from host buffers --
    e2e = None
    if not a.no_e2e:
        Zhost = torch.empty(n_local, D, dtype=torch.float64, pin_memory=True)
        Zhost.copy_(Zdev)
        torch.cuda.synchronize()
        Zhost_np = Zhost.numpy()
        del alg, Zdev
        torch.cuda.empty_cache()
        evals2 = 0
        h2d = n_local*D*8
        d2h = 0
        barrier()
        t0 = time.perf_counter()
        alg2 = make_alg(DeviceRows(eng, Zhost_np, row0=row0, n_total=N))     # H2D of the data rows happens here
        for m in range(1, K+1):
            M0 = alg2.wts.shape[0]
            alg2.build(1, m)
            M1 = alg2.wts.shape[0]
            evals2 += step_evals(N, M1, S, a.opt_itrs)
            h2d += (1 + a.opt_itrs)*(S*D*8 + M1*D*8 + M1*8)
            d2h += a.opt_itrs*M1*8 + 16 + 8 + D*8
        w2, _, i2, _ = alg2.get()
        barrier()
        dt2 = max_over_ranks(time.perf_counter() - t0)
        e2e = {'value': evals2/dt2, 'unit': 'evals/s', 'h2d_bytes_per_step': h2d/K, 'd2h_bytes_per_step': d2h/K,
               'seconds': dt2, 'includes': 'upload of the %.2f GB row shard from pinned host memory (amortised over the %d steps), '
                                           'host sampler, sample uploads, weight read-backs' % (n_local*D*8/1e9, K),
               'indices': [int(i) for i in alg2.idcs],
               'indices_match_value_leg': [int(i) for i in alg2.idcs] == idcs_value[:len(alg2.idcs)]}
        Zdev = alg2.rows.t

    # ------------------------------------------------- cpu_baseline + parity on a bounded sample --
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        ns = min(a.cpu_rows, n_local)
        Zs = Zdev[:ns, :D].cpu().numpy().copy()
        best = None
        for threads in (blas_default, 1):
            if threadpool_limits is not None:
                with threadpool_limits(limits=threads, user_api='blas'):
                    o, ev, t = run_oracle_build(Zs, S, beta, a.cpu_opt_itrs, a.cpu_steps, 0)
            else:
                if threads == 1:
                    continue
                o, ev, t = run_oracle_build(Zs, S, beta, a.cpu_opt_itrs, a.cpu_steps, 0)
            used = threads
            if best is None or ev/t > best[0]:
                best = (ev/t, used, t)
        cpu = {'value': best[0], 'unit': 'evals/s', 'cores': best[1], 'kind': 'port',
               'sample': 'numpy oracle (restatement of the reference, pinned to it): %d build(1,m) steps on the first %d rows, '
                         'D=%d S=%d opt_itrs=%d (%.1f s); faster of default BLAS threads and 1 thread; host has %d cpus'
                         % (a.cpu_steps, ns, D, S, a.cpu_opt_itrs, best[2], os.cpu_count())}
        # the same sample through the CUDA path (tensor-core route): identical index sequence, weights within 1e-6
        prj = bc.BetaBlackBoxProjector(new_sampler(D, 1, a.sampler), S, model_lr.beta_likelihood, model_lr.log_likelihood, None)
        algs = bc.BetaCoreset(Zs, prj, opt_itrs=a.cpu_opt_itrs, step_sched=sched, beta=beta, learn_beta=False)
        for m in range(1, a.cpu_steps+1):
            algs.build(1, m)
        ow, _, oi = o.get()
        gw, _, gi, _ = algs.get()
        same = [int(i) for i in gi] == [int(i) for i in oi]
        nz = (np.abs(ow) > 0) if same else None
        parity = {'rows': ns, 'steps': a.cpu_steps, 'opt_itrs': a.cpu_opt_itrs, 'contraction_digits': digits,
                  'indices_equal': same, 'indices': [int(i) for i in gi], 'oracle_indices': [int(i) for i in oi],
                  'weights': [float(x) for x in gw], 'oracle_weights': [float(x) for x in ow],
                  'max_rel_weight_diff': float(np.max(np.abs(gw[nz]-ow[nz])/np.abs(ow[nz]))) if same and nz.any() else None,
                  'tolerance': 1e-6}

    if rank == 0:
        line = {
            'metric': 'beta_likelihood_evals_per_s', 'value': value, 'unit': 'evals/s', 'n_gpus': world, 'steps': K, 'warmup': W,
            'ms_per_step': 1e3*dt/K, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic', 'config': workload_config(a, world), 'build_seconds_per_point': dt/K,
            'selected_indices': idcs_value, 'roofline': roofline, 'step_breakdown': step_breakdown, 'gpu_launches': int(launches), 'clocks': clk,
            'host': {'cpu_count': os.cpu_count(), 'sampler_calls': sampler_calls, 'sampler_host_ms_per_call': 1e3*sampler_s/max(sampler_calls, 1),
                     'sampler_note': ('optimiser steps call the sampler through its device_step protocol: its algebra runs in kernels on the '
                                      'device-resident weights and the host only queues work -- it runs ahead of the GPU and then BLOCKS on it '
                                      '(launch queue, pinned normal buffers), so the host time per call is waiting, not work')
                                     if a.sampler in ('hybrid', 'device') else 'host sampler: the time per call is on the critical path of a step'},
        }
        if stage2 is not None:
            line['stage2'] = stage2
            line['stage3'] = stage3
        if e2e is not None:
            line['e2e'] = e2e
        if cpu is not None:
            line['cpu_baseline'] = cpu
        if parity is not None:
            line['parity'] = parity
        if emit:
            print(json.dumps(line))
        return line
    return None


def sweep_leg(a):
    """BASELINE config 5: N = 1M .. 100M rows of the north-star problem at this GPU count, one value-leg line per
    point (rows resident, same step, same timing rules as the default arm).  A point needs its fp64 row shard (8 D bytes per
    row) and the 7-plane int8 row image (7 D bytes per row) in HBM at once; one that does not fit is listed as skipped."""
    import copy
    import torch
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    ns = [1_000_000, 3_000_000, 10_000_000, 30_000_000, 100_000_000] if a.sweep == 'default' else [int(float(x)) for x in a.sweep.split(',')]
    out = a.sweep_out or os.path.join(ROOT, 'profiles', 'r02_sweep_n%d.jsonl' % world)
    free, total = torch.cuda.mem_get_info()
    rows = []
    for n in ns:
        need = (n/world)*a.d*(8 + 7) + (2 << 30)
        if need > total*0.97:
            line = {'N': n, 'n_gpus': world, 'skipped': 'row shard + row image need %.0f GB of the %.0f GB HBM of one GPU' % (need/1e9, total/1e9)}
        else:
            b = copy.copy(a)
            b.n, b.no_e2e, b.no_cpu_baseline, b.lean = n, True, True, True
            full = b200_arm(b, emit=False)
            torch.cuda.empty_cache()
            if rank != 0:
                continue
            r = full['roofline']
            line = {'N': n, 'n_gpus': world, 'D': a.d, 'S': a.s, 'opt_itrs': a.opt_itrs, 'steps': a.steps, 'warmup': a.warmup,
                    'evals_per_s': full['value'], 'build_seconds_per_point': full['build_seconds_per_point'],
                    'pass_ms': r['launch_ms'], 'kernel_evals_per_s': r['evals_per_s_kernel'], 'fp64_equiv_tflops': r['achieved'],
                    'share_of_step_in_passes': r['share_of_step'], 'contraction_digits': r.get('contraction_digits'),
                    'selected_indices': full['selected_indices'], 'clocks': full['clocks']}
        if rank == 0:
            rows.append(line)
            with open(out, 'a') as f:
                f.write(json.dumps(line) + '\n')
            print(json.dumps(line), flush=True)


# ------------------------------------------------------------------- BASELINE configs 1-4 --
def configs_leg(a):
    """build seconds of BASELINE.json's configs 1-4 on one GPU with the numpy port of the reference beside each (the CPU
    arm runs fewer build steps of the same problem and is compared per build step).  One JSON line per config.  The
    problems are the seeded ones of tests/golden/problems.py (config 1 = examples/zellner_gaussian/main.py restated line
    by line; config 2 = the tests/test_snnls matrix; configs 3 / 4 = SURVEY 8d's generators at 100K / 1M rows)."""
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
    import problems
    import bayesiancoresets as bc
    import model_lr, gaussian, model_neurlinr
    from bayesiancoresets import _native as nv
    from oracle import np_snnls as osn, np_coresets as oc
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=1, user_api='blas')     # D x D samplers: BLAS thread pools only thrash
    except Exception:
        pass
    torch.cuda.set_device(0)

    def potentials(prob):
        if prob['model'] == 'lr':
            return model_lr.beta_likelihood, model_lr.log_likelihood
        if prob['model'] == 'gauss':
            return gaussian.gaussian_beta_likelihood.bind(**prob['params']), gaussian.gaussian_loglikelihood.bind(**prob['params'])
        return model_neurlinr.neurlinr_beta_likelihood.bind(**prob['params']), model_neurlinr.neurlinr_loglikelihood.bind(**prob['params'])

    def product_sampler(prob):
        """the sampler as a user of this package writes it: same algebra and same numpy stream as the plain host callback
        the CPU arm runs (so both arms draw the same samples), formed on the device with the normals drawn a call ahead"""
        from bayesiancoresets.util import rng
        rng.drain()
        if prob['model'] == 'gauss':
            return gaussian.make_conjugate_sampler(prob['prior']['mu0'], prob['prior']['Sig0inv'], prob['params']['Siginv'], device=True, prefetch=True)
        if prob['model'] == 'nl':
            return model_neurlinr.make_conjugate_sampler(prob['prior']['mu0'], prob['prior']['Sig0inv'], prob['params']['sigsq'], device=True, prefetch=True)
        return model_lr.make_laplace_sampler(prob['data'].shape[1], method='hybrid', prefetch=True)

    def greedy(case, gpu_steps, cpu_steps, label):
        prob = case['make']()
        problems.reseed(case)
        bl, ll = potentials(prob)
        prj = bc.BetaBlackBoxProjector(product_sampler(prob), case['S'], bl, ll, None)
        alg = bc.BetaCoreset(prob['data'], prj, n_subsample_select=case['n_sel'], n_subsample_opt=case['n_opt'], opt_itrs=case['opt_itrs'],
                             step_sched=case['sched'], beta=case['beta'], learn_beta=False)
        alg.build(1, 1)                       # first step: uploads, row image, lazy allocations
        torch.cuda.synchronize()
        l0 = nv.lib().bc_launch_count()
        t0 = time.perf_counter()
        for m in range(2, gpu_steps+1):
            alg.build(1, m)
        torch.cuda.synchronize()
        g = (time.perf_counter()-t0)/max(gpu_steps-1, 1)
        launches = nv.lib().bc_launch_count() - l0
        gi = [int(i) for i in alg.idcs]
        prj.sampler.drain()
        prob = case['make']()
        problems.reseed(case)
        o = oc.GreedyVI(prob['data'], prob['sampler'], case['S'], prob['oracle_betalik'](case['beta']), n_sub_select=case['n_sel'],
                        n_sub_opt=case['n_opt'], opt_itrs=case['opt_itrs'], sched=case['sched'])
        with np.errstate(all='ignore'):
            o.build(1, 1)
            t0 = time.perf_counter()
            for m in range(2, cpu_steps+1):
                o.build(1, m)
        c = (time.perf_counter()-t0)/max(cpu_steps-1, 1)
        oi = [int(i) for i in o.idcs]
        n = min(len(gi), len(oi))
        print(json.dumps({'config': label, 'rows': int(prob['data'].shape[0]), 'D': int(prob['data'].shape[1]), 'S': case['S'],
                          'opt_itrs': case['opt_itrs'], 'subsample': [case['n_sel'], case['n_opt']],
                          'b200_s_per_build_step': g, 'b200_build_steps_timed': gpu_steps-1, 'b200_ms_per_optimiser_step': 1e3*g/(1+case['opt_itrs']),
                          'b200_launches_per_build_step': launches/max(gpu_steps-1, 1),
                          'cpu_port_s_per_build_step': c, 'cpu_build_steps_timed': cpu_steps-1, 'cpu_threads': 1, 'speedup': c/g,
                          'indices_equal_on_common_prefix': gi[:n] == oi[:n], 'indices': gi[:12]}), flush=True)

    cases = {c['name']: c for c in problems.coreset_cases(True)}
    greedy(cases['c1_zellner_gaussian'], 12, 3, 'C1 examples/zellner_gaussian (N=5700 d=100 S=200, 1000 ADAM steps per point, sub-samples 1000/200)')
    # C2: snnls on the 1000 x 100 matrix of tests/test_snnls
    V = problems.snnls_matrix()
    for name, cls in (('giga', bc.snnls.GIGA), ('fw', bc.snnls.FrankWolfe), ('omp', bc.snnls.OrthoPursuit)):
        alg = cls(V.T, V.sum(axis=0)); alg.build(3); alg.reset()
        torch.cuda.synchronize(); t0 = time.perf_counter(); alg.build(100); torch.cuda.synchronize(); g = time.perf_counter()-t0
        o = osn.SOLVERS[name](V.T, V.sum(axis=0)); t0 = time.perf_counter(); o.run(100); c = time.perf_counter()-t0
        print(json.dumps({'config': 'C2 snnls %s, A = 100 x 1000, build(100)' % name, 'b200_s': g, 'cpu_port_s': c, 'speedup': c/g,
                          'weights_rel_diff': float(np.abs(alg.weights()-o.w).max()/np.abs(o.w).max()),
                          'note': 'latency-bound at this size: 0.8 MB of data, a handful of launches and scalar read-backs per iteration'}), flush=True)
    greedy(cases['c3_logreg_100k'], 5, 3, 'C3 examples/zellner_logreg (N=100K D=20 S=100 beta=0.1, full data, Laplace sampler)')
    c4 = dict(cases['c4_neurlin_100k'], make=problems.make_neurlin(1_000_000, 64, 42))
    c4cpu = cases['c4_neurlin_100k']
    # config 4 at its stated 1M rows on the GPU; the CPU port is timed on 100K rows and scaled by 10 (its cost is linear in rows)
    prob = c4['make']()
    problems.reseed(c4)
    bl, ll = potentials(prob)
    prj = bc.BetaBlackBoxProjector(product_sampler(prob), c4['S'], bl, ll, None)
    alg = bc.BetaCoreset(prob['data'], prj, opt_itrs=c4['opt_itrs'], step_sched=c4['sched'], beta=c4['beta'], learn_beta=False)
    alg.build(1, 1); torch.cuda.synchronize(); t0 = time.perf_counter()
    for m in range(2, 6):
        alg.build(1, m)
    torch.cuda.synchronize(); g = (time.perf_counter()-t0)/4.
    prj.sampler.drain()
    prob = c4cpu['make'](); problems.reseed(c4cpu)
    o = oc.GreedyVI(prob['data'], prob['sampler'], c4cpu['S'], prob['oracle_betalik'](c4cpu['beta']), opt_itrs=c4cpu['opt_itrs'], sched=c4cpu['sched'])
    o.build(1, 1); t0 = time.perf_counter(); o.build(1, 2); c = (time.perf_counter()-t0)
    print(json.dumps({'config': 'C4 examples/zellner_neural_linear (N=1M D=64 S=256 beta=0.2, conjugate sampler)', 'rows': 1_000_000,
                      'b200_s_per_build_step': g, 'b200_ms_per_optimiser_step': 1e3*g/(1+c4['opt_itrs']),
                      'cpu_port_s_per_build_step_100k_rows': c, 'cpu_port_s_per_build_step_scaled_to_1M': 10*c, 'speedup': 10*c/g,
                      'indices': [int(i) for i in alg.idcs]}), flush=True)


def _quiet_stdout():
    """libraries (NCCL's version banner, ...) write to fd 1: point it at stderr while the bench runs so that stdout carries
    exactly one JSON line; returns a file object on the real stdout"""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    return real


if __name__ == '__main__':
    args = parse()
    sys.stdout = _quiet_stdout()
    if args.configs:
        configs_leg(args)
    elif args.impl == 'reference':
        reference_arm(args)
    elif args.sweep:
        sweep_leg(args)
    else:
        b200_arm(args)
    try:
        import torch.distributed as _dist
        if _dist.is_available() and _dist.is_initialized():
            _dist.destroy_process_group()
    except Exception:
        pass
