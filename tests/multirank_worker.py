"""One rank of the multi-rank parity job (tests/test_gpu_multirank.py launches WORLD_SIZE of these).  Every rank is handed
the same host data and the same numpy seeds, shards the datapoints like a production job, and rank 0 writes the resulting
coresets to the JSON file named on the command line.  Backend: NCCL when every rank has its own GPU, else gloo with the
ranks sharing cuda:0 (collectives staged through the host: same sharded code path, different transport)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'beta-cores_b200'), os.path.join(ROOT, 'beta-cores_b200', 'examples', 'common'),
          os.path.join(ROOT, 'tests', 'golden')):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist

rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
ngpu = torch.cuda.device_count()
local = rank if ngpu >= world else 0
torch.cuda.set_device(local)
if world > 1:
    if ngpu >= world:
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', local))
    else:
        dist.init_process_group('gloo', rank=rank, world_size=world)
import bayesiancoresets as bc
import model_lr
import problems

out = {'world': world, 'backend': dist.get_backend() if world > 1 else 'none'}
sched = lambda i: 1./(1.+i)


def record(name, alg):
    r = alg.get()
    out[name] = {'idcs': [int(i) for i in r[2]], 'wts': [float(x) for x in r[0]]}


prob = problems.make_logistic(3001, 7, 5)()
Z, sampler = prob['data'], prob['sampler']
S = 48
# 1. beta-Cores, full data (fused passes over the row shards; one double-double part + one (score, index) pair per step)
np.random.seed(3)
alg = bc.BetaCoreset(Z, bc.BetaBlackBoxProjector(sampler, S, model_lr.beta_likelihood, model_lr.log_likelihood, None), opt_itrs=8,
                     step_sched=sched, beta=0.2, learn_beta=False)
for m in range(1, 6):
    alg.build(1, m)
record('beta_full', alg)
# 2. sub-sampled selection / optimisation (each rank gathers the sub-sampled rows it owns)
np.random.seed(4)
alg = bc.BetaCoreset(Z, bc.BetaBlackBoxProjector(sampler, S, model_lr.beta_likelihood, model_lr.log_likelihood, None), opt_itrs=8,
                     n_subsample_select=400, n_subsample_opt=150, step_sched=sched, beta=0.2, learn_beta=False)
for m in range(1, 6):
    alg.build(1, m)
record('beta_sub', alg)
# 2b. the host-free optimiser loop: a sampler with the device_step protocol (two-phase steps around the exchange when sharded)
np.random.seed(7)
smp = model_lr.make_laplace_sampler(7, method='hybrid', prefetch=True)
alg = bc.BetaCoreset(Z, bc.BetaBlackBoxProjector(smp, S, model_lr.beta_likelihood, model_lr.log_likelihood, None), opt_itrs=8,
                     n_subsample_opt=700, step_sched=sched, beta=0.2, learn_beta=False)
for m in range(1, 6):
    alg.build(1, m)
smp.drain()
record('beta_device_loop', alg)
# 3. group-wise selection
np.random.seed(5)
groups = problems.ragged_groups(3001, 9)
alg = bc.SparseVICoreset(Z, bc.BlackBoxProjector(sampler, S, model_lr.log_likelihood, None), opt_itrs=6, step_sched=sched, groups=groups)
for m in range(1, 4):
    alg.build(1, 10**9)
record('svi_groups', alg)
# 4. Hilbert coresets: the datapoint-major projection matrix is sharded; GIGA / Frank-Wolfe / OrthoPursuit on it
for name, solver in (('giga', bc.snnls.GIGA), ('fw', bc.snnls.FrankWolfe), ('omp', bc.snnls.OrthoPursuit)):
    np.random.seed(6)
    alg = bc.HilbertCoreset(Z, bc.BlackBoxProjector(sampler, S, model_lr.log_likelihood, None), snnls=solver)
    for m in range(1, 13):
        alg.build(1, m)
    record('hilbert_'+name, alg)
    out['hilbert_'+name]['error'] = float(alg.error())
    alg.optimize()
    record('hilbert_'+name+'_opt', alg)
# 5. the solvers on a host matrix (every rank keeps its block of columns), more ranks than a tiny problem has columns included
V = problems.snnls_matrix()
for name, solver in (('giga', bc.snnls.GIGA), ('fw', bc.snnls.FrankWolfe), ('omp', bc.snnls.OrthoPursuit)):
    s = solver(V.T, V.sum(axis=0))
    s.build(40)
    w = s.weights()
    out['snnls_'+name] = {'idcs': [int(i) for i in np.nonzero(w)[0]], 'wts': [float(x) for x in w[w > 0]], 'error': float(s.error())}
tiny = np.random.RandomState(1).randn(1, 5)
s = bc.snnls.GIGA(tiny.T, tiny.sum(axis=0))
s.build(1)
out['snnls_one_column'] = {'idcs': [int(i) for i in np.nonzero(s.weights())[0]], 'wts': [float(x) for x in s.weights()]}
torch.cuda.synchronize()
if rank == 0:
    with open(sys.argv[1], 'w') as f:
        json.dump(out, f)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
