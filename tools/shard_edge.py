"""2-rank sanity: a dataset with fewer rows than ranks x tile (one rank nearly / completely empty) builds the same coreset as one rank"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'beta-cores_b200'), os.path.join(ROOT, 'beta-cores_b200', 'examples', 'common')):
    sys.path.insert(0, p)
import numpy as np, torch, torch.distributed as dist
rank = int(os.environ.get('RANK', 0)); world = int(os.environ.get('WORLD_SIZE', 1)); local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
import bayesiancoresets as bc, model_lr
from bayesiancoresets._device import Engine, DeviceRows
from bayesiancoresets._shard import partition_rows
eng = Engine.get()
for N in (1, 3, 300):
    r = np.random.RandomState(N)
    D, S = 6, 40
    Z = r.randn(N, D)
    row0, n_local = partition_rows(N, world, rank)
    rows = DeviceRows(eng, Z[row0:row0+n_local], row0=row0, n_total=N)
    np.random.seed(4)
    prj = bc.BetaBlackBoxProjector(model_lr.make_laplace_sampler(D, method='newton'), S, model_lr.beta_likelihood, model_lr.log_likelihood, None)
    alg = bc.BetaCoreset(rows, prj, opt_itrs=5, step_sched=lambda i: 1./(1.+i), beta=0.3, learn_beta=False)
    for m in range(1, min(N, 4)+1):
        alg.build(1, m)
    if rank == 0:
        print('N', N, 'world', world, 'idcs', alg.idcs.tolist(), 'wts', np.round(alg.wts, 10).tolist(), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
