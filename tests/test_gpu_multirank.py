"""Multi-rank parity on the GPU: a world-2 job (datapoints sharded over the ranks) must build exactly the coresets the
one-rank job builds -- BetaCoreset (full data, sub-sampled), group mode, HilbertCoreset with GIGA / Frank-Wolfe /
OrthoPursuit, and the solvers on a host matrix.  NCCL when the box has two GPUs, else gloo with both ranks on cuda:0."""
import json
import os
import socket
import subprocess
import sys
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(world, path):
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, 'multirank_worker.py'), path], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(o)
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, 'rank %d of %d failed:\n%s' % (r, world, o[-4000:])
    return json.load(open(path))


def test_two_ranks_build_the_same_coresets_as_one(tmp_path):
    one = _run(1, str(tmp_path/'w1.json'))
    two = _run(2, str(tmp_path/'w2.json'))
    assert two['world'] == 2 and two['backend'] in ('nccl', 'gloo')
    keys = [k for k in one if isinstance(one[k], dict)]
    assert len(keys) >= 13
    for k in keys:
        assert one[k]['idcs'] == two[k]['idcs'], k
        np.testing.assert_allclose(two[k]['wts'], one[k]['wts'], rtol=1e-9, atol=1e-12, err_msg=k)
        if 'error' in one[k]:
            np.testing.assert_allclose(two[k]['error'], one[k]['error'], rtol=1e-9, atol=1e-12, err_msg=k)
    assert len(one['beta_full']['idcs']) >= 3 and len(one['hilbert_giga']['idcs']) >= 5      # the builds did select points
