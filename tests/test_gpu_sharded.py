"""GPU: row-sharded tables.  World size 1 runs everywhere; world size 2 needs two GPUs (skipped otherwise)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _run(world, name):
    env = dict(os.environ)
    env.pop('RANK', None)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world), '--master-addr',
           '127.0.0.1', '--master-port', str(29600 + world), os.path.join(HERE, 'dist_shard_check.py'), name]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0 and 'SHARD_OK' in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]


@pytest.mark.parametrize('name', ['sasrec_softmax', 'sasrec_softmax_d128', 'gru_softmax_h32'])
def test_sharded_engine_world1(name):
    _run(1, name)


@pytest.mark.parametrize('name', ['sasrec_softmax', 'sasrec_softmax_d128'])
def test_sharded_engine_world2(name):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    _run(2, name)
