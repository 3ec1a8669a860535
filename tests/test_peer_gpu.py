"""The row grid's H half-step as one exchange over peer-mapped memory (pydnmfk_b200/peer.py, dnmf_xchg_update_h)
and the NCCL communicators owned by libdnmf.so (dnmf_comm_*).

On the one-GPU test box several ranks share cuda:0: CUDA IPC maps each rank's exchange region into the others, the
kernels' flag waits are resolved by the driver's time-slicing between the processes.  With two or more GPUs visible
the same cases (and the library's NCCL collectives) run one rank per GPU under torchrun."""
import os
import subprocess
import sys

import pytest

from oracle import cases as C
from tests import mp_util, workers
from tests.test_parity_gpu import _compare

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PEER_CASES = [c for c in C.CASES if c['grid'][1] == 1 and c['grid'][0] > 1 and c['method'] in ('mu', 'hals')
              and c['itr'] in (1, 10) and c['dtype'] == 'float32' and not c['prune'] and not c['expect_tc']]
PEER_CASES += [C.CASES_BY_NAME[n] for n in ('u64x48k4_2x1_fro_mu_i10_64', 'u64x48k4_2x1_kl_mu_i10_64', 'u2048k32_2x1_fro_mu_i10',
                                            'u2048k32_2x1_kl_mu_i10', 'zeros40x36k3_2x1_kl_mu_prune')]
_cache = {}


def _results(case):
    world = case['grid'][0]
    if world not in _cache:
        batch = [c for c in PEER_CASES if c['grid'][0] == world]
        _cache[world] = mp_util.run(world, workers.fit_many_worker, (batch,), backend='gloo', timeout=1500,
                                    env={'DNMF_PEER_EXCHANGE': '1'})
    res = []
    for r in range(world):
        if case['name'] not in _cache[world][r]:
            pytest.fail('case did not run (an earlier case of this batch failed on rank %d)' % r)
        tag, val = _cache[world][r][case['name']]
        if tag == 'err':
            pytest.fail('rank %d raised:\n%s' % (r, val))
        res.append(val)
    return res


@pytest.mark.parametrize('case', PEER_CASES, ids=[c['name'] for c in PEER_CASES])
def test_peer_exchange_matches_reference(case):
    res = _results(case)
    assert all(r['peer_exchange'] for r in res), 'the H half-step did not take the peer-memory exchange'
    _compare(case, res)


def _torchrun(nproc, script, *args, timeout=900):
    import torch
    if torch.cuda.device_count() < nproc:
        pytest.skip('needs %d GPUs' % nproc)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(nproc), '--master-addr',
           '127.0.0.1', '--master-port', str(mp_util._free_port()), os.path.join(ROOT, script)] + list(args)
    p = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)
    assert p.returncode == 0, p.stdout[-6000:]
    return p.stdout


def test_nccl_parity_two_gpus():
    """Golden cases of every 2-rank grid over NCCL (library communicators + peer exchange), one rank per GPU."""
    out = _torchrun(2, 'tools/nccl_parity.py')
    assert 'failed' in out and ', 0 failed' in out, out[-3000:]


def test_nccl_parity_four_gpus():
    out = _torchrun(4, 'tools/nccl_parity.py')
    assert ', 0 failed' in out, out[-3000:]


def test_library_collectives_two_gpus():
    """dnmf_allreduce / allgather / reduce_scatter / bcast / comm_split of libdnmf.so against torch.distributed."""
    out = _torchrun(2, 'tools/comm_check.py')
    assert 'comm check ok' in out, out[-3000:]
