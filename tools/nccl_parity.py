"""Multi-GPU (NCCL) parity run against the reference's golden vectors, one process per GPU:

    torchrun --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 tools/nccl_parity.py

Runs every golden case whose grid has WORLD_SIZE ranks through PyNMF.fit over NCCL (1-D and 2-D grids, with CUDA
graphs) and prints one line per case plus a summary; exit code 1 on any tolerance violation."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import cases as C                      # noqa: E402
from tests import common as T                      # noqa: E402
from tests.test_parity_gpu import _tol             # noqa: E402
from tests.workers import fit_worker               # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', rank)))
    import tests.workers as Wk
    # fit_worker pins cuda:0 (several ranks on one GPU in the test-suite); here every rank has its own device
    orig = torch.cuda.set_device
    torch.cuda.set_device = lambda d: None
    bad = 0
    todo = [c for c in C.CASES if c['grid'][0] * c['grid'][1] == world and c['itr'] in (10, 100)]
    for case in todo:
        res = fit_worker(rank, world, case)
        g = T.golden_case(case['name'])[rank]
        tf, te = _tol(case)
        dW, dH = T.rel_fro(res['W'], g['W']), T.rel_fro(res['H'], g['H'])
        de = abs(res['err'] - float(g['err'])) / abs(float(g['err']))
        ok = dW <= tf and dH <= tf and de <= max(te, 1e-5 if case['method'] == 'bcd' else te)
        flag = torch.tensor([0 if ok else 1], device='cuda')
        dist.all_reduce(flag)
        if rank == 0:
            print('%-44s backend=%s relW=%.2e relH=%.2e rel_err=%.2e %s' % (case['name'], dist.get_backend(), dW, dH, de,
                                                                        'ok' if flag.item() == 0 else 'FAIL'), flush=True)
        bad += int(flag.item() != 0)
    torch.cuda.set_device = orig
    if rank == 0:
        print('nccl parity: %d cases, %d failed' % (len(todo), bad), flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(1 if bad else 0)


if __name__ == '__main__':
    main()
