"""Multi-GPU (NCCL) parity of the NMFk-level rows against the reference's golden vectors, one process per GPU:

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/nccl_nmfk_parity.py

Runs the 2-rank clustering, nnsvd, nnsvd-initialised fits and NMFk end-to-end cases of tests/test_nmfk_gpu.py with the
collectives on NCCL (the test-suite runs them over gloo on one GPU) plus the replica-mode ensemble, prints one line per
case and a summary; exit code 1 on any tolerance violation."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import nmfk_cases as K                 # noqa: E402
from tests import common as T                      # noqa: E402
from tests import workers as Wk                    # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', rank)))
    torch.cuda.set_device = lambda d: None          # the workers pin cuda:0 (several ranks per GPU in the test-suite)
    with np.load(os.path.join(K.GOLDEN, 'nmfk_cases.npz')) as z:
        gold = {k: z[k] for k in z.files}
    bad = [0]

    def report(name, ok, msg):
        flag = torch.tensor([0 if ok else 1], device='cuda')
        dist.all_reduce(flag)
        if rank == 0:
            print('%-40s backend=%s %s %s' % (name, dist.get_backend(), msg, 'ok' if flag.item() == 0 else 'FAIL'), flush=True)
        bad[0] += int(flag.item() != 0)

    def unwrap(res, name):
        tag, val = res[name]
        if tag == 'err':
            print(val, flush=True)
            return None
        return val

    cases = [c for c in K.CLUSTER_CASES if c['p_r'] == world]
    res = Wk.cluster_worker(rank, world, cases)
    for c in cases:
        o = unwrap(res, c['name'])
        g = lambda key: gold['cluster/%s/%d/%s' % (c['name'], rank, key)]    # noqa: E731
        tol = 1e-10 if c['dtype'] == 'float64' else 1e-5
        ok = o is not None and np.array_equal(o['order'], g('order')) and T.rel_fro(o['W_all'], g('W_all')) <= tol \
            and T.rel_fro(o['H_all'], g('H_all')) <= tol and np.allclose(o['sils'], g('sils'), rtol=0, atol=1e-4)
        report('cluster/' + c['name'], ok, '')
    cases = [c for c in K.NNSVD_CASES if c['grid'][0] * c['grid'][1] == world]
    res = Wk.nnsvd_worker(rank, world, cases)
    for c in cases:
        o = unwrap(res, c['name'])
        dW = T.rel_fro(o['W'], gold['nnsvd/%s/%d/W' % (c['name'], rank)]) if o else 1.0
        dH = T.rel_fro(o['H'], gold['nnsvd/%s/%d/H' % (c['name'], rank)]) if o else 1.0
        report('nnsvd/' + c['name'], dW <= 1e-3 and dH <= 1e-3, 'relW=%.2e relH=%.2e' % (dW, dH))
    cases = [c for c in K.NNSVD_FIT_CASES if c['grid'][0] * c['grid'][1] == world]
    res = Wk.nnsvd_fit_worker(rank, world, cases)
    for c in cases:
        o = unwrap(res, c['name'])
        dW = T.rel_fro(o['W'], gold['nnsvdfit/%s/%d/W' % (c['name'], rank)]) if o else 1.0
        dH = T.rel_fro(o['H'], gold['nnsvdfit/%s/%d/H' % (c['name'], rank)]) if o else 1.0
        report('nnsvdfit/' + c['name'], dW <= 2e-3 and dH <= 2e-3, 'relW=%.2e relH=%.2e' % (dW, dH))
    for c in [c for c in K.E2E_CASES if c['grid'][0] * c['grid'][1] == world]:
        tmp = [tempfile.mkdtemp() if rank == 0 else None]
        dist.broadcast_object_list(tmp, src=0)
        o = Wk.nmfk_e2e_worker(rank, world, c, tmp[0])
        ok = o['nopt'] == int(gold['e2e/%s/%d/nopt' % (c['name'], rank)])
        worst = 0.0
        for k in range(c['start_k'], c['end_k'] + 1):
            worst = max(worst, T.rel_fro(o['k%d/W_reg' % k], gold['e2e/%s/%d/k%d/W_reg' % (c['name'], rank, k)]),
                        T.rel_fro(o['k%d/H_reg' % k], gold['e2e/%s/%d/k%d/H_reg' % (c['name'], rank, k)]))
        report('e2e/' + c['name'], ok and worst <= 2e-2, 'nopt=%d worst rel diff of the regression factors=%.2e' % (o['nopt'], worst))
    # replica-mode ensemble: perturbations spread over the GPUs, identical to the sequential stacking
    seq = Wk.ensemble_worker(rank, world, False)
    par = Wk.ensemble_worker(rank, world, True)
    report('ensemble/replicas', np.array_equal(seq['Wall'], par['Wall']) and np.array_equal(seq['Hall'], par['Hall'])
           and seq['errs'] == par['errs'], '')
    if rank == 0:
        print('nccl nmfk parity: %d failed' % bad[0], flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(1 if bad[0] else 0)


if __name__ == '__main__':
    main()
