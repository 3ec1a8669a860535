"""BASELINE.json configs[4] exactly as stated, on N GPUs (N even):

    data/wtsi.mat (96 x 21) NMFk k-sweep start_k=2 end_k=10, 20 uniform perturbations (noise_var 0.015), KL-MU, nnsvd init,
    1000 iterations per fit, sill_thr 0.9 -- examples/dist_pynmfk_1d_wtsi.py:26-44 of the reference expects nopt == 4.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/bench_cfg5.py

nnsvd needs the 2 x 1 grid the reference asserts for this matrix (dist_svd.py:54-57), so the world is cut into N / 2
replica groups of 2 x 1 ranks (params.ensemble_parallel): every group holds the whole matrix on its own grid and takes
every (N/2)-th perturbation; clustering, the W-fixed regression and the rank selection follow the reference.  Prints one
JSON line: wall time of PyNMFk.fit(), nopt, and the per-k statistics against tests/golden/nmfk_cfg5.npz (the unmodified
reference's run of the same configuration, oracle/gen_golden_nmfk.py cfg5) when that file exists."""
import json
import os
import random
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pydnmfk_b200 import _lib as L  # noqa: E402
from pydnmfk_b200.data_io import read_results  # noqa: E402
from pydnmfk_b200.dist_comm import MPI, MPI_comm  # noqa: E402
from pydnmfk_b200.pyDNMFk import PyNMFk  # noqa: E402
from pydnmfk_b200.utils import parse, determine_block_params  # noqa: E402


GOLDEN = os.path.join(ROOT, 'tests', 'golden')          # fixtures only: the example matrix and the reference's statistics
CFG5_CASE = dict(name='wtsi_2x1_nnsvd_cfg5', grid=(2, 1), init='nnsvd', start_k=2, end_k=10, perturbations=20, itr=1000,
                 noise_var=0.015, sill_thr=0.9, norm='kl', method='mu')      # = oracle/nmfk_cases.py CFG5_CASE (the golden's name)
NNSVD_PY_SEED = 4321                                     # `random.seed` before DistSVD draws its start vectors, as in the golden run


def main():
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    case = dict(CFG5_CASE)
    if '--quick' in sys.argv:                       # script check: a shortened sweep
        case.update(end_k=4, perturbations=4, itr=100)
    if '--quick1000' in sys.argv:                   # profiling: full-length fits, short sweep
        case.update(end_k=3, perturbations=4, itr=1000)
    p_r, p_c = case['grid']
    assert world % (p_r * p_c) == 0, 'world must be a multiple of the 2 x 1 factorization grid'
    X = np.load(os.path.join(GOLDEN, 'wtsi_X.npy')).astype('float32')    # data/wtsi.mat['X'] of the reference, 96 x 21
    pos = rank % (p_r * p_c)
    b = determine_block_params(pos, (p_r, p_c), X.shape).determine_block_index_range_asymm()
    A_ij = np.ascontiguousarray(X[b[0][0]:b[1][0] + 1, b[0][1]:b[1][1] + 1])
    comm = MPI.COMM_WORLD

    def run(tmp, tag):
        p = parse()
        p.comm1 = comm
        p.size, p.rank, p.p_r, p.p_c = world, rank, p_r, p_c
        if world == p_r * p_c:                           # one group: the plain grid of the reference
            grid = MPI_comm(comm, p_r, p_c)
            p.comm, p.row_comm, p.col_comm = grid, grid.cart_1d_row(), grid.cart_1d_column()
        else:
            p.comm = p.row_comm = p.col_comm = None      # set per replica group by PyNMFk
        p.ensemble_parallel = True
        p.fpath, p.fname, p.ftype = 'data/', 'wtsi_' + tag, 'mat'
        p.init, p.itr, p.norm, p.method, p.verbose = case['init'], case['itr'], case['norm'], case['method'], False
        p.start_k, p.end_k, p.step_k, p.sill_thr = case['start_k'], case['end_k'], 1, case['sill_thr']
        p.perturbations, p.noise_var, p.sampling = case['perturbations'], case['noise_var'], 'uniform'
        p.results_path, p.checkpoint, p.precision = tmp + '/', False, 'float32'
        random.seed(NNSVD_PY_SEED)
        comm.barrier()
        torch.cuda.synchronize()
        L.launch_count(reset=True)
        t0 = time.perf_counter()
        nopt = PyNMFk(A_ij, factors=None, params=p).fit()
        torch.cuda.synchronize()
        comm.barrier()
        return time.perf_counter() - t0, int(nopt), L.launch_count()

    tmp = tempfile.mkdtemp() if rank == 0 else None
    tmp = comm.bcast(tmp, root=0)
    t_warm, _, _ = run(tmp, 'warm') if '--no-warmup' not in sys.argv else (0.0, 0, 0)
    dt, nopt, launches = run(tmp, 'run')
    if rank == 0:
        line = {'config': 'cfg5: wtsi 96x21 fp32, PyNMFk.fit k=%d..%d, %d perturbations, %s-%s itr=%d, %s init, sill_thr %.2f, '
                          'noise_var %.3f; %d GPUs = %d replica groups of a %dx%d grid'
                          % (case['start_k'], case['end_k'], case['perturbations'], case['norm'].upper(), case['method'].upper(),
                             case['itr'], case['init'], case['sill_thr'], case['noise_var'], world, world // (p_r * p_c), p_r, p_c),
                'n_gpus': world, 'nmfk_wall_s': dt, 'first_run_wall_s': t_warm, 'nopt': nopt, 'expected_nopt_reference_example': 4,
                'perturbation_fits': (case['end_k'] - case['start_k'] + 1) * case['perturbations'],
                'gpu_launches_rank0': int(launches), 'reference_wall_s_2_cpu_ranks_authoring_container': 72.2}
        gpath = os.path.join(GOLDEN, 'nmfk_cfg5.npz')
        if os.path.exists(gpath) and '--quick' not in sys.argv and '--quick1000' not in sys.argv:
            g = np.load(gpath)
            pre = 'e2e/%s/0/' % case['name']
            line['reference_nopt'] = int(g[pre + 'nopt'])
            cmp_ = {}
            for k in range(case['start_k'], case['end_k'] + 1):
                res = read_results('%s/wtsi_run/%d/' % (tmp, k))
                sil = np.asarray(res['clusterSilhouetteCoefficients'], dtype=np.float64)
                gs = np.asarray(g[pre + 'k%d/clusterSilhouetteCoefficients' % k], dtype=np.float64)
                cmp_[str(k)] = {'min_sil': float(sil.min()), 'ref_min_sil': float(gs.min()),
                                'L_errDist': float(np.asarray(res['L_errDist'])), 'ref_L_errDist': float(g[pre + 'k%d/L_errDist' % k])}
            line['per_k_vs_reference'] = cmp_
        print(json.dumps(line), flush=True)
    comm.barrier()
    sys.stdout.flush()
    if '--quick1000' in sys.argv:
        return
    os._exit(0)


if __name__ == '__main__':
    main()
