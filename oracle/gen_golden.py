"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

TEST INFRASTRUCTURE.  Run in the authoring container only (needs
``/root/reference``):

    python oracle/gen_golden.py            # all cases
    python oracle/gen_golden.py NAME ...   # a subset (merged into the file)

For every case of ``oracle/cases.py`` the reference's own ``PyNMF(...).fit()``
(pyDNMF.py:55,138) is executed on p_r*p_c forked ranks under the mpi4py
stand-in (oracle/refrun) and the per-rank outputs ``W, H, recon_err`` plus the
integer shard geometry and prune masks that ``data_operations`` left on
``params`` are stored.  Also stores known answers for ``determine_block_params``
and the ``sample`` perturbation.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import cases as C            # noqa: E402
from oracle.refrun.launch import run_ranks, reference_available  # noqa: E402

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def _ref_fit(rank, size, case):
    """Runs inside a forked rank with the reference importable."""
    import numpy as np
    from pyDNMFk.pyDNMF import PyNMF
    from pyDNMFk.dist_comm import MPI_comm
    from pyDNMFk.utils import parse, determine_block_params
    from mpi4py import MPI
    p_r, p_c = case['grid']
    comm = MPI.COMM_WORLD
    comms = MPI_comm(comm, p_r, p_c)
    np.random.seed(case['seed'])
    A = C.draw_global(case, np.random)
    args = parse()
    args.size, args.rank, args.comm1, args.comm, args.p_r, args.p_c = size, rank, comms.comm, comms, p_r, p_c
    args.m, args.n, args.k = case['m'], case['n'], case['k']
    args.itr, args.init, args.verbose = case['itr'], 'rand', False
    args.row_comm, args.col_comm = comms.cart_1d_row(), comms.cart_1d_column()
    args.norm, args.method, args.prune = case['norm'], case['method'], case['prune']
    args.W_update = case['W_update']
    blk = determine_block_params(rank, (p_r, p_c), A.shape).determine_block_index_range_asymm()
    A_ij = A[blk[0][0]:blk[1][0] + 1, blk[0][1]:blk[1][1] + 1]
    factors = None
    if case['given_factors']:
        # shard sizes come from the reference itself (a dry data_operations pass)
        from pyDNMFk.utils import data_operations
        args.topo = '2d' if (p_r != 1 and p_c != 1) else '1d'
        dop = data_operations(A_ij, args)
        if args.topo == '2d':
            ml, nl = dop.params.m_loc, dop.params.n_loc
        else:
            ml, nl = A_ij.shape
        factors = list(C.draw_given_factors(case, np.random, ml, nl))
    W, H, err = PyNMF(A_ij, factors=factors, params=args).fit()
    out = dict(W=np.asarray(W), H=np.asarray(H), err=np.float64(err),
               geom=np.array([args.m, args.n, args.m_loc, args.n_loc, args.W_start, args.W_end,
                              args.H_start, args.H_end, blk[0][0], blk[1][0], blk[0][1], blk[1][1]],
                             dtype=np.int64))
    if case['prune']:
        out.update(row_zero_idx_x=np.asarray(args.row_zero_idx_x), col_zero_idx_x=np.asarray(args.col_zero_idx_x),
                   row_zero_idx_w=np.asarray(args.row_zero_idx_w), col_zero_idx_h=np.asarray(args.col_zero_idx_h))
    return out


def _ref_blocks(rank, size, shape, grid):
    from pyDNMFk.utils import determine_block_params
    d = determine_block_params(rank, grid, shape)
    s, e = d.determine_block_index_range_asymm()
    return np.array(list(s) + list(e) + list(d.determine_block_shape_asymm()), dtype=np.int64)


def _ref_sample(rank, size, seed, shape, nv, method):
    import numpy as np
    from pyDNMFk.pyDNMFk import sample
    X = (np.arange(shape[0] * shape[1], dtype=np.float32).reshape(shape) % 17) + 1
    Y = sample(data=X, noise_var=nv, method=method, seed=seed).fit()
    nxt = np.random.rand(3)     # the stream continues into PyNMF.init_factors
    return dict(Y=np.asarray(Y), nxt=nxt)


def main(argv):
    if not reference_available():
        raise SystemExit('reference not mounted; golden vectors can only be generated in the authoring container')
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, 'nmf_cases.npz')
    store = {}
    if argv and os.path.exists(path):
        with np.load(path) as z:
            store = {k: z[k] for k in z.files}
    todo = [C.CASES_BY_NAME[a] for a in argv] if argv else C.CASES
    for case in todo:
        P = case['grid'][0] * case['grid'][1]
        res = run_ranks(P, _ref_fit, (case,), timeout=600)
        for k in [k for k in store if k.startswith(case['name'] + '/')]:
            del store[k]
        for r, out in enumerate(res):
            for key, val in out.items():
                store['%s/%d/%s' % (case['name'], r, key)] = val
        print('%-46s P=%d err=%.6g' % (case['name'], P, float(res[0]['err'])), flush=True)
    np.savez_compressed(path, **store)
    print('wrote', path, os.path.getsize(path), 'bytes')

    if not argv:
        aux = {}
        for shape, grid in (((96, 21), (2, 1)), ((96, 21), (1, 2)), ((1024, 256), (4, 1)), ((26, 14), (2, 2)),
                            ((26, 14), (3, 2)), ((65536, 65536), (8, 1)), ((262144, 131072), (4, 2)), ((7, 5), (1, 1))):
            P = grid[0] * grid[1]
            res = run_ranks(P, _ref_blocks, (shape, grid))
            aux['blocks/%dx%d/%dx%d' % (shape + grid)] = np.stack(res)
        for seed, method in ((0, 'uniform'), (1000, 'uniform'), (3000, 'poisson')):
            res = run_ranks(1, _ref_sample, (seed, (12, 7), 0.015, method))
            aux['sample/%s/%d/Y' % (method, seed)] = res[0]['Y']
            aux['sample/%s/%d/nxt' % (method, seed)] = res[0]['nxt']
        np.savez_compressed(os.path.join(GOLDEN, 'aux_cases.npz'), **aux)
        print('wrote aux_cases.npz')


if __name__ == '__main__':
    main(sys.argv[1:])
