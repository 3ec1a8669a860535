"""Module-level rank functions for tests/mp_util.run (they must be importable by spawned processes)."""
import os

import numpy as np


def comm_worker(rank, world, p_r, p_c):
    """Exercise every collective of pydnmfk_b200.dist_comm on CPU tensors / numpy / scalars."""
    import torch
    from pydnmfk_b200.dist_comm import MPI, MPI_comm
    comm = MPI.COMM_WORLD
    assert comm.size == world and comm.rank == rank
    comms = MPI_comm(comm, p_r, p_c)
    row, col = comms.cart_1d_row(), comms.cart_1d_column()
    i, j = divmod(rank, p_c)
    out = dict(coord=comms.coord2d, row_ranks=row.ranks, col_ranks=col.ranks, row_rank=row.rank, col_rank=col.rank)
    out['sum_int'] = comm.allreduce(rank + 1)
    out['sum_arr'] = comm.allreduce(np.full((2, 3), rank + 1.0, dtype=np.float32))
    out['row_sum'] = row.allreduce(np.array([10.0 * i + j]))
    out['col_sum'] = col.allreduce(np.array([10.0 * i + j]))
    out['gather'] = comm.allgather(('r', rank))
    out['bcast'] = comm.bcast({'from': rank} if rank == 0 else None, root=0)
    t = torch.full((4,), float(rank), dtype=torch.float64)
    out['t_allreduce'] = comm.allreduce_(t.clone()).numpy()
    out['t_gather'] = col.allgather_cat(torch.full((2, 3), float(rank))).numpy()
    sizes = [q + 1 for q in range(row.size)]
    out['t_gather_ragged'] = row.allgather_cat(torch.full((row.rank + 1, 2), float(rank)), sizes).numpy()
    full = torch.arange(float(row.size * 2 * 3)).reshape(row.size * 2, 3) * (rank + 1)
    out['t_rs'] = row.reduce_scatter_rows(full).numpy()
    rag = [1 + q for q in range(col.size)]
    full2 = torch.arange(float(sum(rag) * 2)).reshape(sum(rag), 2) * (rank + 1)
    out['t_rs_ragged'] = col.reduce_scatter_rows(full2, rag).numpy()
    b = torch.full((3,), float(rank + 5))
    out['t_bcast'] = comm.bcast_(b, root=0).numpy()
    send = np.arange(float(world * 3)) * (rank + 1)
    recv = np.empty(3)
    comm.Reduce_scatter(send, recv)
    out['Reduce_scatter'] = recv
    buf = np.full(4, float(rank))
    comm.Bcast(buf, root=world - 1)
    out['Bcast'] = buf
    comm.barrier()
    comms.Free()
    return out


def dims_worker(rank, world, case):
    """data_operations geometry + the RNG order of the rand initialisation, on the host."""
    from oracle import cases as C
    from pydnmfk_b200.dist_comm import MPI, MPI_comm
    from pydnmfk_b200.utils import parse, data_operations, determine_block_params
    from pydnmfk_b200.pyDNMF import draw_rand_factors
    p_r, p_c = case['grid']
    comm = MPI.COMM_WORLD
    comms = MPI_comm(comm, p_r, p_c)
    np.random.seed(case['seed'])
    A = C.draw_global(case, np.random)
    blk = determine_block_params(rank, (p_r, p_c), A.shape).determine_block_index_range_asymm()
    A_ij = A[blk[0][0]:blk[1][0] + 1, blk[0][1]:blk[1][1] + 1]
    p = parse()
    p.comm1, p.row_comm, p.col_comm, p.p_r, p.p_c, p.k = comm, comms.cart_1d_row(), comms.cart_1d_column(), p_r, p_c, case['k']
    p.topo = '2d' if (p_r != 1 and p_c != 1) else '1d'
    data_operations(A_ij, p)
    W, H = draw_rand_factors(p.topo, p_c, rank, A_ij.shape, (p.m_loc, p.n_loc), case['k'], A_ij.dtype)
    if p.topo == '1d':
        if p_c == 1:
            H = comm.bcast(H, root=0)
        else:
            W = comm.bcast(W, root=0)
    geom = [p.m, p.n, p.m_loc, p.n_loc, p.W_start, p.W_end, p.H_start, p.H_end, blk[0][0], blk[1][0], blk[0][1], blk[1][1]]
    return dict(geom=[int(v) for v in geom], W=W, H=H)


def fit_worker(rank, world, case, force_generic=False, resident=True):
    """One PyNMF.fit of a parity case on this rank (GPU; several ranks may share cuda:0 via gloo).  ``resident=False``
    keeps tiny single-rank fits on the per-kernel path instead of the whole-fit on-chip kernel."""
    import torch
    from oracle import cases as C
    from pydnmfk_b200 import _lib as L
    from pydnmfk_b200.dist_comm import MPI, MPI_comm
    from pydnmfk_b200.utils import parse, determine_block_params, data_operations
    from pydnmfk_b200.pyDNMF import PyNMF
    torch.cuda.set_device(0)
    L.set_force_generic(force_generic)
    p_r, p_c = case['grid']
    comm = MPI.COMM_WORLD
    comms = MPI_comm(comm, p_r, p_c)
    np.random.seed(case['seed'])
    A = C.draw_global(case, np.random)
    args = parse()
    args.size, args.rank, args.comm1, args.comm, args.p_r, args.p_c = world, rank, comms.comm, comms, p_r, p_c
    args.m, args.n, args.k = case['m'], case['n'], case['k']
    args.itr, args.init, args.verbose = case['itr'], 'rand', False
    args.row_comm, args.col_comm = comms.cart_1d_row(), comms.cart_1d_column()
    args.norm, args.method, args.prune = case['norm'], case['method'], case['prune']
    args.W_update = case['W_update']
    args.resident_fit = resident
    args.err_monitor = bool(case.get('err_monitor', False))
    blk = determine_block_params(rank, (p_r, p_c), A.shape).determine_block_index_range_asymm()
    A_ij = np.ascontiguousarray(A[blk[0][0]:blk[1][0] + 1, blk[0][1]:blk[1][1] + 1])
    factors = None
    if case['given_factors']:
        args.topo = '2d' if (p_r != 1 and p_c != 1) else '1d'
        dop = data_operations(A_ij, args)
        ml, nl = (dop.params.m_loc, dop.params.n_loc) if args.topo == '2d' else A_ij.shape
        factors = list(C.draw_given_factors(case, np.random, ml, nl))
    L.pass_count(True, reset=True)
    L.pass_count(False, reset=True)
    nmf = PyNMF(A_ij, factors=factors, params=args)
    W, H, err = nmf.fit()
    n_tc, n_generic = L.pass_count(True), L.pass_count(False)
    if case.get('expect_tc'):
        # the large cases exist to pin the tcgen05 kernels: every A-streaming pass must have taken that path
        # (or, in the forced-generic arm of the A/B test, none of them)
        if force_generic:
            assert n_tc == 0 and n_generic > 0, 'forced-generic arm ran %d tcgen05 passes' % n_tc
        else:
            assert n_generic == 0 and n_tc > 0, '%s: %d passes fell back to the generic kernels (tcgen05: %d)' % (
                case['name'], n_generic, n_tc)
    out = dict(W=np.asarray(W), H=np.asarray(H), err=float(err), err_dtype=str(np.asarray(err).dtype),
               tc_passes=int(n_tc), generic_passes=int(n_generic),
               peer_exchange=getattr(getattr(nmf, '_alg', None), '_px', None) is not None,
               err_history=getattr(nmf, 'err_history', None),
               geom=[int(v) for v in (args.m, args.n, args.m_loc, args.n_loc, args.W_start, args.W_end,
                                      args.H_start, args.H_end)])
    if case['prune']:
        out.update(row_zero_idx_x=np.asarray(args.row_zero_idx_x), col_zero_idx_x=np.asarray(args.col_zero_idx_x),
                   row_zero_idx_w=np.asarray(args.row_zero_idx_w), col_zero_idx_h=np.asarray(args.col_zero_idx_h))
    return out


def monitor_worker(rank, world, case, checkpoints):
    """PyNMF.fit with params.err_monitor on a parity case, plus plain fits cut short at `checkpoints` iterations: the
    trace-identity history must reproduce the direct residual (relative_err) of those shorter fits."""
    import copy
    out = {}
    c = copy.deepcopy(case)
    c['err_monitor'] = True
    full = fit_worker(rank, world, c)
    out['full'] = full
    for j in checkpoints:
        cj = copy.deepcopy(case)
        cj['itr'] = j
        cj['expect_tc'] = False
        out['err_at_%d' % j] = fit_worker(rank, world, cj)['err']
    return out


def fit_many_worker(rank, world, cases, force_generic=False):
    """Several parity cases (same world size, any grid) in one process group: amortises process spawn
    and CUDA context creation.  Per-case failures are captured, not raised."""
    import traceback
    out = {}
    for case in cases:
        try:
            out[case['name']] = ('ok', fit_worker(rank, world, case, force_generic))
            ok = 1
        except BaseException:
            out[case['name']] = ('err', traceback.format_exc())
            ok = 0
        from pydnmfk_b200.dist_comm import MPI
        if MPI.COMM_WORLD.allreduce(ok) != world:
            break   # stop the batch on every rank together
    return out


def ensemble_worker(rank, world, spread):
    """PyNMFk.fit_ensemble + fit_regression on wtsi-sized data: sequential, or spread over the ranks (replica mode)."""
    import torch
    from pydnmfk_b200.dist_comm import MPI, MPI_comm, Comm
    from pydnmfk_b200.pyDNMFk import PyNMFk
    from pydnmfk_b200.utils import parse
    torch.cuda.set_device(0)
    comm = MPI.COMM_WORLD
    solo = Comm([comm.ranks[comm.rank]], None)
    g = MPI_comm(solo, 1, 1)
    rs = np.random.RandomState(5)
    A = (rs.rand(96, 21) * 100).astype(np.float32)
    p = parse()
    p.comm1 = comm if spread else solo
    p.ensemble_parallel = bool(spread)
    p.comm, p.row_comm, p.col_comm = g, g.cart_1d_row(), g.cart_1d_column()
    p.p_r, p.p_c, p.init, p.verbose, p.itr, p.norm, p.method, p.prune = 1, 1, 'rand', False, 30, 'kl', 'mu', True
    p.perturbations, p.noise_var, p.sampling, p.checkpoint = 6, 0.015, 'uniform', False
    nmfk = PyNMFk(A, params=p)
    Wall, Hall, errs = nmfk.fit_ensemble(3)
    AvgW, AvgH = Wall[:, :, 0], np.median(Hall, axis=-1)
    Wr, Hr, er = nmfk.fit_regression(AvgW, AvgH)
    return dict(Wall=Wall, Hall=Hall, errs=[float(e) for e in errs], Wr=Wr, Hr=Hr, er=float(er),
                col_err=np.asarray(nmfk.col_err))


# ---- NMFk-level rows (clustering, nnsvd, end to end) ---------------------------------------------------------------
def _grid_params(rank, world, p_r, p_c):
    import torch
    from pydnmfk_b200.dist_comm import MPI, MPI_comm
    from pydnmfk_b200.utils import parse
    torch.cuda.set_device(0)
    comm = MPI.COMM_WORLD
    comms = MPI_comm(comm, p_r, p_c)
    p = parse()
    p.size, p.rank, p.comm1, p.comm, p.p_r, p.p_c = world, rank, comms.comm, comms, p_r, p_c
    p.row_comm, p.col_comm = comms.cart_1d_row(), comms.cart_1d_column()
    return p


def _guarded(fn, cases, *a):
    """Run fn(case) for a batch of cases in one process group; stop the batch on every rank together on a failure."""
    import traceback
    from pydnmfk_b200.dist_comm import MPI
    out = {}
    for case in cases:
        try:
            out[case['name']] = ('ok', fn(case, *a))
            ok = 1
        except BaseException:
            out[case['name']] = ('err', traceback.format_exc())
            ok = 0
        if MPI.COMM_WORLD.allreduce(ok) != MPI.COMM_WORLD.size:
            break
    return out


def cluster_worker(rank, world, cases):
    from oracle import nmfk_cases as K
    from pydnmfk_b200.dist_clustering import custom_clustering

    def one(case):
        W_all, H_all = K.cluster_inputs(case)
        s, e = K.row_split(case['m'], case['p_r'])[rank]
        p = _grid_params(rank, world, case['p_r'], 1)
        p.eps = np.finfo(W_all.dtype).eps
        cl = custom_clustering(np.ascontiguousarray(W_all[s:e]), H_all.copy(), p)
        centroids, cent_std, H_out, sil_k, sil_avg, order = cl.fit()
        return dict(centroids=centroids, cent_std=cent_std, H_all=H_out, W_all=cl.W_all, sil_k=sil_k,
                    sil_avg=float(sil_avg), order=np.asarray(order, dtype=np.int64), sils=cl.dist_silhouettes())
    return _guarded(one, cases)


def _block_of(A, rank, grid):
    from pydnmfk_b200.utils import determine_block_params
    b = determine_block_params(rank, grid, A.shape).determine_block_index_range_asymm()
    return np.ascontiguousarray(A[b[0][0]:b[1][0] + 1, b[0][1]:b[1][1] + 1])


def nnsvd_worker(rank, world, cases):
    import random
    from oracle import nmfk_cases as K
    from pydnmfk_b200.dist_svd import DistSVD

    def one(case):
        A = K.nnsvd_input(case)
        p_r, p_c = case['grid']
        p = _grid_params(rank, world, p_r, p_c)
        p.m, p.n, p.k = case['m'], case['n'], case['k']
        p.eps = np.finfo(A.dtype).eps
        random.seed(K.NNSVD_PY_SEED)
        (W, H), err = DistSVD(p, _block_of(A, rank, (p_r, p_c))).nnsvd(flag=1, verbose=1)
        return dict(W=W, H=H, err_svd=float(err['recon_err_svd']), err_nnsvd=float(err['recon_err_nnsvd']))
    return _guarded(one, cases)


def nnsvd_fit_worker(rank, world, cases):
    import random
    from oracle import nmfk_cases as K
    from pydnmfk_b200.pyDNMF import PyNMF

    def one(case):
        A = K.nnsvd_fit_input(case)
        p_r, p_c = case['grid']
        p = _grid_params(rank, world, p_r, p_c)
        p.m, p.n, p.k = case['m'], case['n'], case['k']
        p.itr, p.init, p.verbose, p.norm, p.method = case['itr'], 'nnsvd', False, case['norm'], case['method']
        random.seed(K.NNSVD_PY_SEED)
        W, H, err = PyNMF(_block_of(A, rank, (p_r, p_c)), factors=None, params=p).fit()
        return dict(W=np.asarray(W), H=np.asarray(H), err=float(err))
    return _guarded(one, cases)


def nmfk_e2e_worker(rank, world, case, tmp):
    """PyNMFk.fit() on the 96 x 21 example matrix; returns nopt and the per-k results this rank wrote."""
    import random
    from oracle import nmfk_cases as K
    from pydnmfk_b200.data_io import read_results
    from pydnmfk_b200.pyDNMFk import PyNMFk
    X = K.wtsi().astype('float32')
    p_r, p_c = case['grid']
    p = _grid_params(rank, world, p_r, p_c)
    p.fpath, p.fname, p.ftype = 'data/', 'wtsi', 'mat'
    p.init, p.itr, p.norm, p.method, p.verbose = case['init'], case['itr'], case['norm'], case['method'], False
    p.start_k, p.end_k, p.step_k, p.sill_thr = case['start_k'], case['end_k'], 1, case['sill_thr']
    p.perturbations, p.noise_var, p.sampling = case['perturbations'], case['noise_var'], 'uniform'
    p.results_path, p.checkpoint, p.precision = tmp + '/', False, 'float32'
    random.seed(K.NNSVD_PY_SEED)
    nopt = PyNMFk(_block_of(X, rank, (p_r, p_c)), factors=None, params=p).fit()
    out = dict(nopt=int(nopt))
    for k in range(case['start_k'], case['end_k'] + 1):
        d = '%s/wtsi/%d/' % (tmp, k)
        if rank == 0:
            for key, val in read_results(d).items():
                out['k%d/%s' % (k, key)] = np.asarray(val, dtype=np.float64)
        if p_r == 1 and p_c == 1:
            wname, hname = 'W_reg_factors/W_0.npy', 'H_reg_factors/H_0.npy'
        elif p_c == 1:
            wname, hname = 'W_reg_factors/W_%d.npy' % rank, 'H_reg_factors/H.npy'
        else:
            wname, hname = 'W_reg_factors/W.npy', 'H_reg_factors/H_%d.npy' % rank
        out['k%d/W_reg' % k] = np.load(d + wname)
        out['k%d/H_reg' % k] = np.load(d + hname)
    return out


def nmfk_replica_worker(rank, world, case, tmp):
    """PyNMFk.fit() with the perturbation ensemble spread over replica groups: the world is larger than the case's
    p_r x p_c grid, every group of p_r * p_c ranks holds the whole matrix (params.ensemble_parallel)."""
    import random
    import torch
    from oracle import nmfk_cases as K
    from pydnmfk_b200.data_io import read_results
    from pydnmfk_b200.dist_comm import MPI
    from pydnmfk_b200.pyDNMFk import PyNMFk
    from pydnmfk_b200.utils import parse
    torch.cuda.set_device(int(os.environ.get('DNMF_TEST_DEVICE', '0')))
    X = K.wtsi().astype('float32')
    p_r, p_c = case['grid']
    pos = rank % (p_r * p_c)
    p = parse()
    p.comm1 = MPI.COMM_WORLD
    p.size, p.rank, p.p_r, p.p_c = world, rank, p_r, p_c
    p.comm = p.row_comm = p.col_comm = None          # set per replica group by PyNMFk
    p.ensemble_parallel = True
    p.fpath, p.fname, p.ftype = 'data/', 'wtsi', 'mat'
    p.init, p.itr, p.norm, p.method, p.verbose = case['init'], case['itr'], case['norm'], case['method'], False
    p.start_k, p.end_k, p.step_k, p.sill_thr = case['start_k'], case['end_k'], 1, case['sill_thr']
    p.perturbations, p.noise_var, p.sampling = case['perturbations'], case['noise_var'], 'uniform'
    p.results_path, p.checkpoint, p.precision = tmp + '/', False, 'float32'
    random.seed(K.NNSVD_PY_SEED)
    nopt = PyNMFk(_block_of(X, pos, (p_r, p_c)), factors=None, params=p).fit()
    out = dict(nopt=int(nopt), pos=pos)
    for k in range(case['start_k'], case['end_k'] + 1):
        d = '%s/wtsi/%d/' % (tmp, k)
        if rank == 0:
            for key, val in read_results(d).items():
                out['k%d/%s' % (k, key)] = np.asarray(val, dtype=np.float64)
    return out


def reference_suite_worker(rank, world):
    """The reference's own MPI tests (tests/test_dist_nmf_1d.py, test_dist_nmf_2d.py, test_dist_nmf_1d_nnsvd_init.py,
    test_dist_utils.py), restated against the `pyDNMFk` import alias with the reference's thresholds, on 2 ranks."""
    import random
    import torch
    torch.cuda.set_device(0)
    ns = {}
    exec('import pyDNMFk.config as config\nconfig.init(0)\nfrom pyDNMFk.pyDNMF import *\nfrom pyDNMFk.dist_comm import *', ns)
    np_, MPI, MPI_comm, parse, PyNMF = ns['np'], ns['MPI'], ns['MPI_comm'], ns['parse'], ns['PyNMF']
    determine_block_params, data_operations = ns['determine_block_params'], ns['data_operations']
    comm = MPI.COMM_WORLD
    out = {}

    def run_grid(A, grid, k, itr, init, combos):
        p_r, p_c = grid
        comms = MPI_comm(comm, p_r, p_c)
        args = parse()
        args.size, args.rank, args.comm1, args.comm, args.p_r, args.p_c = comm.size, comm.rank, comms.comm, comms, p_r, p_c
        args.m, args.n, args.k = A.shape[0], A.shape[1], k
        args.itr, args.init = itr, init
        args.row_comm, args.col_comm = comms.cart_1d_row(), comms.cart_1d_column()
        args.verbose = False
        b = determine_block_params(comm.rank, (p_r, p_c), A.shape).determine_block_index_range_asymm()
        A_ij = np_.ascontiguousarray(A[b[0][0]:b[1][0] + 1, b[0][1]:b[1][1] + 1])
        errs = {}
        for mthd, norm in combos:
            args.method, args.norm = mthd, norm
            W_ij, H_ij, rel_error = PyNMF(A_ij, factors=None, params=args).fit()
            errs['%s_%s' % (norm, mthd)] = float(rel_error)
        return errs

    combos = [('mu', 'fro'), ('mu', 'kl'), ('bcd', 'fro'), ('hals', 'fro')]
    # tests/test_dist_nmf_1d.py:11-39 (threshold 1e-3) and test_dist_nmf_2d.py (same recipe, threshold 1e-4 on [2, 1])
    np_.random.seed(100)
    m, k, n = 24, 2, 12
    A = np_.random.rand(m, k) @ np_.random.rand(k, n)
    for grid in ([1, 2], [2, 1]):
        out['nmf_%dx%d' % tuple(grid)] = run_grid(A, grid, k, 2000, 'rand', combos)
    # tests/test_dist_nmf_2d.py:11-40: the same recipe, freshly seeded, grid [2, 1] only, threshold 1e-4
    np_.random.seed(100)
    A = np_.random.rand(m, k) @ np_.random.rand(k, n)
    out['nmf2d_2x1'] = run_grid(A, [2, 1], k, 2000, 'rand', combos)
    # tests/test_dist_nmf_1d_nnsvd_init.py:12-40 (threshold 1e-1)
    np_.random.seed(100)
    random.seed(7)
    for grid, (mm, nn) in zip([[2, 1], [1, 2]], [(24, 12), (12, 24)]):
        A = np_.random.rand(mm, 2) @ np_.random.rand(2, nn)
        out['nnsvd_%dx%d' % tuple(grid)] = run_grid(A, grid, 2, 2000, 'nnsvd', combos)
    # tests/test_dist_utils.py:10-51: prune / un-prune keeps the factor shapes (wtsi with 3 zero rows / columns added)
    from oracle import nmfk_cases as K
    for grid in ([1, 2], [2, 1]):
        A = K.wtsi().astype(np_.float64)
        z_c = np_.zeros((A.shape[0], 1))
        A = np_.hstack((z_c, A[:, :A.shape[1] // 2], z_c, A[:, A.shape[1] // 2:], z_c))
        z_r = np_.zeros((1, A.shape[1]))
        A = np_.vstack((z_r, A[:A.shape[0] // 2, :], z_r, A[A.shape[0] // 2:, :], z_r))
        args = parse()
        comms = MPI_comm(comm, grid[0], grid[1])
        args.topo = '1d'
        args.size, args.rank, args.comm, args.p_r, args.p_c = comms.size, comms.rank, comms, grid[0], grid[1]
        args.row_comm, args.col_comm, args.comm1 = comms.cart_1d_row(), comms.cart_1d_column(), comms.comm
        args.k = 4
        b = determine_block_params(comms.rank, (grid[0], grid[1]), A.shape).determine_block_index_range_asymm()
        A_ij = np_.ascontiguousarray(A[b[0][0]:b[1][0] + 1, b[0][1]:b[1][1] + 1])
        data_op = data_operations(A_ij, args)
        data_op.zero_idx_prune()
        W_i_ = np_.random.rand(data_op.params.m_loc, args.k)
        H_j_ = np_.random.rand(args.k, data_op.params.n_loc)
        A_p, W_i, H_j = data_op.prune_all(W_i_, H_j_)
        W_u, H_u = data_op.unprune_factors(W_i, H_j)
        out['prune_%dx%d' % tuple(grid)] = dict(orig=(W_i_.shape, H_j_.shape), pruned=(tuple(A_p.shape), tuple(W_i.shape), tuple(H_j.shape)),
                                                unpruned=(tuple(W_u.shape), tuple(H_u.shape)))
    return out
