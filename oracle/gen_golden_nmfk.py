"""Golden vectors for the NMFk-level rows (clustering, silhouettes, nnsvd, rank selection, NMFk end to end),
produced by the UNMODIFIED reference on forked ranks.  TEST INFRASTRUCTURE; authoring container only.

    python oracle/gen_golden_nmfk.py [cluster] [nnsvd] [nnsvdfit] [pvalue] [e2e] [cfg5]

Writes tests/golden/nmfk_cases.npz, copies the reference's own small fixtures for this path (tests/sill.npy,
tests/nnsvd_factors_*.npy -> tests/golden/ref_*.npy/.npz) and the example matrices (data/wtsi.mat -> wtsi_X.npy, 96 x 21;
data/swim.mat -> swim_X.npy, 1024 x 256).
"""
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import nmfk_cases as K       # noqa: E402
from oracle.refrun.launch import run_ranks, reference_available, REFERENCE  # noqa: E402


def _args(p_r, p_c, size, rank):
    from pyDNMFk.dist_comm import MPI_comm
    from pyDNMFk.utils import parse
    from mpi4py import MPI
    comms = MPI_comm(MPI.COMM_WORLD, p_r, p_c)
    args = parse()
    args.size, args.rank, args.comm1, args.comm, args.p_r, args.p_c = size, rank, comms.comm, comms, p_r, p_c
    args.row_comm, args.col_comm = comms.cart_1d_row(), comms.cart_1d_column()
    return args


def _ref_cluster(rank, size, case):
    from pyDNMFk.dist_clustering import custom_clustering
    W_all, H_all = K.cluster_inputs(case)
    s, e = K.row_split(case['m'], case['p_r'])[rank]
    args = _args(case['p_r'], 1, size, rank)
    args.eps = np.finfo(W_all.dtype).eps
    cl = custom_clustering(W_all[s:e].copy(), H_all.copy(), args)
    centroids, cent_std, H_out, sil_k, sil_avg, order = cl.fit()
    return dict(centroids=centroids, cent_std=cent_std, H_all=H_out, W_all=cl.W_all, sil_k=sil_k,
                sil_avg=np.float64(sil_avg), order=np.asarray(order, dtype=np.int64),
                sils=cl.dist_silhouettes())          # third pass, as tests/test_dist_clustering.py:44-45 does


def _block(A, rank, grid):
    from pyDNMFk.utils import determine_block_params
    b = determine_block_params(rank, grid, A.shape).determine_block_index_range_asymm()
    return A[b[0][0]:b[1][0] + 1, b[0][1]:b[1][1] + 1]


def _ref_nnsvd(rank, size, case):
    import random
    from pyDNMFk.dist_svd import DistSVD
    A = K.nnsvd_input(case)
    p_r, p_c = case['grid']
    args = _args(p_r, p_c, size, rank)
    args.m, args.n, args.k = case['m'], case['n'], case['k']
    args.eps = np.finfo(A.dtype).eps
    random.seed(K.NNSVD_PY_SEED)
    (W, H), err = DistSVD(args, _block(A, rank, (p_r, p_c))).nnsvd(flag=1, verbose=1)
    return dict(W=np.asarray(W), H=np.asarray(H), err_svd=np.float64(err['recon_err_svd']),
                err_nnsvd=np.float64(err['recon_err_nnsvd']))


def _ref_nnsvd_fit(rank, size, case):
    import random
    from pyDNMFk.pyDNMF import PyNMF
    A = K.nnsvd_fit_input(case)
    p_r, p_c = case['grid']
    args = _args(p_r, p_c, size, rank)
    args.m, args.n, args.k = case['m'], case['n'], case['k']
    args.itr, args.init, args.verbose = case['itr'], 'nnsvd', False
    args.norm, args.method = case['norm'], case['method']
    random.seed(K.NNSVD_PY_SEED)
    W, H, err = PyNMF(_block(A, rank, (p_r, p_c)), factors=None, params=args).fit()
    return dict(W=np.asarray(W), H=np.asarray(H), err=np.float64(err))


def _ref_pvalue(rank, size, sc, tmp):
    from h5py import File
    from pyDNMFk.pyDNMFk import PyNMFk
    from pyDNMFk.utils import parse
    p = parse()
    p.start_k, p.end_k, p.results_path = sc['start_k'], sc['end_k'], tmp + '/'
    for i, k in enumerate(range(sc['start_k'], sc['end_k'] + 1, sc['step_k'])):
        os.makedirs('%s/%d' % (tmp, k), exist_ok=True)
        with File('%s/%d/results.h5' % (tmp, k), 'w') as hf:
            hf.create_dataset('L_err', data=sc['L_err'][i])
            hf.create_dataset('clusterSilhouetteCoefficients', data=np.array([sc['sil_min'][i], 1.0]))
    obj = object.__new__(PyNMFk)
    obj.params, obj.step_k, obj.sill_thr = p, sc['step_k'], sc['sill_thr']
    nopt, pvalue = obj.pvalueAnalysis()
    return dict(nopt=np.int64(nopt), pvalue=np.asarray(pvalue, dtype=np.float64))


def _ref_e2e(rank, size, case, tmp):
    import random
    from h5py import File
    from pyDNMFk.pyDNMFk import PyNMFk
    X = K.wtsi().astype('float32')
    p_r, p_c = case['grid']
    args = _args(p_r, p_c, size, rank)
    args.fpath, args.fname, args.ftype = 'data/', 'wtsi', 'mat'
    args.init, args.itr, args.norm, args.method, args.verbose = case['init'], case['itr'], case['norm'], case['method'], False
    args.start_k, args.end_k, args.step_k, args.sill_thr = case['start_k'], case['end_k'], 1, case['sill_thr']
    args.perturbations, args.noise_var, args.sampling = case['perturbations'], case['noise_var'], 'uniform'
    args.results_path = tmp + '/'
    args.checkpoint = False
    args.precision = 'float32'
    random.seed(K.NNSVD_PY_SEED)
    A_ij = np.ascontiguousarray(_block(X, rank, (p_r, p_c)))
    nopt = PyNMFk(A_ij, factors=None, params=args).fit()
    out = dict(nopt=np.int64(nopt))
    for k in range(case['start_k'], case['end_k'] + 1):
        d = '%s/wtsi/%d/' % (tmp, k)
        if rank == 0:
            with File(d + 'results.h5', 'r') as hf:
                for key in ('clusterSilhouetteCoefficients', 'avgSilhouetteCoefficients', 'L_err', 'L_errDist', 'avgErr',
                            'ErrTol', 'AIC'):
                    out['k%d/%s' % (k, key)] = np.asarray(hf[key], dtype=np.float64)
        wname = 'W_reg_factors/W_%d.npy' % rank if (p_r != 1 or p_c == 1) else 'W_reg_factors/W.npy'
        if p_r == 1 and p_c == 1:
            wname, hname = 'W_reg_factors/W_0.npy', 'H_reg_factors/H_0.npy'
        elif p_c == 1:
            hname = 'H_reg_factors/H.npy'
        else:
            hname = 'H_reg_factors/H_%d.npy' % rank
        out['k%d/W_reg' % k] = np.load(d + wname)
        out['k%d/H_reg' % k] = np.load(d + hname)
    return out


def main(argv):
    if not reference_available():
        raise SystemExit('reference not mounted; golden vectors can only be generated in the authoring container')
    what = set(argv) or {'cluster', 'nnsvd', 'nnsvdfit', 'pvalue', 'e2e', 'fixtures'}
    os.makedirs(K.GOLDEN, exist_ok=True)
    path = os.path.join(K.GOLDEN, 'nmfk_cases.npz')
    store = {}
    if os.path.exists(path):
        with np.load(path) as z:
            store = {k: z[k] for k in z.files}

    def put(prefix, res):
        for k in [k for k in store if k.startswith(prefix + '/')]:
            del store[k]
        for r, out in enumerate(res):
            for key, val in out.items():
                store['%s/%d/%s' % (prefix, r, key)] = val

    if 'fixtures' in what:
        from scipy.io import loadmat
        np.save(os.path.join(K.GOLDEN, 'wtsi_X.npy'), loadmat(os.path.join(REFERENCE, 'data', 'wtsi.mat'))['X'])
        np.save(os.path.join(K.GOLDEN, 'swim_X.npy'), loadmat(os.path.join(REFERENCE, 'data', 'swim.mat'))['X'])
        shutil.copy(os.path.join(REFERENCE, 'tests', 'sill.npy'), os.path.join(K.GOLDEN, 'ref_sill.npy'))
        for t in ('24x16', '16x24'):
            f = np.load(os.path.join(REFERENCE, 'tests', 'nnsvd_factors_%s.npy' % t), allow_pickle=True).item()
            np.savez(os.path.join(K.GOLDEN, 'ref_nnsvd_factors_%s.npz' % t), W=f['W'], H=f['H'])
    if 'cluster' in what:
        for case in K.CLUSTER_CASES:
            res = run_ranks(case['p_r'], _ref_cluster, (case,), timeout=600)
            put('cluster/' + case['name'], res)
            print('cluster %-26s sil_avg=%.6f' % (case['name'], float(res[0]['sil_avg'])), flush=True)
    if 'nnsvd' in what:
        for case in K.NNSVD_CASES:
            res = run_ranks(case['grid'][0] * case['grid'][1], _ref_nnsvd, (case,), timeout=600)
            put('nnsvd/' + case['name'], res)
            print('nnsvd %-26s err_svd=%.3g err_nnsvd=%.4f' % (case['name'], res[0]['err_svd'], res[0]['err_nnsvd']), flush=True)
    if 'nnsvdfit' in what:
        for case in K.NNSVD_FIT_CASES:
            res = run_ranks(case['grid'][0] * case['grid'][1], _ref_nnsvd_fit, (case,), timeout=600)
            put('nnsvdfit/' + case['name'], res)
            print('nnsvdfit %-30s err=%.6g dtype=%s' % (case['name'], res[0]['err'], res[0]['W'].dtype), flush=True)
    if 'pvalue' in what:
        for name, sc in K.pvalue_scenarios().items():
            with tempfile.TemporaryDirectory() as tmp:
                res = run_ranks(1, _ref_pvalue, (sc, tmp), timeout=120)
            put('pvalue/' + name, res)
            print('pvalue %-20s nopt=%d p=%s' % (name, res[0]['nopt'], np.round(res[0]['pvalue'], 5)), flush=True)
    if 'e2e' in what:
        for case in K.E2E_CASES:
            with tempfile.TemporaryDirectory() as tmp:
                res = run_ranks(case['grid'][0] * case['grid'][1], _ref_e2e, (case, tmp), timeout=3000)
            put('e2e/' + case['name'], res)
            print('e2e %-20s nopt=%d' % (case['name'], res[0]['nopt']), flush=True)
    if 'cfg5' in argv:
        case = K.CFG5_CASE
        with tempfile.TemporaryDirectory() as tmp:
            res = run_ranks(case['grid'][0] * case['grid'][1], _ref_e2e, (case, tmp), timeout=6000)
        out = {}
        for r, o in enumerate(res):
            for key, val in o.items():
                out['e2e/%s/%d/%s' % (case['name'], r, key)] = val
        p5 = os.path.join(K.GOLDEN, 'nmfk_cfg5.npz')
        np.savez_compressed(p5, **out)
        print('cfg5 %-20s nopt=%d' % (case['name'], res[0]['nopt']), 'wrote', p5, os.path.getsize(p5), 'bytes', flush=True)
        if what == {'cfg5'}:
            return
    np.savez_compressed(path, **store)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main(sys.argv[1:])
