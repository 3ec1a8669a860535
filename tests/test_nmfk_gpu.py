"""NMFk-level rows on the GPU: clustering + silhouettes, nnsvd initialisation, rank selection
and PyNMFk.fit end to end, against the unmodified reference's outputs (tests/golden/nmfk_cases.npz), the reference's
own fixtures and the numpy oracle.  Multi-rank cases run one process per rank on cuda:0 over gloo.

Tolerances: assignments / permutation orders bit-exact; clustered factors 1e-5 (fp32) / 1e-10 (fp64) relative Frobenius;
silhouettes 1e-4 absolute.  nnsvd factors 1e-3: the power iteration stops at |<v_new, v_old>| > 1 - eps_fp32
(dist_svd.py:126), i.e. the reference's own factors are only determined to ~1e-3 (its test uses rtol 1e-3 as well).
"""
import os
import tempfile

import numpy as np
import pytest
import torch

from oracle import nmfk_cases as K
from oracle import nmfk_oracle as NK
from tests import common as T
from tests import mp_util, workers

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def gold():
    with np.load(os.path.join(K.GOLDEN, 'nmfk_cases.npz')) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope='module')
def ops():
    from pydnmfk_b200 import device as D
    return D.default_ops()


def _dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


# ---- kernels ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_group_kernels(ops, dtype):
    rs = np.random.RandomState(1)
    m, k, P, n = 37, 5, 6, 9
    W = rs.rand(m, k, P).astype(dtype)
    H = rs.rand(k, n, P).astype(dtype)
    eps = float(np.finfo(dtype).eps)
    tol = 1e-6 if dtype == np.float32 else 1e-14
    nrm = ops.colsum_wide(_dev(W).view(m, k * P), squares=True)
    ref = (W.astype(np.float64) ** 2).sum(0)
    assert np.allclose(nrm.cpu().numpy().reshape(k, P), ref, rtol=10 * tol)
    assert np.allclose(ops.colsum_wide(_dev(W).view(m, k * P)).cpu().numpy().reshape(k, P), W.astype(np.float64).sum(0), rtol=10 * tol)
    s = rs.rand(k, P).astype(dtype) + 0.5
    for mode, f in ((0, lambda x, t: x * t), (1, lambda x, t: x / t), (2, lambda x, t: x / np.sqrt(t + dtype(eps))),
                    (3, lambda x, t: x * np.sqrt(t + dtype(eps))), (4, lambda x, t: x / (t + dtype(eps))),
                    (5, lambda x, t: x * (t + dtype(eps)))):
        Wd = ops.scale_groups(_dev(W), _dev(s), (0, P, 1), mode, eps).cpu().numpy()
        assert np.allclose(Wd, f(W, s.reshape(1, k, P)), rtol=4 * tol, atol=0), mode
        Hd = ops.scale_groups(_dev(H), _dev(s), (P, 0, 1), mode, eps).cpu().numpy()
        assert np.allclose(Hd, f(H, s.reshape(k, 1, P)), rtol=4 * tol, atol=0), mode
    order = np.stack([rs.permutation(k) for _ in range(P)]).astype(np.int32)
    Wp = ops.permute_groups(_dev(W), _dev(order), axis=1).cpu().numpy()
    Hp = ops.permute_groups(_dev(H), _dev(order), axis=0).cpu().numpy()
    Hs = ops.permute_groups(_dev(H), _dev(order), axis=0, sequential=True).cpu().numpy()
    Hseq = H.copy()
    for p in range(P):
        assert np.array_equal(Wp[:, :, p], W[:, :, p][:, order[p]])
        assert np.array_equal(Hp[:, :, p], H[:, :, p][order[p]])
        Hseq[:, :, p] = [Hseq[:, :, p][q] for q in order[p]]      # numpy's in-place list-of-views assignment
    assert np.array_equal(Hs, Hseq)
    for PP in (1, 2, 5, 6, 20, 33):
        X = rs.randn(11, 7, PP).astype(dtype)
        X[0, 0, :] = 1.0                                  # ties
        med, mad = ops.median_last(_dev(X), want_mad=True)
        assert np.array_equal(med.cpu().numpy(), np.median(X, axis=-1))
        assert np.array_equal(mad.cpu().numpy(), np.median(np.abs(X - np.median(X, axis=-1, keepdims=True)), axis=-1))


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
@pytest.mark.parametrize('k,P', [(1, 3), (2, 5), (3, 4), (8, 20), (17, 7), (64, 3)])
def test_greedy_assignment_is_bit_exact(ops, dtype, k, P):
    rs = np.random.RandomState(k * 100 + P)
    D3 = rs.rand(P, k, k).astype(dtype)
    D3[0] = np.round(D3[0] * 4) / 4                      # many exact ties: np.argmax takes the first in row-major order
    if P > 1:
        D3[1] = 0.5
    flat = np.ascontiguousarray(D3.transpose(1, 2, 0).reshape(k, k * P))      # [r, c*P + p]
    got = ops.greedy_lsa(_dev(flat), k, P).cpu().numpy()
    for p in range(P):
        assert [int(v) for v in got[p]] == NK.change_order(NK.greedy_lsa(D3[p])), p


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_silhouette_kernel(ops, dtype):
    rs = np.random.RandomState(4)
    for k, P in ((1, 4), (2, 3), (5, 6), (9, 20)):
        Wf = rs.rand(50, k * P)
        Wf /= np.linalg.norm(Wf, axis=0)
        G = (Wf.T @ Wf).astype(dtype)
        G[0, 0] = 1.0000001                               # clip
        d = np.arccos(np.clip(G, -1.0, 1.0)).reshape(k, P, k, P).astype(np.float64)
        if k == 1:
            ref = np.ones((k, P))
        else:
            a = np.array([[1 / (P - 1) * d[kk, n, kk].sum() for n in range(P)] for kk in range(k)])
            b = np.zeros((k, P))
            for kk in range(k):
                for n in range(P):
                    t = d[kk, n].sum(axis=1)
                    t[kk] = np.inf
                    b[kk, n] = 1 / P * t.min()
            ref = (b - a) / np.maximum(a, b)
        got = ops.silhouettes(_dev(G), k, P).cpu().numpy()
        assert np.allclose(got, ref, rtol=0, atol=2e-5 if dtype == np.float32 else 1e-12)


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_svd_helper_kernels(ops, dtype):
    rs = np.random.RandomState(9)
    m, n = 300, 70
    A = rs.rand(m, n).astype(dtype)
    u, v, sig = rs.randn(m), rs.randn(n), np.array([1.7])
    M = ops.rank1_sub(_dev(A.copy()), _dev(u), _dev(v), _dev(sig)).cpu().numpy()
    ref = A.copy()
    ref -= sig[0] * np.outer(u, v)
    assert np.array_equal(M, ref)
    x, xt = rs.randn(n), rs.randn(m)
    assert np.allclose(ops.matvec_f64(_dev(A), _dev(x)).cpu().numpy(), A @ x, rtol=1e-13, atol=1e-13)
    assert np.allclose(ops.matvec_f64(_dev(A), _dev(xt), trans=True).cpu().numpy(), A.T @ xt, rtol=1e-13, atol=1e-12)
    big = rs.rand(3000, 129).astype(dtype)
    xb = rs.randn(3000)
    assert np.allclose(ops.matvec_f64(_dev(big), _dev(xb), trans=True).cpu().numpy(), big.T @ xb, rtol=1e-12, atol=1e-11)
    y, last = rs.randn(5000), rs.randn(5000)
    r = torch.zeros(1, dtype=torch.float64, device='cuda')
    vn = ops.power_normalize(_dev(y), _dev(last), r).cpu().numpy()
    assert np.allclose(vn, y / np.linalg.norm(y), rtol=1e-14) and abs(r.item() - np.dot(y / np.linalg.norm(y), last)) < 1e-12
    dst = torch.zeros((4, m), dtype=torch.float64, device='cuda')
    ops.div_store(_dev(xt), _dev(np.array([4.0])), dst[2])
    assert np.array_equal(dst.cpu().numpy()[2], xt / 2.0) and float(dst[0].abs().sum() + dst[3].abs().sum()) == 0.0
    U = rs.randn(m, 4)
    pn = ops.posneg_colsumsq(_dev(U)).cpu().numpy()
    assert np.allclose(pn[0], (np.where(U > 0, U, 0) ** 2).sum(0), rtol=1e-13)
    assert np.allclose(pn[1], (np.where(U < 0, -U, 0) ** 2).sum(0), rtol=1e-13)
    coef = rs.rand(16) + 0.5
    pos = np.array([1, 0, 1, 0], dtype=np.int32)
    ref = np.where(pos.astype(bool), coef[0:4] * np.where(U > 0, U, 0) / coef[4:8], coef[8:12] * np.where(U < 0, -U, 0) / coef[12:16])
    assert np.array_equal(ops.nnsvd_pick(_dev(U), _dev(coef), _dev(pos)).cpu().numpy(), ref)
    assert np.array_equal(ops.nnsvd_pick(_dev(U), _dev(coef), _dev(pos), transpose_out=True).cpu().numpy(), ref.T)
    X = rs.rand(500, 150).astype(dtype)
    tol = 2e-6 if dtype == np.float32 else 1e-13
    assert T.rel_fro(ops.gram_wide(_dev(X)).cpu().numpy(), X.astype(np.float64).T @ X.astype(np.float64)) <= tol
    Xw, Yw = rs.rand(70, 900).astype(dtype), rs.rand(70, 900).astype(dtype)
    assert T.rel_fro(ops.outer_gram_wide(_dev(Xw), _dev(Yw)).cpu().numpy(), Xw.astype(np.float64) @ Yw.astype(np.float64).T) <= tol


# ---- clustering / silhouettes vs the reference -------------------------------------------------------------------
_batches = {}


def _batch(kind, worker, cases, world):
    key = (kind, world)
    if key not in _batches:
        todo = [c for c in cases if (c['p_r'] if 'p_r' in c else c['grid'][0] * c['grid'][1]) == world]
        _batches[key] = mp_util.run(world, worker, (todo,), backend='gloo', timeout=900)
    return _batches[key]


def _result(per_rank, name, world):
    out = []
    for r in range(world):
        if name not in per_rank[r]:
            pytest.fail('case did not run (an earlier case of this batch failed on rank %d)' % r)
        tag, val = per_rank[r][name]
        if tag == 'err':
            pytest.fail('rank %d raised:\n%s' % (r, val))
        out.append(val)
    return out


@pytest.mark.parametrize('case', K.CLUSTER_CASES, ids=[c['name'] for c in K.CLUSTER_CASES])
def test_clustering_matches_reference(gold, case):
    world = case['p_r']
    res = _result(_batch('cluster', workers.cluster_worker, K.CLUSTER_CASES, world), case['name'], world)
    tol = 1e-10 if case['dtype'] == 'float64' else 1e-5
    for r, o in enumerate(res):
        g = lambda key: gold['cluster/%s/%d/%s' % (case['name'], r, key)]   # noqa: E731
        assert np.array_equal(o['order'], g('order')), 'assignment orders differ'
        assert o['W_all'].dtype == g('W_all').dtype and o['W_all'].shape == g('W_all').shape
        assert T.rel_fro(o['W_all'], g('W_all')) <= tol and T.rel_fro(o['H_all'], g('H_all')) <= tol
        assert T.rel_fro(o['centroids'], g('centroids')) <= tol
        assert np.allclose(o['cent_std'], g('cent_std'), rtol=0, atol=tol)
        assert np.allclose(o['sils'], g('sils'), rtol=0, atol=1e-4) and np.allclose(o['sil_k'], g('sil_k'), rtol=0, atol=1e-4)
        assert abs(o['sil_avg'] - float(g('sil_avg'))) <= 1e-4
    if case['name'] == 'reftest_2x1':                      # the reference's own fixture (tests/test_dist_clustering.py:46-50)
        assert np.allclose(res[0]['sils'], np.load(os.path.join(K.GOLDEN, 'ref_sill.npy')), rtol=1e-3, atol=1e-3)


# ---- nnsvd vs the reference --------------------------------------------------------------------------------------
@pytest.mark.parametrize('case', K.NNSVD_CASES, ids=[c['name'] for c in K.NNSVD_CASES])
def test_nnsvd_matches_reference(gold, case):
    world = case['grid'][0] * case['grid'][1]
    res = _result(_batch('nnsvd', workers.nnsvd_worker, K.NNSVD_CASES, world), case['name'], world)
    for r, o in enumerate(res):
        gW, gH = gold['nnsvd/%s/%d/W' % (case['name'], r)], gold['nnsvd/%s/%d/H' % (case['name'], r)]
        assert o['W'].shape == gW.shape and o['H'].shape == gH.shape and o['W'].dtype == gW.dtype
        dW, dH = T.rel_fro(o['W'], gW), T.rel_fro(o['H'], gH)
        assert dW <= 1e-3 and dH <= 1e-3, (dW, dH)
        assert abs(o['err_nnsvd'] - float(gold['nnsvd/%s/%d/err_nnsvd' % (case['name'], r)])) <= 1e-3
    if case['name'] == 'tall24x16k2_2x1':                  # tests/test_dist_nnsvd.py:36-41
        f = np.load(os.path.join(K.GOLDEN, 'ref_nnsvd_factors_24x16.npz'))
        assert np.allclose(np.vstack([o['W'] for o in res]), f['W'], rtol=1e-3, atol=1e-3)
        assert res[0]['err_svd'] < 1e-14 and res[0]['err_nnsvd'] < .11
    if case['name'] == 'short16x24k2_1x2':                 # tests/test_dist_nnsvd.py:60-64
        f = np.load(os.path.join(K.GOLDEN, 'ref_nnsvd_factors_16x24.npz'))
        assert np.allclose(res[0]['W'], f['W'], rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize('case', K.NNSVD_FIT_CASES, ids=[c['name'] for c in K.NNSVD_FIT_CASES])
def test_fit_with_nnsvd_init_matches_reference(gold, case):
    world = case['grid'][0] * case['grid'][1]
    res = _result(_batch('nnsvdfit', workers.nnsvd_fit_worker, K.NNSVD_FIT_CASES, world), case['name'], world)
    for r, o in enumerate(res):
        g = lambda key: gold['nnsvdfit/%s/%d/%s' % (case['name'], r, key)]   # noqa: E731
        dW, dH = T.rel_fro(o['W'], g('W')), T.rel_fro(o['H'], g('H'))
        assert dW <= 2e-3 and dH <= 2e-3, (dW, dH)
        assert abs(o['err'] - float(g('err'))) <= 1e-3 * float(g('err'))


# ---- NMFk end to end -----------------------------------------------------------------------------------------------
@pytest.mark.parametrize('case', K.E2E_CASES, ids=[c['name'] for c in K.E2E_CASES])
def test_nmfk_end_to_end_matches_reference(gold, case):
    world = case['grid'][0] * case['grid'][1]
    with tempfile.TemporaryDirectory() as tmp:
        res = mp_util.run(world, workers.nmfk_e2e_worker, (case, tmp), backend='gloo', timeout=1500)
    loose = case['init'] == 'nnsvd'
    for r, o in enumerate(res):
        assert o['nopt'] == int(gold['e2e/%s/%d/nopt' % (case['name'], r)])
        for k in range(case['start_k'], case['end_k'] + 1):
            g = lambda key: gold['e2e/%s/%d/k%d/%s' % (case['name'], r, k, key)]   # noqa: E731
            if r == 0:
                assert np.allclose(o['k%d/clusterSilhouetteCoefficients' % k], g('clusterSilhouetteCoefficients'), rtol=0,
                                   atol=5e-3 if loose else 1e-3)
                assert np.allclose(o['k%d/ErrTol' % k], g('ErrTol'), rtol=2e-3 if loose else 1e-4)
                assert np.allclose(o['k%d/L_err' % k], g('L_err'), rtol=2e-2 if loose else 2e-3, atol=1e-5)
                assert abs(float(o['k%d/L_errDist' % k]) - float(g('L_errDist'))) <= (2e-3 if loose else 1e-4) * float(g('L_errDist'))
                assert abs(float(o['k%d/AIC' % k]) - float(g('AIC'))) <= 1e-3 * abs(float(g('AIC')))
            tolf = 2e-2 if loose else 2e-3
            assert T.rel_fro(o['k%d/W_reg' % k], g('W_reg')) <= tolf and T.rel_fro(o['k%d/H_reg' % k], g('H_reg')) <= tolf


@pytest.mark.parametrize('name,world', [('wtsi_1x1_rand', 2), ('wtsi_2x1_rand', 4)])
def test_nmfk_ensemble_over_replica_groups_matches_reference(gold, name, world):
    """PyNMFk.fit() end to end with params.ensemble_parallel: the world is cut into replica groups of p_r x p_c ranks,
    each taking every n_groups-th perturbation (BASELINE.json configs[4]: "perturbation ensemble batched over
    8 x B200").  Statistics and the selected rank must equal the sequential reference run."""
    case = K.E2E_BY_NAME[name]
    with tempfile.TemporaryDirectory() as tmp:
        res = mp_util.run(world, workers.nmfk_replica_worker, (case, tmp), backend='gloo', timeout=1500)
    for o in res:
        assert o['nopt'] == int(gold['e2e/%s/%d/nopt' % (name, o['pos'])])
    o = res[0]
    for k in range(case['start_k'], case['end_k'] + 1):
        g = lambda key: gold['e2e/%s/0/k%d/%s' % (name, k, key)]   # noqa: E731
        assert np.allclose(o['k%d/clusterSilhouetteCoefficients' % k], g('clusterSilhouetteCoefficients'), rtol=0, atol=1e-3)
        assert np.allclose(o['k%d/ErrTol' % k], g('ErrTol'), rtol=1e-4)
        assert np.allclose(o['k%d/L_err' % k], g('L_err'), rtol=2e-3, atol=1e-5)


@pytest.mark.parametrize('d', [1, 5, 21, 96, 300, 512])
@pytest.mark.parametrize('dtype', [np.float32, np.float64], ids=['f32', 'f64'])
def test_power_iterate_equals_the_stepwise_loop(d, dtype):
    """dnmf_power_iterate (the whole loop of dist_svd.py:117-134 in one launch) must return exactly the vector of the
    launch-per-step loop (dnmf_matvec_f64 + dnmf_power_normalize until |<v, v_prev>| > 1 - eps)."""
    from pydnmfk_b200 import device as D
    ops = D.default_ops()
    rs = np.random.RandomState(d)
    X = rs.rand(3 * d + 2, d)
    B = (X.T @ X).astype(dtype)                        # a Gram matrix, like svd1D's
    v0 = rs.randn(d)
    v0 /= np.linalg.norm(v0)
    Bd = _dev(B)
    for eps in (float(np.finfo(np.float32).eps), np.finfo(np.float32).eps, np.finfo(np.float64).eps * 1e4):
        thr = 1. - eps                                   # np.float32 for the float32 eps: numpy then compares in float32
        cur = _dev(v0)
        r = torch.zeros(1, dtype=torch.float64, device='cuda')
        steps = 0
        while True:
            cur = ops.power_normalize(ops.matvec_f64(Bd, cur), cur, r)
            steps += 1
            if abs(float(r.item())) > thr or steps > 100000:
                break
        got = ops.power_iterate(Bd, _dev(v0), thr)
        assert torch.equal(got, cur), (type(thr), steps, float((got - cur).abs().max()))
