"""The NMFk-level oracle (oracle/nmfk_oracle.py) against the unmodified reference's outputs
(tests/golden/nmfk_cases.npz, produced by oracle/gen_golden_nmfk.py) and the reference's own fixtures."""
import os
import random

import numpy as np
import pytest

from oracle import nmf_oracle as O
from oracle import nmfk_cases as K
from oracle import nmfk_oracle as NK
from tests import common as T


@pytest.fixture(scope='module')
def gold():
    with np.load(os.path.join(K.GOLDEN, 'nmfk_cases.npz')) as z:
        return {k: z[k] for k in z.files}


def _tol(dtype):
    return 1e-11 if dtype == 'float64' else 2e-5


@pytest.mark.parametrize('case', K.CLUSTER_CASES, ids=[c['name'] for c in K.CLUSTER_CASES])
def test_clustering_matches_reference(gold, case):
    W_all, H_all = K.cluster_inputs(case)
    rows = K.row_split(case['m'], case['p_r'])
    cl = NK.Clustering([W_all[s:e] for s, e in rows], [H_all] * case['p_r'], case['p_r'], np.finfo(W_all.dtype).eps)
    res = cl.fit()
    sils = cl.silhouettes()
    tol = _tol(case['dtype'])
    for r in range(case['p_r']):
        g = lambda key: gold['cluster/%s/%d/%s' % (case['name'], r, key)]   # noqa: E731
        assert np.array_equal(np.asarray(res[r][5]), g('order'))
        assert T.rel_fro(res[r][0], g('centroids')) <= tol
        assert np.allclose(res[r][1], g('cent_std'), rtol=0, atol=tol)
        assert T.rel_fro(res[r][2], g('H_all')) <= tol
        assert T.rel_fro(cl.W[r], g('W_all')) <= tol
        assert np.allclose(res[r][3], g('sil_k'), rtol=0, atol=50 * tol)
        assert abs(res[r][4] - float(g('sil_avg'))) <= 50 * tol
        assert np.allclose(sils, g('sils'), rtol=0, atol=50 * tol)


def test_clustering_matches_reference_fixture(gold):
    """tests/test_dist_clustering.py:46-50 of the reference: silhouettes of its seeded example equal its sill.npy."""
    sil = np.load(os.path.join(K.GOLDEN, 'ref_sill.npy'))
    assert np.allclose(gold['cluster/reftest_2x1/0/sils'], sil, rtol=1e-3, atol=1e-3)
    case = K.CLUSTER_BY_NAME['reftest_2x1']
    W_all, H_all = K.cluster_inputs(case)
    cl = NK.Clustering([W_all[:8], W_all[8:]], [H_all, H_all], 2, np.finfo(np.float64).eps)
    cl.fit()
    assert np.allclose(cl.silhouettes(), sil, rtol=1e-3, atol=1e-3)


def _nnsvd_blocks(case, A):
    p_r, p_c = case['grid']
    return [A[s[0]:e[0] + 1, s[1]:e[1] + 1] for s, e in
            (O.block_range(r, (p_r, p_c), A.shape) for r in range(p_r * p_c))]


@pytest.mark.parametrize('case', K.NNSVD_CASES, ids=[c['name'] for c in K.NNSVD_CASES])
def test_nnsvd_matches_reference(gold, case):
    A = K.nnsvd_input(case)
    p_r, p_c = case['grid']
    rng = random.Random(K.NNSVD_PY_SEED)
    out = NK.nnsvd(_nnsvd_blocks(case, A), case['m'], case['n'], case['k'], p_r, p_c, np.finfo(A.dtype).eps, rng)
    for r, (W, H) in enumerate(out):
        gW, gH = gold['nnsvd/%s/%d/W' % (case['name'], r)], gold['nnsvd/%s/%d/H' % (case['name'], r)]
        assert W.shape == gW.shape and H.shape == gH.shape and W.dtype == gW.dtype
        assert T.rel_fro(W, gW) <= 1e-9 and T.rel_fro(H, gH) <= 1e-9, (T.rel_fro(W, gW), T.rel_fro(H, gH))


def test_nnsvd_matches_reference_fixtures(gold):
    """tests/test_dist_nnsvd.py:36-41,60-64 of the reference: W equals the stored sklearn factors at 1e-3."""
    f = np.load(os.path.join(K.GOLDEN, 'ref_nnsvd_factors_24x16.npz'))
    W = np.vstack([gold['nnsvd/tall24x16k2_2x1/%d/W' % r] for r in range(2)])
    assert np.allclose(W, f['W'], rtol=1e-3, atol=1e-3)
    f = np.load(os.path.join(K.GOLDEN, 'ref_nnsvd_factors_16x24.npz'))
    assert np.allclose(gold['nnsvd/short16x24k2_1x2/0/W'], f['W'], rtol=1e-3, atol=1e-3)
    assert gold['nnsvd/tall24x16k2_2x1/0/err_svd'] < 1e-15 and gold['nnsvd/tall24x16k2_2x1/0/err_nnsvd'] < .11


@pytest.mark.parametrize('case', K.NNSVD_FIT_CASES, ids=[c['name'] for c in K.NNSVD_FIT_CASES])
def test_fit_with_nnsvd_init_matches_reference(gold, case):
    A = K.nnsvd_fit_input(case)
    p_r, p_c = case['grid']
    blocks = _nnsvd_blocks(case, A)
    out = O.fit(blocks, p_r, p_c, case['k'], case['norm'], case['method'], case['itr'], init='nnsvd',
                py_rng=random.Random(K.NNSVD_PY_SEED), prune=True)
    for r, (W, H, err) in enumerate(out):
        g = lambda key: gold['nnsvdfit/%s/%d/%s' % (case['name'], r, key)]   # noqa: E731
        assert W.dtype == g('W').dtype
        assert T.rel_fro(W, g('W')) <= 1e-7 and T.rel_fro(H, g('H')) <= 1e-7
        assert abs(float(err) - float(g('err'))) <= 1e-7 * float(g('err'))


def test_rank_selection_matches_reference(gold):
    for name, sc in K.pvalue_scenarios().items():
        kr = range(sc['start_k'], sc['end_k'] + 1, sc['step_k'])
        nopt, pv = NK.pvalue_analysis(sc['L_err'], [round(s, 2) for s in sc['sil_min']], kr, sc['sill_thr'])
        assert nopt == int(gold['pvalue/%s/0/nopt' % name])
        assert np.allclose(pv, gold['pvalue/%s/0/pvalue' % name], rtol=1e-12, atol=0)
