"""Input recipes for the NMFk-level rows.  TEST INFRASTRUCTURE.

Shared by the golden generator (unmodified reference), the numpy oracle and the GPU parity tests, so all three see the
same seeded inputs.
"""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def _gauss(n, mean, std):
    return np.exp(-(np.linspace(1, n, n) - mean) ** 2 / std)


# ---- clustering (dist_clustering.py) -----------------------------------------------------------
CLUSTER_CASES = [
    # name, m, k, P, n, p_r, dtype
    dict(name='reftest_2x1', m=16, k=3, P=4, n=5, p_r=2, dtype='float64', recipe='reftest'),
    dict(name='reftest_1x1', m=16, k=3, P=4, n=5, p_r=1, dtype='float64', recipe='reftest'),
    dict(name='c40k5p6_1x1_64', m=40, k=5, P=6, n=7, p_r=1, dtype='float64', recipe='bumps'),
    dict(name='c40k5p6_2x1_64', m=40, k=5, P=6, n=7, p_r=2, dtype='float64', recipe='bumps'),
    dict(name='c40k5p6_4x1_32', m=40, k=5, P=6, n=7, p_r=4, dtype='float32', recipe='bumps'),
    dict(name='c96k8p20_1x1_32', m=96, k=8, P=20, n=21, p_r=1, dtype='float32', recipe='bumps'),
    dict(name='c33k2p5_2x1_64', m=33, k=2, P=5, n=4, p_r=2, dtype='float64', recipe='bumps'),
    dict(name='c24k1p4_1x1_64', m=24, k=1, P=4, n=3, p_r=1, dtype='float64', recipe='bumps'),
    dict(name='c64k6p7_2x1_64_noisy', m=64, k=6, P=7, n=9, p_r=2, dtype='float64', recipe='noisy'),
]
CLUSTER_BY_NAME = {c['name']: c for c in CLUSTER_CASES}


def cluster_inputs(case):
    """Global (W_all [m,k,P], H_all [k,n,P]); W is split by rows over p_r ranks, H is replicated (the reference's
    test layout, tests/test_dist_clustering.py:21-40)."""
    m, k, P, n = case['m'], case['k'], case['P'], case['n']
    rs = np.random.RandomState(100)
    if case['recipe'] == 'reftest':
        W = np.vstack([_gauss(m, 3, 3), _gauss(m, 8, 2), _gauss(m, 14, 3)]).T
        W_all = np.stack([W[:, rs.permutation(k)] + rs.rand(m, k) * .1 for _ in range(P)], axis=-1)
        H_all = rs.rand(k, n, P)
    else:
        centers = np.linspace(2, m - 2, k)
        W = np.vstack([_gauss(m, c, 2.0 + (i % 3)) for i, c in enumerate(centers)]).T
        amp = .1 if case['recipe'] == 'bumps' else .45
        W_all = np.stack([W[:, rs.permutation(k)] + rs.rand(m, k) * amp for _ in range(P)], axis=-1)
        H_all = rs.rand(k, n, P) + 0.1
    return W_all.astype(case['dtype']), H_all.astype(case['dtype'])


def row_split(m, p_r):
    """Row ranges of a p_r x 1 grid (utils.py:15-46)."""
    out = []
    for i in range(p_r):
        s = i * (m // p_r) + min(i, m % p_r)
        e = (i + 1) * (m // p_r) + min(i + 1, m % p_r)
        out.append((s, e))
    return out


# ---- nnsvd (dist_svd.py) ------------------------------------------------------------------------
NNSVD_CASES = [
    dict(name='tall24x16k2_2x1', m=24, n=16, k=2, grid=(2, 1), dtype='float64', recipe='reftest_tall'),
    dict(name='short16x24k2_1x2', m=16, n=24, k=2, grid=(1, 2), dtype='float64', recipe='reftest_short'),
    dict(name='tall96x21k4_2x1_32', m=96, n=21, k=4, grid=(2, 1), dtype='float32', recipe='uniform'),
    dict(name='tall96x21k4_4x1_64', m=96, n=21, k=4, grid=(4, 1), dtype='float64', recipe='uniform'),
    dict(name='short21x96k3_1x2_32', m=21, n=96, k=3, grid=(1, 2), dtype='float32', recipe='uniform'),
    dict(name='square32k3_1x1_64', m=32, n=32, k=3, grid=(1, 1), dtype='float64', recipe='uniform'),
    dict(name='tall200x64k6_2x1_32', m=200, n=64, k=6, grid=(2, 1), dtype='float32', recipe='lowrank'),
]
NNSVD_BY_NAME = {c['name']: c for c in NNSVD_CASES}
NNSVD_PY_SEED = 4321            # `random.seed` before DistSVD draws its start vectors (dist_svd.py:80-85)


def nnsvd_input(case):
    m, n, k = case['m'], case['n'], case['k']
    rs = np.random.RandomState(0)
    if case['recipe'] == 'reftest_tall':            # tests/test_dist_nnsvd.py:14-21
        A = rs.rand(m, k) @ rs.rand(k, n)
    elif case['recipe'] == 'reftest_short':         # second draw of the same stream (tests/test_dist_nnsvd.py:44-48)
        rs.rand(24, 2), rs.rand(2, 16)
        A = rs.rand(m, k) @ rs.rand(k, n)
    elif case['recipe'] == 'lowrank':
        A = rs.rand(m, k) @ rs.rand(k, n) + 0.01 * rs.rand(m, n)
    else:
        A = rs.rand(m, n)
    return A.astype(case['dtype'])


# ---- PyNMF with init='nnsvd' (pyDNMF.py:131-135) ------------------------------------------------
NNSVD_FIT_CASES = [
    dict(name='fit24x12k2_2x1_fro_mu_64', m=24, n=12, k=2, grid=(2, 1), norm='fro', method='mu', itr=60, dtype='float64'),
    dict(name='fit12x24k2_1x2_kl_mu_64', m=12, n=24, k=2, grid=(1, 2), norm='kl', method='mu', itr=60, dtype='float64'),
    dict(name='fit96x21k4_2x1_kl_mu_32', m=96, n=21, k=4, grid=(2, 1), norm='kl', method='mu', itr=40, dtype='float32'),
    dict(name='fit96x21k3_2x1_fro_hals_32', m=96, n=21, k=3, grid=(2, 1), norm='fro', method='hals', itr=20, dtype='float32'),
]
NNSVD_FIT_BY_NAME = {c['name']: c for c in NNSVD_FIT_CASES}


def nnsvd_fit_input(case):
    rs = np.random.RandomState(100)
    m, n, k = case['m'], case['n'], case['k']
    if case['dtype'] == 'float64':
        A = rs.rand(m, k) @ rs.rand(k, n)           # tests/test_dist_nmf_1d_nnsvd_init.py:27-29
    else:
        A = rs.rand(m, n)
    return A.astype(case['dtype'])


# ---- rank selection (pyDNMFk.py:261-299) -----------------------------------------------------------
def pvalue_scenarios():
    """Fabricated per-k regression-error vectors and minimum silhouettes."""
    rs = np.random.RandomState(11)
    out = {}
    n = 21
    base = rs.rand(n) + 1.0
    out['drop_until_4'] = dict(start_k=2, end_k=7, step_k=1, sill_thr=0.9,
                               L_err=[base, base * 0.6, base * 0.3, base * 0.29 + 0.01 * rs.rand(n), base * 0.28, base * 0.1],
                               sil_min=[0.99, 0.97, 0.95, 0.5, 0.3, 0.2])
    out['never_significant'] = dict(start_k=1, end_k=4, step_k=1, sill_thr=0.9,
                                    L_err=[base + 0.2 * (rs.rand(n) - 0.5) for _ in range(4)], sil_min=[1.0, 0.95, 0.93, 0.91])
    out['step2'] = dict(start_k=2, end_k=10, step_k=2, sill_thr=0.8,
                        L_err=[base / (i + 1) for i in range(5)], sil_min=[0.99, 0.85, 0.7, 0.9, 0.95])
    return out


# ---- NMFk end to end (pyDNMFk.py:169-258) ------------------------------------------------------------
E2E_CASES = [
    dict(name='wtsi_1x1_rand', grid=(1, 1), init='rand', start_k=2, end_k=5, perturbations=6, itr=250, noise_var=0.015,
         sill_thr=0.9, norm='kl', method='mu'),
    dict(name='wtsi_2x1_rand', grid=(2, 1), init='rand', start_k=2, end_k=4, perturbations=5, itr=200, noise_var=0.015,
         sill_thr=0.9, norm='kl', method='mu'),
    dict(name='wtsi_2x1_nnsvd', grid=(2, 1), init='nnsvd', start_k=2, end_k=4, perturbations=5, itr=200, noise_var=0.015,
         sill_thr=0.9, norm='kl', method='mu'),
]
E2E_BY_NAME = {c['name']: c for c in E2E_CASES}

# BASELINE.json configs[4] exactly as stated (examples/dist_pynmfk_1d_wtsi.py:26-44 expects nopt == 4): nnsvd init on the
# 2 x 1 grid it requires, k = 2..10, 20 perturbations, KL-MU, 1000 iterations.  ~5 min through the unmodified reference,
# so it has its own golden file (tests/golden/nmfk_cfg5.npz, `python oracle/gen_golden_nmfk.py cfg5`) and is checked by
# tools/bench_cfg5.py on the GPUs rather than by every test run.
CFG5_CASE = dict(name='wtsi_2x1_nnsvd_cfg5', grid=(2, 1), init='nnsvd', start_k=2, end_k=10, perturbations=20, itr=1000,
                 noise_var=0.015, sill_thr=0.9, norm='kl', method='mu')


def wtsi():
    """The 96 x 21 mutation-count matrix of the reference's NMFk example (data/wtsi.mat, key 'X'), stored as a fixture."""
    return np.load(os.path.join(GOLDEN, 'wtsi_X.npy'))
