"""Shared parity-case definitions (TEST INFRASTRUCTURE).

One case = one ``PyNMF(A_ij, params).fit()`` on a p_r x p_c grid.  The same
recipe is replayed three ways and must give the same factors:

  * the unmodified reference on forked ranks     (oracle/gen_golden.py -> tests/golden)
  * the numpy restatement                         (oracle/nmf_oracle.py)
  * the CUDA product                              (pydnmfk_b200, tests -m gpu)

RNG convention (mirrors the reference tests, e.g. tests/test_dist_nmf_1d.py:14):
every rank seeds its *process-global* legacy numpy stream with ``seed``,
generates the full global matrix from that stream, slices its block, optionally
re-seeds with ``reseed``, and then ``PyNMF.init_factors`` keeps drawing from the
same stream (pyDNMF.py:107-129).
"""
import numpy as np


def _case(name, m, n, k, grid, norm, method, itr, dtype='float32', data='uniform',
          seed=100, reseed=None, prune=False, W_update=True, given_factors=False, k0=None):
    return dict(name=name, m=m, n=n, k=k, grid=tuple(grid), norm=norm, method=method, itr=itr,
                dtype=dtype, data=data, seed=seed, reseed=reseed, prune=prune,
                W_update=W_update, given_factors=given_factors, k0=k0 or k, expect_tc=False)


def draw_global(case, rs):
    """Generate the global matrix from legacy stream ``rs`` (a RandomState or
    the ``np.random`` module itself)."""
    m, n, k0 = case['m'], case['n'], case['k0']
    kind = case['data']
    if kind == 'lowrank':          # tests/test_dist_nmf_1d.py:16-20
        W = rs.rand(m, k0)
        H = rs.rand(k0, n)
        A = W @ H
    elif kind == 'uniform':
        A = rs.rand(m, n)
    elif kind == 'zeros':          # uniform with exact-zero rows / columns (prune path)
        A = rs.rand(m, n)
        A[rs.permutation(m)[:max(1, m // 8)], :] = 0
        A[:, rs.permutation(n)[:max(1, n // 8)]] = 0
    elif kind == 'sparse':         # 60 % zeros + a few zero rows/cols
        A = rs.rand(m, n)
        A[rs.rand(m, n) < 0.6] = 0
        A[rs.permutation(m)[:2], :] = 0
        A[:, rs.permutation(n)[:2]] = 0
    elif kind == 'swim':           # the reference's data/swim.mat (1024 x 256 uint8 images), stored as a test fixture
        import os
        A = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'swim_X.npy'))
        assert A.shape == (m, n)
    else:
        raise ValueError(kind)
    A = A.astype(case['dtype'])
    if case['reseed'] is not None:
        rs.seed(case['reseed'])
    return A


def draw_given_factors(case, rs, m_loc, n_loc):
    """For W_update=False / given-factor cases: per-rank factors drawn after
    the data from the same stream."""
    W = rs.rand(m_loc, case['k'])
    H = rs.rand(case['k'], n_loc)
    return W, H


def _grid_cases():
    cs = []
    # mirror of the reference's own 1-D test (fp64, exact rank-2 data, no reseed)
    for g in ((1, 1), (2, 1), (1, 2)):
        for norm, method in (('fro', 'mu'), ('kl', 'mu'), ('fro', 'hals'), ('fro', 'bcd')):
            cs.append(_case('reftest_%dx%d_%s_%s' % (g + (norm, method)), 24, 12, 2, g, norm, method,
                            300, dtype='float64', data='lowrank', prune=True))
    # small fp32 / fp64 sweeps over grids, methods and iteration counts
    for g in ((1, 1), (2, 1), (1, 2), (2, 2), (4, 1), (4, 2)):
        for norm, method in (('fro', 'mu'), ('kl', 'mu'), ('fro', 'hals'), ('fro', 'bcd')):
            for itr in (1, 10, 100):
                if method == 'bcd' and itr == 100:
                    continue
                for dt in ('float32', 'float64'):
                    if dt == 'float64' and (itr != 10 or g in ((4, 1),)):
                        continue
                    cs.append(_case('u64x48k4_%dx%d_%s_%s_i%d_%s' % (g + (norm, method, itr, dt[-2:])),
                                    64, 48, 4, g, norm, method, itr, dtype=dt, reseed=7))
    # ragged (non-divisible) shapes, 1-D only plus one 2x2 (SURVEY probe 2)
    for g in ((2, 1), (1, 2), (3, 1), (2, 2)):
        for norm, method in (('fro', 'mu'), ('kl', 'mu')):
            cs.append(_case('ragged26x14k3_%dx%d_%s_%s' % (g + (norm, method)), 26, 14, 3, g, norm, method,
                            10, reseed=11))
    # the headline configuration in miniature: k = 32, square
    for g in ((1, 1), (4, 1), (2, 2)):
        for norm in ('fro', 'kl'):
            for itr in (10, 100):
                cs.append(_case('u512k32_%dx%d_%s_mu_i%d' % (g + (norm, itr)), 512, 512, 32, g, norm, 'mu',
                                itr, reseed=7))
    cs.append(_case('u512k32_1x1_fro_mu_i10_64', 512, 512, 32, (1, 1), 'fro', 'mu', 10, dtype='float64', reseed=7))
    cs.append(_case('u512k32_1x1_kl_mu_i10_64', 512, 512, 32, (1, 1), 'kl', 'mu', 10, dtype='float64', reseed=7))
    cs.append(_case('u256x384k16_1x1_fro_hals_i10', 256, 384, 16, (1, 1), 'fro', 'hals', 10, reseed=7))
    cs.append(_case('u256x384k16_2x1_fro_hals_i10', 256, 384, 16, (2, 1), 'fro', 'hals', 10, reseed=7))
    cs.append(_case('u256x384k16_1x1_fro_bcd_i10', 256, 384, 16, (1, 1), 'fro', 'bcd', 10, reseed=7))
    cs.append(_case('u256x384k64_1x1_fro_mu_i10', 256, 384, 64, (1, 1), 'fro', 'mu', 10, reseed=7))
    cs.append(_case('u256x384k64_4x2_fro_mu_i10', 256, 384, 64, (4, 2), 'fro', 'mu', 10, reseed=7))
    # the tcgen05 path end to end: shards of >= 2^20 elements per rank (csrc/dnmf_tc.cu tc_eligible), so that every
    # A-streaming pass of these fits runs tc_pass_kernel / tc_kl_kernel (asserted in tests/workers.py:fit_worker)
    big = [
        ('u2048', 2048, 2048, (1, 1), 'fro', 32, 100), ('u2048', 2048, 2048, (1, 1), 'fro', 10, 10),
        ('u2048', 2048, 2048, (1, 1), 'fro', 64, 10), ('u2048', 2048, 2048, (1, 1), 'kl', 32, 100),
        ('u2048', 2048, 2048, (1, 1), 'kl', 10, 10),
        ('u2048', 2048, 2048, (2, 1), 'fro', 32, 10), ('u2048', 2048, 2048, (2, 1), 'kl', 32, 10),
        ('u2048', 2048, 2048, (1, 2), 'fro', 32, 10), ('u2048', 2048, 2048, (1, 2), 'kl', 32, 10),
        ('u2048', 2048, 2048, (2, 2), 'fro', 32, 10), ('u2048', 2048, 2048, (2, 2), 'kl', 32, 10),
        ('u2048', 2048, 2048, (2, 2), 'fro', 64, 10),
        ('u4096x1024', 4096, 1024, (1, 1), 'fro', 32, 10), ('u4096x1024', 4096, 1024, (2, 1), 'kl', 32, 100),
        ('u4096x1024', 4096, 1024, (2, 2), 'fro', 64, 100), ('u4096x1024', 4096, 1024, (1, 1), 'kl', 10, 100),
        # KL with 32 < k <= 64: the 64-wide build of the fused tcgen05 kernel (dnmf_tc_kl64.cu)
        ('u2048', 2048, 2048, (1, 1), 'kl', 64, 10), ('u2048', 2048, 2048, (2, 1), 'kl', 48, 10),
        ('u4096x1024', 4096, 1024, (1, 1), 'kl', 64, 100),
    ]
    for tag, m, n, g, norm, k, itr in big:
        c = _case('%sk%d_%dx%d_%s_mu_i%d' % (tag, k, g[0], g[1], norm, itr), m, n, k, g, norm, 'mu', itr, reseed=7)
        c['expect_tc'] = True
        cs.append(c)
    # BCD / HALS on the tensor path (cfg4 in miniature: k = 16): the BCD iteration's fused A H^T + residual pass
    for g in ((1, 1), (2, 1), (2, 2)):
        for method in ('bcd', 'hals'):
            c = _case('u2048k16_%dx%d_fro_%s_i10' % (g + (method,)), 2048, 2048, 16, g, 'fro', method, 10, reseed=7)
            c['expect_tc'] = True
            cs.append(c)
    # BCD beyond 10 iterations in fp32: 100 accept / restore decisions on the device-resident control state
    cs.append(_case('u64x48k4_1x1_fro_bcd_i100_32', 64, 48, 4, (1, 1), 'fro', 'bcd', 100, reseed=7))
    cs.append(_case('u64x48k4_2x1_fro_bcd_i100_32', 64, 48, 4, (2, 1), 'fro', 'bcd', 100, reseed=7))
    c = _case('u2048k16_1x1_fro_bcd_i100', 2048, 2048, 16, (1, 1), 'fro', 'bcd', 100, reseed=7)
    c['expect_tc'] = True
    cs.append(c)
    # BASELINE.json configs[0] itself: data/swim.mat, FRO-MU, k = 4, rand init, 1000 iterations on a 4 x 1 grid (the
    # reference's `mpirun -n 4 python main.py --p_r=4 --p_c=1 --k=4 --fname=swim --init=rand --itr=1000 --norm=fro
    # --method=mu`, prune off as main.py:28 defaults), plus the same fit on one rank
    cs.append(_case('swim_4x1_fro_mu_i1000', 1024, 256, 4, (4, 1), 'fro', 'mu', 1000, data='swim', reseed=7))
    cs.append(_case('swim_1x1_fro_mu_i1000', 1024, 256, 4, (1, 1), 'fro', 'mu', 1000, data='swim', reseed=7))
    # prune path: exact-zero rows/cols, fp32 in -> float64 out (utils.py:195,198)
    for g in ((1, 1), (2, 1), (1, 2), (2, 2)):
        for norm in ('fro', 'kl'):
            cs.append(_case('zeros40x36k3_%dx%d_%s_mu_prune' % (g + (norm,)), 40, 36, 3, g, norm, 'mu', 10,
                            data='zeros', reseed=5, prune=True))
    cs.append(_case('sparse96x64k4_2x1_kl_mu_prune', 96, 64, 4, (2, 1), 'kl', 'mu', 10,
                    data='sparse', reseed=5, prune=True))
    # regression-style fit: factors given, W fixed (pyDNMFk.py:245-247)
    for g in ((1, 1), (2, 1), (2, 2)):
        for norm in ('fro', 'kl'):
            cs.append(_case('given64x48k4_%dx%d_%s_mu_Wfixed' % (g + (norm,)), 64, 48, 4, g, norm, 'mu', 10,
                            reseed=9, W_update=False, given_factors=True))
    return cs


CASES = _grid_cases()
CASES_BY_NAME = {c['name']: c for c in CASES}
assert len(CASES_BY_NAME) == len(CASES)
