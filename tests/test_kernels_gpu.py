"""Per-kernel parity of libdnmf (through the C-ABI) against float64 numpy on the same seeded inputs.

Tolerances: fp32 results are compared with the float64 answer at 2e-5 relative Frobenius (an fp32 GEMM
with n <= 8192 terms sits around 1e-6); fp64 at 1e-12; integer / index work bit-exact.
"""
import numpy as np
import pytest
import torch

from tests import common as T

pytestmark = pytest.mark.gpu

TOL = {np.float32: 2e-5, np.float64: 1e-12}
SHAPES = [(1, 1, 1), (7, 5, 2), (24, 12, 2), (96, 21, 4), (130, 257, 3), (256, 384, 16), (300, 1000, 10),
          (513, 2049, 32), (1024, 256, 4), (700, 900, 64), (2048, 2048, 32), (4100, 300, 7)]


@pytest.fixture(scope='module')
def ops():
    from pydnmfk_b200 import device as D
    return D.default_ops()


def _mk(shape, dt, seed, lo=0.0):
    rs = np.random.RandomState(seed)
    return (rs.rand(*shape) + lo).astype(dt)


def _dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


@pytest.mark.parametrize('dt', [np.float32, np.float64], ids=['f32', 'f64'])
@pytest.mark.parametrize('m,n,k', SHAPES)
def test_ah_and_wta(ops, m, n, k, dt):
    A, H, W = _mk((m, n), dt, 1), _mk((k, n), dt, 2), _mk((m, k), dt, 3)
    A64, H64, W64 = A.astype(np.float64), H.astype(np.float64), W.astype(np.float64)
    V = ops.ah(_dev(A), _dev(H)).cpu().numpy()
    assert V.dtype == dt and T.rel_fro(V, A64 @ H64.T) <= TOL[dt]
    Y = ops.wta(_dev(A), _dev(W)).cpu().numpy()
    assert T.rel_fro(Y, W64.T @ A64) <= TOL[dt]
    Yt = ops.wta(_dev(A), _dev(W), transposed_out=True).cpu().numpy()
    assert Yt.shape == (n, k) and T.rel_fro(Yt, (W64.T @ A64).T) <= TOL[dt]


@pytest.mark.parametrize('dt', [np.float32, np.float64], ids=['f32', 'f64'])
@pytest.mark.parametrize('m,n,k', SHAPES)
def test_kl_contractions(ops, m, n, k, dt):
    A, H, W = _mk((m, n), dt, 4), _mk((k, n), dt, 5, 0.1), _mk((m, k), dt, 6, 0.1)
    eps = float(np.finfo(dt).eps)
    A64, H64, W64 = A.astype(np.float64), H.astype(np.float64), W.astype(np.float64)
    U = A64 / (W64 @ H64 + eps)
    V = ops.kl_uht(_dev(A), _dev(W), _dev(H), eps).cpu().numpy()
    assert T.rel_fro(V, U @ H64.T) <= TOL[dt]
    Y = ops.kl_wtu(_dev(A), _dev(W), _dev(H), eps).cpu().numpy()
    assert T.rel_fro(Y, W64.T @ U) <= TOL[dt]
    Yt = ops.kl_wtu(_dev(A), _dev(W), _dev(H), eps, transposed_out=True).cpu().numpy()
    assert T.rel_fro(Yt, (W64.T @ U).T) <= TOL[dt]


@pytest.mark.parametrize('dt', [np.float32, np.float64], ids=['f32', 'f64'])
def test_kl_zero_entries_and_sparse_data(ops, dt):
    """A has exact zeros (the swim / pruned-data situation): U must be exactly 0 there, no NaN."""
    rs = np.random.RandomState(0)
    A = _mk((200, 150), dt, 7)
    A[rs.rand(200, 150) < 0.6] = 0
    W, H = _mk((200, 5), dt, 8), _mk((5, 150), dt, 9)
    eps = float(np.finfo(dt).eps)
    V = ops.kl_uht(_dev(A), _dev(W), _dev(H), eps).cpu().numpy()
    U = A.astype(np.float64) / (W.astype(np.float64) @ H.astype(np.float64) + eps)
    assert np.isfinite(V).all() and T.rel_fro(V, U @ H.astype(np.float64).T) <= TOL[dt]


@pytest.mark.parametrize('dt', [np.float32, np.float64], ids=['f32', 'f64'])
def test_strided_shards(ops, dt):
    """A is a window of a larger matrix (lda > n) and not 16-byte aligned: the scalar-load path."""
    big = _mk((70, 101), dt, 10)
    A = torch.from_numpy(big).cuda()[3:68, 1:98]           # 65 x 97 window, ld = 101
    H, W = _mk((6, 97), dt, 11), _mk((65, 6), dt, 12)
    A64 = big[3:68, 1:98].astype(np.float64)
    V = ops.ah(A, _dev(H)).cpu().numpy()
    Y = ops.wta(A, _dev(W)).cpu().numpy()
    assert T.rel_fro(V, A64 @ H.astype(np.float64).T) <= TOL[dt]
    assert T.rel_fro(Y, W.astype(np.float64).T @ A64) <= TOL[dt]


@pytest.mark.parametrize('dt', [np.float32, np.float64], ids=['f32', 'f64'])
@pytest.mark.parametrize('rows,k', [(1, 1), (33, 3), (1000, 4), (5000, 16), (65536, 32), (777, 64)])
def test_gram(ops, rows, k, dt):
    X = _mk((rows, k), dt, 13)
    G = ops.gram(_dev(X), trans=False).cpu().numpy()
    assert T.rel_fro(G, X.astype(np.float64).T @ X.astype(np.float64)) <= TOL[dt]
    Xt = np.ascontiguousarray(X.T)
    G2 = ops.gram(_dev(Xt), trans=True).cpu().numpy()
    assert T.rel_fro(G2, X.astype(np.float64).T @ X.astype(np.float64)) <= TOL[dt]


@pytest.mark.parametrize('dt', [np.float32, np.float64], ids=['f32', 'f64'])
@pytest.mark.parametrize('m,n,k', [(5, 7, 2), (300, 500, 4), (1000, 70, 32), (129, 4097, 64), (64, 48, 10)])
def test_updates(ops, m, n, k, dt):
    eps = float(np.finfo(dt).eps)
    W, H = _mk((m, k), dt, 14), _mk((k, n), dt, 15)
    V, Y = _mk((m, k), dt, 16), _mk((k, n), dt, 17)
    G = _mk((k, k), dt, 18)
    x = _mk((k,), dt, 19)
    f = np.float64
    # MU
    Wd = _dev(W); ops.mu_update_w(Wd, _dev(V), _dev(G), eps)
    assert T.rel_fro(Wd.cpu().numpy(), W.astype(f) * (V.astype(f) / (W.astype(f) @ G.astype(f) + eps))) <= TOL[dt]
    Hd = _dev(H); ops.mu_update_h(Hd, _dev(Y), _dev(G), eps)
    ref_h = H.astype(f) * (Y.astype(f) / (H.astype(f).T @ G.astype(f) + eps).T)
    assert T.rel_fro(Hd.cpu().numpy(), ref_h) <= TOL[dt]
    Hd = _dev(H); ops.mu_update_h(Hd, _dev(np.ascontiguousarray(Y.T)), _dev(G), eps, y_transposed=True)
    assert T.rel_fro(Hd.cpu().numpy(), ref_h) <= TOL[dt]
    # KL
    Wd = _dev(W); ops.kl_update_w(Wd, _dev(V), _dev(x), eps)
    assert T.rel_fro(Wd.cpu().numpy(), W.astype(f) * (V.astype(f) / (x.astype(f)[None, :] + eps))) <= TOL[dt]
    Hd = _dev(H); ops.kl_update_h(Hd, _dev(Y), _dev(x), eps)
    assert T.rel_fro(Hd.cpu().numpy(), H.astype(f) * (Y.astype(f) / (x.astype(f)[:, None] + eps))) <= TOL[dt]
    # clamp
    Z = (_mk((m, k), dt, 20) - 0.5).astype(dt)
    Zd = _dev(Z); ops.clamp_min(Zd, eps)
    assert np.array_equal(Zd.cpu().numpy(), np.maximum(Z, dt(eps)))
    # BCD projected-gradient steps
    L = 3.7
    Wd = _dev(np.zeros_like(W)); ops.bcd_pg_w(Wd, _dev(W), _dev(V), _dev(G), L)
    assert T.rel_fro(Wd.cpu().numpy(), np.maximum(0, W.astype(f) - (W.astype(f) @ G.astype(f) - V.astype(f)) / L)) <= 10 * TOL[dt]
    Hd = _dev(np.zeros_like(H)); ops.bcd_pg_h(Hd, _dev(H), _dev(Y), _dev(G), L)
    assert T.rel_fro(Hd.cpu().numpy(), np.maximum(0, H.astype(f) - (G.astype(f) @ H.astype(f) - Y.astype(f)) / L)) <= 10 * TOL[dt]
    # HALS H sweep (Gauss-Seidel over rows)
    Hd = _dev(H); ops.hals_h(Hd, _dev(Y), _dev(G), eps)
    Hr = H.astype(f).copy()
    for kk in range(k):
        Hr[kk, :] = np.maximum(Hr[kk, :] + Y.astype(f)[kk, :] - G.astype(f)[kk, :].dot(Hr), eps)
    assert T.rel_fro(Hd.cpu().numpy(), Hr) <= 50 * TOL[dt]
    # HALS W column + normalisation
    Wd = _dev(W)
    Wr = W.astype(f).copy()
    for kk in range(min(k, 3)):
        sq = ops.hals_w_col(Wd, _dev(V), _dev(G), kk, eps)
        t = Wr[:, kk] * G.astype(f)[kk, kk] + V.astype(f)[:, kk] - Wr.dot(G.astype(f)[:, kk])
        Wr[:, kk] = np.maximum(t, eps)
        assert abs(sq.item() - np.sum(Wr[:, kk] ** 2)) <= 20 * TOL[dt] * np.sum(Wr[:, kk] ** 2)
        ops.div_col(Wd, kk, sq)
        Wr[:, kk] /= np.sqrt(np.sum(Wr[:, kk] ** 2))
    assert T.rel_fro(Wd.cpu().numpy(), Wr) <= 50 * TOL[dt]
    # column scaling and axpby
    Wd = _dev(W); ops.div_cols(Wd, _dev(x + 1))
    assert T.rel_fro(Wd.cpu().numpy(), W.astype(f) / (x.astype(f) + 1)[None, :]) <= TOL[dt]
    out = torch.empty_like(Wd); ops.axpby(out, _dev(W), _dev(V), 1.25, -0.5)
    assert T.rel_fro(out.cpu().numpy(), 1.25 * W.astype(f) - 0.5 * V.astype(f)) <= TOL[dt]


@pytest.mark.parametrize('dt', [np.float32, np.float64], ids=['f32', 'f64'])
@pytest.mark.parametrize('m,n,k', [(3, 4, 2), (500, 300, 4), (4096, 32, 32), (32, 9000, 32), (1000, 1000, 64)])
def test_sums_normalize_residual(ops, m, n, k, dt):
    f = np.float64
    eps = float(np.finfo(dt).eps)
    A, W, H = _mk((m, n), dt, 21), _mk((m, k), dt, 22), _mk((k, n), dt, 23)
    assert T.rel_fro(ops.colsum(_dev(W)).cpu().numpy(), W.astype(f).sum(0)) <= TOL[dt]
    assert T.rel_fro(ops.rowsum(_dev(H)).cpu().numpy(), H.astype(f).sum(1)) <= TOL[dt]
    for X in (A, W, H):
        assert abs(ops.sqnorm(_dev(X)).item() - np.sum(X.astype(f) ** 2)) <= 1e-12 * np.sum(X.astype(f) ** 2)
    s = W.sum(0).astype(dt)
    Wd, Hd = _dev(W), _dev(H)
    ops.normalize(Wd, Hd, _dev(s), eps)
    assert T.rel_fro(Wd.cpu().numpy(), W.astype(f) / (s.astype(f) + eps)) <= TOL[dt]
    assert T.rel_fro(Hd.cpu().numpy(), H.astype(f) * s.astype(f)[:, None]) <= TOL[dt]
    r = ops.residual_sqnorm(_dev(A), _dev(W), _dev(H)).cpu().numpy()
    R = A.astype(f) - W.astype(f) @ H.astype(f)
    assert abs(r[0] - np.sum(R ** 2)) <= 10 * TOL[dt] * np.sum(R ** 2)
    assert abs(r[1] - np.sum(A.astype(f) ** 2)) <= 1e-12 * np.sum(A.astype(f) ** 2)
    num, den = ops.column_err(_dev(A), _dev(W), _dev(H))
    assert T.rel_fro(num.cpu().numpy(), np.sum(R ** 2, 0)) <= 10 * TOL[dt]
    assert T.rel_fro(den.cpu().numpy(), np.sum(A.astype(f) ** 2, 0)) <= 1e-12


@pytest.mark.parametrize('dt', [np.float32, np.float64], ids=['f32', 'f64'])
def test_shard_ops_bit_exact(ops, dt):
    rs = np.random.RandomState(5)
    A = _mk((257, 130), dt, 24)
    A[rs.rand(257, 130) < 0.7] = 0
    A[[0, 17, 256], :] = 0
    A[:, [3, 129]] = 0
    rows, cols = ops.nnz_counts(_dev(A))
    assert np.array_equal(rows.cpu().numpy(), np.sum(A != 0, 1)) and np.array_equal(cols.cpu().numpy(), np.sum(A != 0, 0))
    rmask, cmask = np.sum(A != 0, 1) > 0, np.sum(A != 0, 0) > 0
    ri, ci = torch.from_numpy(np.flatnonzero(rmask)).cuda(), torch.from_numpy(np.flatnonzero(cmask)).cuda()
    P = ops.compact(_dev(A), ri, ci).cpu().numpy()
    assert np.array_equal(P, A[np.ix_(rmask, cmask)])
    W = _mk((int(rmask.sum()), 4), dt, 25)
    B = np.zeros((257, 4)); B[rmask, :] = W
    assert np.array_equal(ops.scatter_rows(_dev(W), ri, 257).cpu().numpy(), B)
    H = _mk((4, int(cmask.sum())), dt, 26)
    Cc = np.zeros((4, 130)); Cc[:, cmask] = H
    assert np.array_equal(ops.scatter_cols(_dev(H), ci, 130).cpu().numpy(), Cc)
    # perturbation: bit-identical to numpy's separate ufunc roundings (pyDNMFk.py:42-44)
    U = rs.random_sample(A.shape).astype(dt)
    nv = 0.015
    M = 2 * nv * U + nv
    M = M + 1
    assert np.array_equal(ops.perturb_uniform(_dev(A + 1), _dev(U), nv).cpu().numpy(), np.multiply(A + 1, M))


def test_sample_matches_reference_golden():
    """sample(...).fit() reproduces the reference's perturbed matrix and leaves the RNG stream where the
    reference leaves it."""
    from pydnmfk_b200.pyDNMFk import sample
    aux = T.golden('aux_cases.npz')
    X = (np.arange(12 * 7, dtype=np.float32).reshape(12, 7) % 17) + 1
    for method, seed in (('uniform', 0), ('uniform', 1000), ('poisson', 3000)):
        Y = sample(data=X, noise_var=0.015, method=method, seed=seed).fit()
        assert np.array_equal(np.asarray(Y), aux['sample/%s/%d/Y' % (method, seed)])
        assert np.array_equal(np.random.rand(3), aux['sample/%s/%d/nxt' % (method, seed)])


def test_determinism(ops):
    """Two runs of the split-reduction kernels give bit-identical results."""
    A, H, W = _mk((3000, 5000), np.float32, 27), _mk((32, 5000), np.float32, 28), _mk((3000, 32), np.float32, 29)
    Ad, Hd, Wd = _dev(A), _dev(H), _dev(W)
    v1, v2 = ops.ah(Ad, Hd).cpu().numpy(), ops.ah(Ad, Hd).cpu().numpy()
    y1, y2 = ops.wta(Ad, Wd).cpu().numpy(), ops.wta(Ad, Wd).cpu().numpy()
    k1 = ops.kl_wtu(Ad, Wd, Hd, 1e-7).cpu().numpy()
    k2 = ops.kl_wtu(Ad, Wd, Hd, 1e-7).cpu().numpy()
    assert np.array_equal(v1, v2) and np.array_equal(y1, y2) and np.array_equal(k1, k2)


def test_errors(ops):
    from pydnmfk_b200 import _lib as L
    A = torch.zeros((8, 8), device='cuda')
    H = torch.zeros((65, 8), device='cuda')
    with pytest.raises(L.DnmfError, match='DNMF_MAX_K'):
        ops.ah(A, H)


# ---- whole-fit on-chip multiplicative updates (dnmf_mu_fit_resident) --------------------------------------------
@pytest.mark.parametrize('dtype', [np.float32, np.float64])
@pytest.mark.parametrize('norm', ['fro', 'kl'])
@pytest.mark.parametrize('m,n,k', [(96, 21, 4), (40, 64, 7), (130, 33, 1), (17, 200, 16), (64, 48, 64), (1024, 256, 4),
                                   (300, 100, 8), (2050, 64, 5), (515, 300, 32)])
def test_resident_fit_equals_per_kernel_loop(dtype, norm, m, n, k):
    """The on-chip loop follows the per-kernel update + clamp sequence (same math, different summation order)."""
    import torch
    from pydnmfk_b200 import device as D
    from pydnmfk_b200.dist_comm import MPI, MPI_comm
    from pydnmfk_b200.dist_nmf import nmf_algorithms_1D
    from pydnmfk_b200.utils import parse
    ops = D.default_ops()
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    if not ops.resident_fit_fits(m, n, k, norm, tdt):
        pytest.skip('does not fit in shared memory')
    MPI._reset()
    comm = MPI.COMM_WORLD
    comms = MPI_comm(comm, 1, 1)
    rs = np.random.RandomState(m + n + k)
    eps = float(np.finfo(dtype).eps)
    B = 3
    A = [rs.rand(m, n).astype(dtype) for _ in range(B)]
    A[0][rs.rand(m, n) < 0.3] = 0
    W0 = [rs.rand(m, k).astype(dtype) for _ in range(B)]
    H0 = [rs.rand(k, n).astype(dtype) for _ in range(B)]
    dev = lambda x: torch.from_numpy(x.copy()).cuda()     # noqa: E731
    for w_update in (True, False):
        itr = 23
        ref = []
        for b in range(B):
            p = parse()
            p.m, p.n, p.p_r, p.p_c, p.k, p.comm1, p.norm, p.method = m, n, 1, 1, k, comm, norm, 'mu'
            p.row_comm, p.col_comm = comms.cart_1d_row(), comms.cart_1d_column()
            p.eps, p.W_update, p.itr = np.finfo(dtype).eps, w_update, itr
            Ad, Wd, Hd = dev(A[b]), dev(W0[b]), dev(H0[b])
            alg = nmf_algorithms_1D(Ad, Wd, Hd, params=p)
            for i in range(itr):
                alg.update()
                if i % 10 == 0:
                    ops.clamp_min(Hd, eps)
                    ops.clamp_min(Wd, eps)
            ref.append((Wd.cpu().numpy(), Hd.cpu().numpy()))
        As, Ws, Hs = [dev(a) for a in A], [dev(w) for w in W0], [dev(h) for h in H0]
        ops.mu_fit_resident(As, Ws, Hs, norm, w_update, 0, 9, eps)      # split ranges: the clamp phase follows `it`
        ops.mu_fit_resident(As, Ws, Hs, norm, w_update, 9, itr, eps)
        torch.cuda.synchronize()
        tol = 2e-4 if dtype == np.float32 else 1e-11
        for b in range(B):
            dW, dH = T.rel_fro(Ws[b].cpu().numpy(), ref[b][0]), T.rel_fro(Hs[b].cpu().numpy(), ref[b][1])
            assert dW <= tol and dH <= tol, (w_update, b, dW, dH)
            if not w_update:
                assert np.array_equal(Ws[b].cpu().numpy()[W0[b] > eps], W0[b][W0[b] > eps])


def test_resident_fit_limits():
    import torch
    from pydnmfk_b200 import device as D
    ops = D.default_ops()
    assert ops.resident_fit_fits(96, 21, 10, 'kl', torch.float32) and ops.resident_fit_ctas(96, 21, 10, 'kl', torch.float32) == 1
    assert ops.resident_fit_ctas(1024, 256, 4, 'kl', torch.float32) == 16     # swim: a 16-CTA cluster, 64 rows per CTA
    assert ops.resident_fit_ctas(130, 33, 1, 'fro', torch.float64) == 2
    assert not ops.resident_fit_fits(8192, 4096, 4, 'kl', torch.float32)      # 128 MiB: the A-streaming path
    assert not ops.resident_fit_fits(96, 21, 65, 'fro', torch.float32)
    assert not ops.resident_fit_fits(96, 21, 4, 'l1', torch.float32)


@pytest.mark.parametrize('dtype', ['float32', 'float64'])
@pytest.mark.parametrize('m,k', [(1, 1), (130, 3), (4100, 7), (70000, 32), (2048, 64)])
def test_trace_terms(ops, m, k, dtype):
    """dnmf_trace_terms: <W, V> and <G1, G2> in float64, and the device-side slot counter that appends to a history."""
    rs = np.random.RandomState(5)
    W, V = rs.rand(m, k).astype(dtype), rs.rand(m, k).astype(dtype)
    G1, G2 = rs.rand(k, k).astype(dtype), rs.rand(k, k).astype(dtype)
    ref = np.array([np.sum(W.astype(np.float64) * V.astype(np.float64)), np.sum(G1.astype(np.float64) * G2.astype(np.float64))])
    dW, dV, dG1, dG2 = (torch.from_numpy(x).cuda() for x in (W, V, G1, G2))
    got = ops.trace_terms(dW, dV, dG1, dG2).cpu().numpy()[0]
    assert np.allclose(got, ref, rtol=1e-12, atol=0)
    hist = torch.zeros((3, 2), dtype=torch.float64, device='cuda')
    counter = torch.zeros(1, dtype=torch.int64, device='cuda')
    for _ in range(5):                                   # two samples more than the history holds
        ops.trace_terms(dW, dV, dG1, dG2, out=hist, slot_counter=counter)
    assert int(counter.item()) == 5
    assert np.allclose(hist.cpu().numpy(), np.tile(ref, (3, 1)), rtol=1e-12, atol=0)
