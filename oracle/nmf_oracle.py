"""CPU oracle: numpy restatement of pyDNMFk's distributed NMF update loop.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline legs may import this module; the product package
(``pydnmfk_b200``) never does and has no CPU fallback.

Parity status: PINNED.  ``oracle/gen_golden.py`` runs the *unmodified*
reference (``/root/reference``, numpy + a fork-based mpi4py stand-in) and
stores its outputs under ``tests/golden/``; ``tests/test_oracle_golden.py``
checks this restatement against those files (and, when ``/root/reference`` is
mounted, against live reference runs).

The reference is SPMD (one process per MPI rank).  Here the P ranks are
"virtual": every per-rank quantity is a Python list indexed by world rank and
every collective is a small function over those lists, evaluated in
communicator-rank order (what the stand-in does, and what pickle-based
``allreduce`` does up to floating-point association).

Each function cites the reference lines it restates (paths relative to
``/root/reference``).
"""
import numpy as np

# ---------------------------------------------------------------------------
# shard index maps                                    pyDNMFk/utils.py:15-46
# ---------------------------------------------------------------------------


def block_range(rank, pgrid, shape):
    """Inclusive [start, end] per dimension of ``rank``'s block of an array of
    ``shape`` split over ``pgrid`` (utils.py:36-41).  ``rank`` is forced to 0
    for a 1-element grid (utils.py:33)."""
    pgrid = tuple(int(g) for g in pgrid)
    if int(np.prod(pgrid)) <= 1:
        rank = 0
    chunk = np.unravel_index(rank, pgrid)
    start = [int(i * (n // g) + min(i, n % g)) for n, g, i in zip(shape, pgrid, chunk)]
    end = [int((i + 1) * (n // g) + min(i + 1, n % g) - 1) for n, g, i in zip(shape, pgrid, chunk)]
    return start, end


def block_shape(rank, pgrid, shape):
    """utils.py:43-46."""
    s, e = block_range(rank, pgrid, shape)
    return [b - a + 1 for a, b in zip(s, e)]


def split_matrix(A, p_r, p_c):
    """Per-rank blocks of a global matrix (data_io.py:81-83)."""
    blocks = []
    for r in range(p_r * p_c):
        s, e = block_range(r, (p_r, p_c), A.shape)
        blocks.append(np.ascontiguousarray(A[s[0]:e[0] + 1, s[1]:e[1] + 1]))
    return blocks


# ---------------------------------------------------------------------------
# virtual communicators                               pyDNMFk/dist_comm.py:16-56
# ---------------------------------------------------------------------------


class VGrid:
    """p_r x p_c Cartesian grid, row-major rank -> (i, j) (dist_comm.py:22,
    utils.py:38).  ``row`` groups = reference's cart_1d_row (keeps dim 0: the
    p_r ranks of one grid *column*, ordered by i); ``col`` groups =
    cart_1d_column (the p_c ranks of one grid *row*, ordered by j)."""

    def __init__(self, p_r, p_c):
        self.p_r, self.p_c = int(p_r), int(p_c)
        self.p = self.p_r * self.p_c
        self.coords = [divmod(r, self.p_c) for r in range(self.p)]
        self.world = [list(range(self.p))]
        self.row = [[i * self.p_c + j for i in range(self.p_r)] for j in range(self.p_c)]
        self.col = [[i * self.p_c + j for j in range(self.p_c)] for i in range(self.p_r)]

    def row_rank(self, r):
        return self.coords[r][0]

    def col_rank(self, r):
        return self.coords[r][1]


def _seq_sum(vals):
    acc = vals[0]
    for v in vals[1:]:
        acc = acc + v
    return acc


def allreduce(vals, groups):
    out = list(vals)
    for g in groups:
        s = _seq_sum([vals[r] for r in g])
        for r in g:
            out[r] = s.copy() if isinstance(s, np.ndarray) else s
    return out


def allgather(vals, groups):
    out = list(vals)
    for g in groups:
        lst = [vals[r] for r in g]
        for r in g:
            out[r] = lst
    return out


def reduce_scatter(send, recv_sizes, groups, dtype):
    """Flat element-wise sum; member q of a group receives the contiguous
    chunk of length recv_sizes[q] (dist_nmf.py:169,202,315,341)."""
    out = [None] * len(send)
    for g in groups:
        total = _seq_sum([np.ascontiguousarray(send[r]).ravel() for r in g])
        off = 0
        for r in g:
            out[r] = np.array(total[off:off + recv_sizes[r]], dtype=dtype)
            off += recv_sizes[r]
    return out


# ---------------------------------------------------------------------------
# dims, prune, unprune                               pyDNMFk/utils.py:49-217
# ---------------------------------------------------------------------------


class Shards:
    """Everything ``data_operations`` leaves on ``params`` (per rank)."""
    pass


def compute_dims(A, grid, k):
    """Global m, n (utils.py:73-93) and factor-shard geometry (utils.py:97-115)."""
    p_r, p_c, P = grid.p_r, grid.p_c, grid.p
    topo = '2d' if (p_r != 1 and p_c != 1) else '1d'
    loc = [a.shape for a in A]
    if p_r != 1 and p_c == 1:
        n = [loc[r][1] for r in range(P)]
        m = [sum(loc[q][0] for q in range(P))] * P
    elif p_c != 1 and p_r == 1:
        n = [sum(loc[q][1] for q in range(P))] * P
        m = [loc[r][0] for r in range(P)]
    else:
        mm = sum(loc[r][0] for r in range(P) if r % p_c == 0)
        nn = sum(loc[r][1] for r in range(P) if r // p_c == 0)
        m, n = [mm] * P, [nn] * P
    sh = Shards()
    sh.topo, sh.m, sh.n = topo, m, n
    sh.m_loc, sh.n_loc, sh.W_start, sh.W_end, sh.H_start, sh.H_end = [], [], [], [], [], []
    for r in range(P):
        if topo == '2d':
            rm, gm, shp_m = grid.col_rank(r), (p_c, 1), (loc[r][0], k)
            rn, gn, shp_n = grid.row_rank(r), (1, p_r), (k, loc[r][1])
        else:
            rm, gm, shp_m = r, (p_r, 1), (m[r], k)
            rn, gn, shp_n = r, (1, p_c), (k, n[r])
        ws, we = block_range(rm, gm, shp_m)
        hs, he = block_range(rn, gn, shp_n)
        sh.m_loc.append(we[0] - ws[0] + 1)
        sh.n_loc.append(he[1] - hs[1] + 1)
        sh.W_start.append(ws[0]); sh.W_end.append(we[0] + 1)
        sh.H_start.append(hs[1]); sh.H_end.append(he[1] + 1)
    return sh


def zero_idx_prune(A, grid, sh):
    """Non-zero row/column masks (utils.py:117-135).  Integer work: exact."""
    P = grid.p
    row_sum = [np.sum(a != 0, 1) for a in A]
    col_sum = [np.sum(a != 0, 0) for a in A]
    if sh.topo == '2d':
        row_sum = allreduce(row_sum, grid.col)
        col_sum = allreduce(col_sum, grid.row)
    else:
        if grid.p_c > 1:
            row_sum = allreduce(row_sum, grid.world)
        if grid.p_r > 1:
            col_sum = allreduce(col_sum, grid.world)
    rx = [s > 0 for s in row_sum]
    cx = [s > 0 for s in col_sum]
    if sh.topo == '2d':
        ch = [col_sum[r][sh.H_start[r]:sh.H_end[r]] > 0 for r in range(P)]
        rw = [row_sum[r][sh.W_start[r]:sh.W_end[r]] > 0 for r in range(P)]
    else:
        rw, ch = [s > 0 for s in row_sum], [s > 0 for s in col_sum]
    return rx, cx, rw, ch


def prune_all(A, W, H, masks):
    """utils.py:137-176."""
    rx, cx, rw, ch = masks
    A2 = [a[np.ix_(rx[r], cx[r])] for r, a in enumerate(A)]
    W2 = [w[rw[r], :] for r, w in enumerate(W)]
    H2 = [h[:, ch[r]] for r, h in enumerate(H)]
    return A2, W2, H2


def unprune_factors(W, H, masks):
    """utils.py:178-217: scatter back into float64 zeros, guarded by len>1."""
    rx, cx, rw, ch = masks
    Wo, Ho = [], []
    for r in range(len(W)):
        if len(rw[r]) > 1:
            B = np.zeros((len(rw[r]), W[r].shape[1]))
            B[rw[r], :] = W[r]
        else:  # reference would raise UnboundLocalError; keep factor as is
            B = W[r]
        Wo.append(B)
        if len(ch[r]) > 1:
            C = np.zeros((H[r].shape[0], len(ch[r])))
            C[:, ch[r]] = H[r]
        else:
            C = H[r]
        Ho.append(C)
    return Wo, Ho


# ---------------------------------------------------------------------------
# init                                              pyDNMFk/pyDNMF.py:107-135
# ---------------------------------------------------------------------------


def init_factors_rand(A, grid, sh, k, rngs):
    """RNG call order is part of the parity contract."""
    P, dt = grid.p, A[0].dtype
    W, H = [None] * P, [None] * P
    if sh.topo == '2d':
        for r in range(P):
            W[r] = rngs[r].rand(sh.m_loc[r], k).astype(dt)
            H[r] = rngs[r].rand(k, sh.n_loc[r]).astype(dt)
    elif grid.p_c == 1:
        for r in range(P):
            W[r] = rngs[r].rand(A[r].shape[0], k).astype(dt)
        H0 = rngs[0].rand(k, A[0].shape[1]).astype(dt)
        H = [H0.copy() for _ in range(P)]
    else:  # p_r == 1
        for r in range(P):
            H[r] = rngs[r].rand(k, A[r].shape[1]).astype(dt)
        W0 = rngs[0].rand(A[0].shape[0], k).astype(dt)
        W = [W0.copy() for _ in range(P)]
    return W, H


# ---------------------------------------------------------------------------
# perturbation                                      pyDNMFk/pyDNMFk.py:37-50
# ---------------------------------------------------------------------------


def perturb(X, noise_var, method, rng):
    if method == 'uniform':
        M = 2 * noise_var * rng.random_sample(X.shape).astype(X.dtype) + noise_var
        M = M + 1
        return np.multiply(X, M)
    if method == 'poisson':
        return rng.poisson(X).astype(X.dtype)
    return 0


# ---------------------------------------------------------------------------
# update ops
# ---------------------------------------------------------------------------


class _State:
    pass


def _gram(X, grid, do_reduce):
    """X^T X per rank + optional world allreduce (dist_nmf.py:94-116, 662-685)."""
    g = [np.matmul(x.T, x) for x in X]
    return allreduce(g, grid.world) if do_reduce else g


def _mm(Xs, Ys, grid, do_reduce):
    """dist_nmf.py:687-711."""
    g = [np.matmul(x, y) for x, y in zip(Xs, Ys)]
    return allreduce(g, grid.world) if do_reduce else g


def _gather_H(st):
    """H_ij -> H_j over the row communicator (dist_nmf.py:195-197, 284-287)."""
    lst = allgather(st.H, st.grid.row)
    return [np.hstack(l) for l in lst]


def _gather_W(st):
    """W_ij -> W_i over the column communicator (dist_nmf.py:163-165, 288-291)."""
    lst = allgather(st.W, st.grid.col)
    return [np.vstack(l) for l in lst]


def _AH(st, H=None):
    """A H^T: 1-D dist_nmf.py:730; 2-D AH_glob dist_nmf.py:174-205."""
    g, dt = st.grid, st.dt
    if st.topo == '1d':
        Hs = st.H if H is None else H
        return _mm(st.A, [h.T for h in Hs], g, g.p_c != 1)
    saved = st.H
    if H is not None:
        st.H = H
    Hj = _gather_H(st)
    st.H = saved
    V = [np.matmul(a, h.T) for a, h in zip(st.A, Hj)]
    sizes = [st.W[r].shape[0] * st.k for r in range(g.p)]
    out = reduce_scatter(V, sizes, g.col, dt)
    return [o.reshape(st.W[r].shape[0], st.k) for r, o in enumerate(out)]


def _WTA(st):
    """W^T A: 1-D dist_nmf.py:749; 2-D ATW_glob dist_nmf.py:144-172."""
    g, dt = st.grid, st.dt
    if st.topo == '1d':
        return _mm([w.T for w in st.W], st.A, g, g.p_r != 1)
    Wi = _gather_W(st)
    Y = [np.matmul(w.T, a) for w, a in zip(Wi, st.A)]
    sizes = [st.H[r].shape[1] * st.k for r in range(g.p)]
    out = reduce_scatter([y.T.copy() for y in Y], sizes, g.row, dt)
    return [o.reshape(st.H[r].shape[1], st.k).T for r, o in enumerate(out)]


def _gramW(st):
    return _gram(st.W, st.grid, True if st.topo == '2d' else st.grid.p_r != 1)


def _gramHT(st, H=None):
    Hs = st.H if H is None else H
    return _gram([h.T for h in Hs], st.grid, True if st.topo == '2d' else st.grid.p_c != 1)


def fro_mu_update(st):
    """dist_nmf.py:207-263 (2-D) / 715-771 (1-D): W first, then H with new W."""
    P = st.grid.p
    if st.W_update:
        HHT = _gramHT(st)
        AH = _AH(st)
        for r in range(P):
            WHTH = np.matmul(st.W[r], HHT[r]) + st.eps
            st.W[r] *= AH[r] / WHTH
    WTW = _gramW(st)
    AtW = _WTA(st)
    for r in range(P):
        HWtW = np.matmul(st.H[r].T, WTW[r]) + st.eps
        st.H[r] *= AtW[r] / HWtW.T


def _axis_sum(X, axis, st, reduce_p):
    """sum_along_axis dist_nmf.py:775-801 / sum_axis dist_nmf.py:345-349."""
    s = [x.sum(axis=axis) for x in X]
    if st.topo == '2d' or reduce_p != 1:
        s = allreduce(s, st.grid.world)
    return s


def kl_mu_update(st):
    """dist_nmf.py:293-407 (2-D) / 803-869 (1-D)."""
    g, P, dt = st.grid, st.grid.p, st.dt
    if st.W_update:
        x2 = _axis_sum(st.H, 1, st, g.p_c)
        if st.topo == '1d':
            U = [a / (w @ h + st.eps) for a, w, h in zip(st.A, st.W, st.H)]
            sk = _mm(U, [h.T for h in st.H], g, g.p_c != 1)
        else:
            Hj, Wi = _gather_H(st), _gather_W(st)
            U = [a / (w.dot(h) + st.eps) for a, w, h in zip(st.A, Wi, Hj)]
            V = [u.dot(h.T) for u, h in zip(U, Hj)]
            sizes = [st.W[r].shape[0] * st.k for r in range(P)]
            out = reduce_scatter(V, sizes, g.col, dt)
            sk = [o.reshape(st.W[r].shape[0], st.k) for r, o in enumerate(out)]
        for r in range(P):
            X2 = np.tile(x2[r], (st.W[r].shape[0], 1))
            st.W[r] *= sk[r] / (X2 + st.eps)
    x1 = _axis_sum(st.W, 0, st, g.p_r)
    if st.topo == '1d':
        U = [a / (w @ h + st.eps) for a, w, h in zip(st.A, st.W, st.H)]
        ks = _mm([w.T for w in st.W], U, g, g.p_r != 1)
    else:
        Hj, Wi = _gather_H(st), _gather_W(st)
        U = [a / (w.dot(h) + st.eps) for a, w, h in zip(st.A, Wi, Hj)]
        Y = [w.T.dot(u) for w, u in zip(Wi, U)]
        sizes = [st.H[r].shape[1] * st.k for r in range(P)]
        out = reduce_scatter([y.T.copy() for y in Y], sizes, g.row, dt)
        ks = [o.reshape(st.H[r].shape[1], st.k).T for r, o in enumerate(out)]
    for r in range(P):
        X1 = np.tile(x1[r], (st.H[r].shape[1], 1)).T
        st.H[r] *= ks[r] / (X1 + st.eps)


def _dist_col_norm(cols, st):
    """utils.norm (utils.py:367-391): ||.||_2^2 per rank, allreduce iff p_r != 1."""
    nm = [np.linalg.norm(c, ord=2) ** 2 for c in cols]
    if st.grid.p_r != 1:
        nm = allreduce(nm, st.grid.world)
    return [np.sqrt(v) for v in nm]


def fro_hals_update(st):
    """dist_nmf.py:411-470 (2-D) / 873-934 (1-D)."""
    P, k = st.grid.p, st.k
    if st.W_update:
        HHT = _gramHT(st)
        AH = _AH(st)
        for kk in range(k):
            for r in range(P):
                W = st.W[r]
                t = W[:, kk] * HHT[r][kk, kk] + AH[r][:, kk] - W.dot(HHT[r][:, kk])
                W[:, kk] = np.maximum(t, st.eps)
            ss = _dist_col_norm([st.W[r][:, kk] for r in range(P)], st)
            for r in range(P):
                if ss[r] > 0:
                    st.W[r][:, kk] /= ss[r]
    WTW = _gramW(st)
    AtW = _WTA(st)
    for r in range(P):
        H = st.H[r]
        for kk in range(k):
            t = H[kk, :] + AtW[r][kk, :] - WTW[r][kk, :].dot(H)
            H[kk, :] = np.maximum(t, st.eps)


def _global_sqnorm(X, st, reduce_p=-1):
    """globalSqNorm dist_nmf.py:474-480 (2-D: always allreduce) / 939-949."""
    v = []
    for x in X:
        nx = np.linalg.norm(x)
        v.append(nx * nx)
    if st.topo == '2d' or reduce_p != 1:
        v = allreduce(v, st.grid.world)
    return v


def _residual(st):
    """A - W H on each rank with the block factors (dist_nmf.py:555-556, 1024)."""
    if st.topo == '1d':
        return [a - w @ h for a, w, h in zip(st.A, st.W, st.H)]
    Hj, Wi = _gather_H(st), _gather_W(st)
    return [a - w @ h for a, w, h in zip(st.A, Wi, Hj)]


def fro_bcd_update(st, itr):
    """dist_nmf.py:474-579 (2-D) / 939-1047 (1-D).  Runs its own ``itr`` loop;
    ignores W_update.  Scalars follow the reference's numpy promotion rules
    (under numpy>=2 ``np.min([...])`` yields float64 and promotes the
    extrapolated factors)."""
    g, P = st.grid, st.grid.p
    Xnorm = _global_sqnorm(st.A, st)
    nW = _global_sqnorm(st.W, st, g.p_r)
    nH = _global_sqnorm(st.H, st, g.p_c)
    W_old = [st.W[r] / np.sqrt(nW[r]) * np.sqrt(np.sqrt(Xnorm[r])) for r in range(P)]
    H_old = [st.H[r] / np.sqrt(nH[r]) * np.sqrt(np.sqrt(Xnorm[r])) for r in range(P)]
    Wm = [w.copy() for w in W_old]
    Hm = [h.copy() for h in H_old]
    HHT = _gramHT(st, H_old)
    if st.topo == '2d':
        st.H = H_old            # dist_nmf.py:498
    AHT = _AH(st, H_old)
    obj_old = [0.5 * x for x in Xnorm]
    rw = 1
    t_old = 1
    HHTnorm = [1] * P
    WTWnorm = [1] * P
    for _ in range(itr):
        HHTnorm_old = HHTnorm
        HHTnorm = [np.linalg.norm(x) for x in HHT]
        for r in range(P):
            GW = Wm[r] @ HHT[r] - AHT[r]
            st.W[r] = np.maximum(0, Wm[r] - GW / HHTnorm[r])
        ws = [np.sum(w, 0, keepdims=True) for w in st.W]
        if st.topo == '2d' or g.p_r != 1:
            ws = allreduce(ws, g.world)
        for r in range(P):
            st.W[r] = st.W[r] / ws[r]
        WTW = _gramW(st)
        WTWnorm_old = WTWnorm
        WTWnorm = [np.linalg.norm(x) for x in WTW]
        WTA = _WTA(st)
        for r in range(P):
            GH = WTW[r] @ Hm[r] - WTA[r]
            st.H[r] = np.maximum(0, Hm[r] - GH / WTWnorm[r])
        HHT = _gramHT(st)
        AHT = _AH(st)
        jt = _global_sqnorm(_residual(st), st)
        obj = [0.5 * v for v in jt]
        t = (1 + np.sqrt(1 + 4 * t_old ** 2)) / 2
        if obj[0] >= obj_old[0]:
            Wm = [w.copy() for w in W_old]
            Hm = [h.copy() for h in H_old]
            HHT = _gramHT(st, H_old)
            AHT = _AH(st, H_old)
        else:
            w = (t_old - 1) / t
            for r in range(P):
                ww = np.min([w, rw * np.sqrt(HHTnorm_old[r] / HHTnorm[r])])
                wh = min([w, rw * np.sqrt(WTWnorm_old[r] / WTWnorm[r])])
                Wm[r] = st.W[r] + ww * (st.W[r] - W_old[r])
                Hm[r] = st.H[r] + wh * (st.H[r] - H_old[r])
            W_old = [w.copy() for w in st.W]
            H_old = [h.copy() for h in st.H]
            t_old = t
            obj_old = obj


def update(st, itr):
    """Dispatch + error messages of dist_nmf.py:66-92 / 634-660."""
    nrm, mth = st.norm.upper(), st.method.upper()
    if nrm == 'FRO':
        if mth == 'MU':
            fro_mu_update(st)
        elif mth == 'HALS':
            fro_hals_update(st)
        elif mth == 'BCD':
            fro_bcd_update(st, itr)
        else:
            raise Exception('Not a valid method: Choose (mu/hals/bcd)')
    elif nrm == 'KL':
        if mth == 'MU':
            kl_mu_update(st)
        else:
            raise Exception('Not a valid method: Choose (mu)')
    else:
        raise Exception('Not a valid norm: Choose (fro/kl)')


# ---------------------------------------------------------------------------
# driver                                           pyDNMFk/pyDNMF.py:55-218
# ---------------------------------------------------------------------------


def normalize_features(st):
    """pyDNMF.py:185-194: W /= colsum+eps ; H *= colsum (no eps)."""
    g = st.grid
    s = [w.sum(axis=0, keepdims=True) for w in st.W]
    if st.topo == '2d' or g.p_r != 1:
        s = allreduce(s, g.world)
    for r in range(g.p):
        st.W[r] /= s[r] + st.eps
        st.H[r] *= s[r].T


def relative_err(st):
    """pyDNMF.py:197-218: fp32 ``np.linalg.norm`` per rank (BLAS dot), squared,
    allreduced, sqrt."""
    R = _residual(st)
    e = allreduce([np.linalg.norm(x, ord='fro') ** 2 for x in R], st.grid.world)
    a = allreduce([np.linalg.norm(x, ord='fro') ** 2 for x in st.A], st.grid.world)
    return [np.sqrt(e[r]) / np.sqrt(a[r]) for r in range(st.grid.p)]


def fit(A_blocks, p_r, p_c, k, norm='kl', method='mu', itr=5000, rngs=None,
        factors=None, prune=True, W_update=True, init='rand', return_state=False, py_rng=None):
    """PyNMF(A_ij, factors, params).fit() on every virtual rank.

    ``A_blocks``: list of per-rank shards (world-rank order, all one dtype).
    ``rngs``: list of per-rank ``np.random.RandomState`` standing for each
    process's global legacy numpy stream.  ``factors``: optional per-rank
    ``[(W, H), ...]``.  Returns ``[(W, H, recon_err)] * P``.
    """
    grid = VGrid(p_r, p_c)
    P = grid.p
    assert len(A_blocks) == P
    st = _State()
    st.grid, st.k, st.norm, st.method, st.W_update = grid, int(k), norm, method, W_update
    st.A = [np.asarray(a) for a in A_blocks]
    st.dt = st.A[0].dtype
    st.eps = np.finfo(st.dt).eps                                  # pyDNMF.py:68
    sh = compute_dims(st.A, grid, st.k)
    st.topo, st.sh = sh.topo, sh
    if factors is not None:
        st.W = [np.asarray(f[0]).astype(st.dt) for f in factors]  # pyDNMF.py:90-96
        st.H = [np.asarray(f[1]).astype(st.dt) for f in factors]
    elif init == 'rand':
        st.W, st.H = init_factors_rand(st.A, grid, sh, st.k, rngs)
    elif init == 'nnsvd':                                         # pyDNMF.py:131-135 (float64 factors, not cast)
        if sh.topo != '1d':
            raise Exception('NNSVD init only available for 1D topology, please try with 1d topo.')
        from . import nmfk_oracle
        wh = nmfk_oracle.nnsvd(st.A, sh.m[0], sh.n[0], st.k, p_r, p_c, st.eps, py_rng)
        st.W, st.H = [w for w, _ in wh], [h for _, h in wh]
    else:
        raise ValueError(init)
    masks = None
    if prune:
        masks = zero_idx_prune(st.A, grid, sh)
        st.A, st.W, st.H = prune_all(st.A, st.W, st.H, masks)
    st.masks = masks
    err = None
    for i in range(itr):
        if method.lower() == 'bcd':
            i = itr - 1                                           # pyDNMF.py:152
        update(st, itr)
        if i % 10 == 0:                                           # pyDNMF.py:155-157
            st.H = [np.maximum(h, st.eps) for h in st.H]
            st.W = [np.maximum(w, st.eps) for w in st.W]
        if i == itr - 1:
            normalize_features(st)
            err = relative_err(st)
            break
    W, H = st.W, st.H
    if prune:
        W, H = unprune_factors(W, H, masks)
    out = [(W[r], H[r], err[r]) for r in range(P)]
    if return_state:
        return out, st
    return out


def column_err(st):
    """pyDNMF.py:221-239 per-column relative L2 error (length n, fp64)."""
    g, P = st.grid, st.grid.p
    n = st.sh.n[0]
    num = [np.zeros(n) for _ in range(P)]
    den = [np.zeros(n) for _ in range(P)]
    if st.topo == '1d':
        Wi, Hj = st.W, st.H
    else:
        Wi, Hj = _gather_W(st), _gather_H(st)
    for r in range(P):
        s, _ = block_range(r, (g.p_r, g.p_c), (st.sh.m[0], n))
        rec = Wi[r] @ Hj[r]
        nl = st.A[r].shape[1]
        num[r][s[1]:s[1] + nl] = np.sum((st.A[r] - rec) ** 2, axis=0)
        den[r][s[1]:s[1] + nl] = np.sum(st.A[r] ** 2, axis=0)
    num = allreduce(num, g.world)
    den = allreduce(den, g.world)
    return [np.sqrt(a / b) for a, b in zip(num, den)]
