"""CPU oracle for the NMFk-level rows: numpy restatement of pyDNMFk's distributed
clustering / silhouettes, nnsvd initialisation, rank selection and the per-k NMFk driver.

TEST INFRASTRUCTURE ONLY (same rules as oracle/nmf_oracle.py: never imported by the product).
Parity status: PINNED -- tests/test_oracle_nmfk.py checks every function against tests/golden/nmfk_cases.npz, which
oracle/gen_golden_nmfk.py produced by running the unmodified reference on forked ranks, and against the reference's own
fixtures for this path (tests/sill.npy, tests/nnsvd_factors_*.npy; copied to tests/golden/ref_*).

Virtual ranks as in nmf_oracle.py: per-rank quantities are lists in world-rank order.  File:line citations are relative
to the reference root.
"""
import math

import numpy as np

from . import nmf_oracle as O


def _world(P):
    return [list(range(P))]


# ------------------------------------------------------------------------------------------------
# custom_clustering                                             pyDNMFk/dist_clustering.py:5-188
# ------------------------------------------------------------------------------------------------
def greedy_lsa(A):
    """dist_clustering.py:58-69: repeatedly take the largest remaining entry, strike its row and column."""
    X = A.copy()
    pairs = []
    for _ in range(X.shape[0]):
        ind = np.unravel_index(np.argmax(X), X.shape)
        pairs.append(ind)
        X[:, ind[1]] = -np.inf
        X[ind[0], :] = -np.inf
    return pairs


def change_order(pairs):
    """dist_clustering.py:49-55."""
    ans = list(range(len(pairs)))
    for p in pairs:
        ans[p[0]] = p[1]
    return ans


class Clustering:
    """State of ``custom_clustering`` on P = p_r virtual ranks (W_all split by rows, H_all per rank)."""

    def __init__(self, W_blocks, H_blocks, p_r, eps):
        self.W = [np.array(w) for w in W_blocks]
        self.H = [np.array(h) for h in H_blocks]
        self.P = len(self.W)
        self.p_r = p_r
        self.eps = eps

    def _reduce(self, vals):
        return O.allreduce(vals, _world(self.P)) if self.p_r != 1 else vals

    def normalize_by_W(self):
        """dist_clustering.py:31-39."""
        nrm = self._reduce([(w * w).sum(axis=0) for w in self.W])
        for r in range(self.P):
            t = np.sqrt(nrm[r] + self.eps)
            self.W[r] /= t.reshape(1, t.shape[0], t.shape[1])
            self.H[r] *= t.reshape(t.shape[0], 1, t.shape[1])

    def custom_clustering(self):
        """dist_clustering.py:84-130 (centroids=None)."""
        orders = []
        self.normalize_by_W()
        cent = [w[:, :, 0].copy() for w in self.W]
        n_pert = self.W[0].shape[-1]
        for _ in range(100):
            for p in range(n_pert):
                dist = self._reduce([cent[r].T @ self.W[r][:, :, p] for r in range(self.P)])
                j = change_order(greedy_lsa(dist[0]))          # identical on every rank
                orders.append(j)
                for r in range(self.P):
                    self.W[r][:, :, p] = self.W[r][:, :, p][:, j]
                    self.H[r][:, :, p] = [self.H[r][:, :, p][q] for q in j]
            cent = [np.median(w, axis=-1) for w in self.W]
            cn = self._reduce([(c ** 2).sum(axis=0) for c in cent])
            for r in range(self.P):
                cent[r] /= np.sqrt(cn[r] + self.eps)
        return cent, orders

    def silhouettes(self):
        """dist_clustering.py:132-160 (re-runs the clustering first, :142)."""
        self.custom_clustering()
        N, k, n_pert = self.W[0].shape
        G = self._reduce([(w.reshape(w.shape[0], k * n_pert).T @ w.reshape(w.shape[0], k * n_pert)).reshape(k, n_pert, k, n_pert)
                          for w in self.W])[0]
        d = np.arccos(np.clip(G, -1.0, 1.0))
        if k == 1:
            return np.ones((k, n_pert))
        a = np.zeros((k, n_pert))
        b = np.zeros((k, n_pert))
        for kk in range(k):
            for n in range(n_pert):
                a[kk, n] = 1 / (n_pert - 1) * np.sum(d[kk, n, kk, :])
                tmp = np.sum(d[kk, n, :, :], axis=1)
                tmp[kk] = np.inf
                b[kk, n] = 1 / n_pert * np.min(tmp)
        return (b - a) / np.maximum(a, b)

    def fit(self):
        """dist_clustering.py:162-188 -> per-rank [centroids, mad, H_all, sil per cluster, mean sil, orders]."""
        cent, orders = self.custom_clustering()
        mad = [np.nanmedian(np.absolute(w - np.nanmedian(w, axis=-1, keepdims=True)), axis=-1) for w in self.W]
        sils = self.silhouettes()
        return [[cent[r], mad[r], self.H[r], sils.mean(axis=1), sils.flatten().mean(), orders] for r in range(self.P)]


# ------------------------------------------------------------------------------------------------
# DistSVD / nnsvd                                                     pyDNMFk/dist_svd.py:9-267
# ------------------------------------------------------------------------------------------------
def random_unit_vector(d, py_rng):
    """dist_svd.py:80-85 (``py_rng`` stands for the process's global Python ``random`` module)."""
    un = [py_rng.normalvariate(0, 1) for _ in range(d)]
    nrm = math.sqrt(sum(x * x for x in un))
    return np.asarray([x / nrm for x in un], dtype='float64')


def nnsvd(A_blocks, m, n, k, p_r, p_c, eps, py_rng):
    """``DistSVD(args, A).nnsvd(flag=1)`` on a 1-D grid; returns per-rank (W, H).  Only rank 0's start vectors matter
    (they are broadcast, dist_svd.py:101-103) but every rank draws, so ``py_rng`` is advanced once per component."""
    P = p_r * p_c
    assert len(A_blocks) == P
    if m > n:
        assert p_r > p_c, "m>n , ensure p_r>p_c"             # dist_svd.py:53-56
    elif m < n:
        assert p_r < p_c, "m<n , ensure p_r<p_c"
    A = [np.asarray(a) for a in A_blocks]
    world = _world(P)
    so_far = []
    for i in range(k):                                        # dist_svd.py:155-178
        mats = [a.copy() for a in A]
        for sig, u, v in so_far[:i]:
            for r in range(P):
                mats[r] -= sig[r] * np.outer(u[r], v[r])
        v0 = random_unit_vector(min(m, n), py_rng)             # svd1D, dist_svd.py:97-137
        if m >= n:
            B = O.allreduce([x.T @ x for x in mats], world)[0]
        else:
            B = O.allreduce([x @ a.T for x, a in zip(mats, A)], world)[0]
        cur = v0
        while True:
            last = cur
            cur = B @ cur
            cur = cur / np.linalg.norm(cur)
            if abs(np.dot(cur, last).item()) > 1. - eps:
                break
        cur = np.zeros(cur.shape) + cur
        if m > n:
            v = [cur] * P
            un = [a @ cur for a in A]
            s = np.sqrt(O.allreduce([sum(x * x) for x in un], world)[0])
            u = [x / s for x in un]
        else:
            u = [cur] * P
            un = [a.T @ cur for a in A]
            s = np.sqrt(O.allreduce([sum(x * x) for x in un], world)[0])
            v = [x / s for x in un]
        so_far.append(([s] * P, u, v))
    out = []
    S = np.asarray([t[0][0] for t in so_far])
    Us = [np.asarray([t[1][r] for t in so_far]).T for r in range(P)]        # (rows, k)
    Vs = [np.asarray([t[2][r] for t in so_far]).T for r in range(P)]        # V.T of dist_svd.py:221
    UP = [np.where(U > 0, U, 0) for U in Us]
    UN = [np.where(U < 0, -U, 0) for U in Us]
    VP = [np.where(V > 0, V, 0) for V in Vs]
    VN = [np.where(V < 0, -V, 0) for V in Vs]
    UPn = np.sqrt(O.allreduce([np.sum(np.square(x), 0) for x in UP], world)[0])
    UNn = np.sqrt(O.allreduce([np.sum(np.square(x), 0) for x in UN], world)[0])
    if m > n:
        UPn, UNn = UPn / P, UNn / P                            # dist_svd.py:236-237
    W, H = [], []
    for r in range(P):
        VPn = np.sqrt(np.sum(np.square(VP[r]), 0))             # local only (dist_svd.py:232-235)
        VNn = np.sqrt(np.sum(np.square(VN[r]), 0))
        mp = np.sqrt(UPn * VPn * S)
        mn = np.sqrt(UNn * VNn * S)
        W.append(np.where(mp > mn, mp * UP[r] / (UPn + eps), mn * UN[r] / (UNn + eps)))
        H.append(np.where(mp > mn, mp * VP[r] / (VPn + eps), mn * VN[r] / (VNn + eps)).T)
    # normalize_by_W (dist_svd.py:68-78): column sums (not norms), eps added before both scalings
    cs = [w.sum(axis=0, keepdims=True) for w in W]
    if p_r != 1:
        cs = O.allreduce(cs, world)
    for r in range(P):
        c = cs[r] + eps
        out.append((W[r] / c, H[r] * c.T))
    return out


# ------------------------------------------------------------------------------------------------
# rank selection                                                   pyDNMFk/pyDNMFk.py:261-299
# ------------------------------------------------------------------------------------------------
def pvalue_analysis(L_err, sil_min, k_range, sill_thr):
    """``L_err[i]`` / ``sil_min[i]`` belong to ``k_range[i]``; sil_min already rounded to 2 decimals (pyDNMFk.py:281)."""
    from scipy.stats import wilcoxon
    pvalue = np.ones(len(k_range))
    one = L_err[0]
    i, nopt = 1, 1
    while i < len(k_range):
        if sil_min[i - 1] > sill_thr:
            pvalue[i] = wilcoxon(one, L_err[i])[1]
            if pvalue[i] < 0.05:
                nopt = i
                one = np.copy(L_err[i])
        i += 1
    return k_range[nopt - 1], pvalue
