"""Clustering of the perturbation ensemble and its silhouettes, device resident.

Mirrors ``pyDNMFk/dist_clustering.py`` (``custom_clustering``: same constructor, method names and return values).
The ensemble tensors ``W_all [m_loc, k, P]`` / ``H_all [k, n_loc, P]`` stay in HBM; what the reference does with
``100 x P`` tiny matmuls + blocking allreduces per k is batched over the perturbations:

  * one contraction ``centroids^T @ W_all`` for all P perturbations (``dnmf_wta`` on the m x kP view, one allreduce),
  * one launch of the greedy assignment for all P (``dnmf_greedy_lsa``; within an iteration every perturbation is
    matched against the same centroids, dist_clustering.py:110-116, so they are independent),
  * gather-permute of W columns / H rows, median over P, renormalisation -- one kernel each.

The (kP)^2 cosine Gram of the silhouettes runs on the A-streaming contraction in column chunks.
"""
import numpy as np
import torch

from . import device as D
from .utils import *  # noqa: F401,F403  (the reference module star-exports utils)
from .utils import comm_timing


class custom_clustering():
    """Greedy quadratic-assignment clustering of P groups of k vectors (dist_clustering.py:5-29)."""

    @comm_timing()
    def __init__(self, Wall, Hall, params):
        self._numpy_in = not isinstance(Wall, torch.Tensor)
        self.ops = D.default_ops()
        self._W = D.to_device(Wall).contiguous()
        self._H = D.to_device(Hall, self._W.dtype).contiguous()
        if self._W is Wall:                      # the reference works in place on the caller's arrays; device tensors
            self._W = self._W.clone()            # passed in are left untouched (results are on .W_all / .H_all)
        if self._H is Hall:
            self._H = self._H.clone()
        self.p_r, self.p_c = params.p_r, params.p_c
        self.comm1 = params.comm1
        self.eps = params.eps
        self.p = self.p_r * self.p_c

    # ---- the ensemble as the caller sees it ---------------------------------------------------------
    def _out(self, t):
        return t.cpu().numpy() if self._numpy_in else t

    @property
    def W_all(self):
        return self._out(self._W)

    @property
    def H_all(self):
        return self._out(self._H)

    def _reduce(self, t):
        return self.comm1.allreduce_(t) if self.p_r != 1 else t

    # ---- dist_clustering.py:31-39 -----------------------------------------------------------------------
    @comm_timing()
    def normalize_by_W(self):
        m, k, P = self._W.shape
        nrm = self._reduce(self.ops.colsum_wide(self._W.view(m, k * P), squares=True))       # [k*P] at kk*P + p
        self.ops.scale_groups(self._W, nrm, (0, P, 1), mode=2, eps=self.eps)                  # W /= sqrt(nrm + eps)
        self.ops.scale_groups(self._H, nrm, (P, 0, 1), mode=3, eps=self.eps)                  # H *= sqrt(nrm + eps)

    # ---- dist_clustering.py:41-47 -----------------------------------------------------------------------
    @comm_timing()
    def mad(self, data, flag=1, axis=-1):
        if flag != 1 or axis not in (-1, data.ndim - 1):
            raise NotImplementedError('only the median absolute deviation over the last axis is used by NMFk')
        t = D.to_device(data).contiguous()
        out = self.ops.median_last(t, want_mad=True)[1]
        return out if isinstance(data, torch.Tensor) else out.cpu().numpy()

    # ---- dist_clustering.py:49-69 (host versions, for callers of the public helpers) ------------------------
    def change_order(self, tens):
        ans = list(range(len(tens)))
        for p in tens:
            ans[p[0]] = p[1]
        return ans

    def greedy_lsa(self, A):
        """The greedy assignment of one k x k similarity matrix as (centroid, feature) pairs, one per centroid (the
        reference lists them in pick order; ``change_order`` gives the same result either way).  Device kernel, k <= 64."""
        t = D.to_device(np.ascontiguousarray(A) if not isinstance(A, torch.Tensor) else A).contiguous()
        k = t.shape[0]
        order = self.ops.greedy_lsa(t.view(k, k), k, 1).cpu().numpy()[0]
        return [(r, int(order[r])) for r in range(k)]

    @comm_timing()
    def dist_feature_ordering(self, centroids, W_sub):
        """dist_clustering.py:71-82 for one perturbation."""
        c = D.to_device(centroids, self._W.dtype).contiguous()
        w = D.to_device(W_sub, self._W.dtype).contiguous()
        k = w.shape[1]
        dist = self._reduce(self.ops.wta(w, c))
        j = self.ops.greedy_lsa(dist, k, 1)
        w = self.ops.permute_groups(w.view(w.shape[0], k, 1), j, axis=1).view(w.shape[0], k)
        return (w if isinstance(W_sub, torch.Tensor) else w.cpu().numpy()), [int(v) for v in j.cpu().numpy()[0]]

    # ---- dist_clustering.py:84-130 ----------------------------------------------------------------------------
    @comm_timing()
    def dist_custom_clustering(self, centroids=None, vb=0):
        self.normalize_by_W()
        m, k, P = self._W.shape
        if centroids is None:
            cent = self._W[:, :, 0].contiguous()
        else:
            cent = D.to_device(centroids, self._W.dtype).contiguous().clone()
        orders = []
        for _ in range(100):
            dist = self._reduce(self.ops.wta(self._W.view(m, k * P), cent))      # [k, k*P]: every perturbation at once
            order = self.ops.greedy_lsa(dist, k, P)
            self._W = self.ops.permute_groups(self._W, order, axis=1)
            # H rows: the reference assigns a list of row *views* back into the same slice (dist_clustering.py:116); under
            # numpy >= 1.20 that copies row by row in place, so rows overwritten earlier are read back.  Reproduced as is.
            self._H = self.ops.permute_groups(self._H, order, axis=0, sequential=True)
            orders.append(order)
            cent = self.ops.median_last(self._W)
            cn = self._reduce(self.ops.colsum_wide(cent, squares=True))
            self.ops.scale_groups(cent, cn, (0, 1), mode=2, eps=self.eps)
        permute_order = [[int(v) for v in row] for row in torch.stack(orders).cpu().numpy().reshape(100 * P, k)]
        return self._out(cent), self.W_all, self.H_all, permute_order

    # ---- dist_clustering.py:132-160 ---------------------------------------------------------------------------
    @comm_timing()
    def dist_silhouettes(self):
        self.dist_custom_clustering()
        m, k, P = self._W.shape
        G = self._reduce(self.ops.gram_wide(self._W.view(m, k * P)))
        return self.ops.silhouettes(G, k, P).cpu().numpy()

    # ---- dist_clustering.py:162-188 ---------------------------------------------------------------------------
    @comm_timing()
    def fit(self):
        centroids, _, _, IDX_F2 = self.dist_custom_clustering()
        CentStd = self._out(self.ops.median_last(self._W, want_mad=True)[1])
        cluster_coefficients = self.dist_silhouettes()
        S_avg = cluster_coefficients.flatten().mean()
        return [centroids, CentStd, self.H_all, cluster_coefficients.mean(axis=1), S_avg, IDX_F2]
