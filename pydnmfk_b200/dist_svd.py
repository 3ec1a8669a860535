"""Distributed truncated SVD by deflated power iteration and the nnsvd initialisation built on it, device resident.

Mirrors ``pyDNMFk/dist_svd.py`` (``DistSVD``: same constructor, method names, grid assertions and return values).
Data movement differs from the reference, results do not:

  * the d x d Gram (d = min(m, n)) of the deflated shard is built by the A-streaming contractions in 64-wide chunks
    (``dnmf_wta`` / ``dnmf_ah``) and all-reduced; the shard is deflated in place by rank-1 updates in the reference's
    order (dist_svd.py:157-162), so the k-th Gram sees bit-identical input without re-copying A k times;
  * the power iteration (dist_svd.py:117-134; rank 0 + Bcast in the reference) runs redundantly on every GPU on the
    identical all-reduced Gram, in float64 like numpy's promotion of ``B @ currV``; one fused normalise + convergence
    dot per iteration;
  * start vectors come from Python's global ``random`` exactly like ``randomUnitVector`` (dist_svd.py:80-85), so a
    seeded run follows the reference's trajectory.
"""
from datetime import datetime
from math import sqrt
from random import normalvariate

import os

import numpy as np
import torch

from . import device as D
from .utils import *  # noqa: F401,F403
from .utils import comm_timing


class DistSVD():
    @comm_timing()
    def __init__(self, args, A):
        self.args = args
        self.globalm = self.args.m
        self.globaln = self.args.n
        self.k = self.args.k if self.args.k else min(self.globalm, self.globaln)
        self.svdSoFar = []
        self.comm = args.comm
        self.rank = self.comm.rank
        self.p = self.comm.size
        self.grid_comm = self.comm.cartesian2d
        self.coords = self.comm.coord2d
        self.proc_rows = self.args.p_r
        self.proc_cols = self.args.p_c
        if self.globalm > self.globaln:
            assert self.proc_rows > self.proc_cols, "m>n , ensure p_r>p_c"
        elif self.globalm < self.globaln:
            assert self.proc_rows < self.proc_cols, "m<n , ensure p_r<p_c"
        self.eps = self.args.eps
        self._numpy_in = not isinstance(A, torch.Tensor)
        self.ops = D.default_ops()
        self.A = D.to_device(A)
        self.mat_for_1D = self.A.clone()
        try:
            self.seed = self.args.seed
        except AttributeError:
            self.seed = datetime.now().timestamp()

    def _out(self, t):
        return t.cpu().numpy() if self._numpy_in else t

    # ---- dist_svd.py:68-78 ---------------------------------------------------------------------------------
    @comm_timing()
    def normalize_by_W(self, Wall, Hall, comm1):
        W = D.to_device(Wall).contiguous()
        H = D.to_device(Hall, W.dtype).contiguous()
        cs = self.ops.colsum_wide(W)
        if self.proc_rows != 1:
            cs = comm1.allreduce_(cs)
        self.ops.scale_groups(W, cs, (0, 1), mode=4, eps=self.eps)          # W /= (colsum + eps)
        self.ops.scale_groups(H, cs, (1, 0), mode=5, eps=self.eps)          # H *= (colsum + eps)^T
        return W, H

    # ---- dist_svd.py:80-85 ---------------------------------------------------------------------------------
    def randomUnitVector(self, d):
        unnormalized = [normalvariate(0, 1) for _ in range(d)]
        theNorm = sqrt(sum(x * x for x in unnormalized))
        return np.asarray([x / theNorm for x in unnormalized], dtype='float64')

    # ---- dist_svd.py:87-92 -----------------------------------------------------------------------------------
    @comm_timing()
    def globalGram(self, X, Y):
        """``X @ Y`` summed over the grid.  ``X`` streams through the A H^T contraction with 64-row chunks of ``Y^T`` as
        the skinny factor (the reference's call sites: ``(mat.T, mat)`` and ``(mat, A.T)``)."""
        Xd = D.to_device(X).contiguous()
        Yt = D.to_device(Y, Xd.dtype).t().contiguous()
        return self.grid_comm.allreduce_(self.ops.outer_gram_wide(Xd, Yt))

    # ---- dist_svd.py:94-137 --------------------------------------------------------------------------------
    @comm_timing()
    def svd1D(self):
        d = min(self.globalm, self.globaln)
        cur = D.to_device(self.randomUnitVector(d))
        if self.grid_comm.size > 1:
            cur = self.grid_comm.bcast_(cur, root=0)
        if self.globalm >= self.globaln:
            B = self.grid_comm.allreduce_(self.ops.gram_wide(self.mat_for_1D))                 # mat^T mat   (n x n)
        else:
            B = self.grid_comm.allreduce_(self.ops.outer_gram_wide(self.mat_for_1D, self.A))   # mat A^T     (m x m)
        if d <= 512 and B.dtype in (torch.float32, torch.float64) and os.environ.get('DNMF_POWER_ITERATE', '1') != '0':
            # small Gram matrix: the whole loop in one launch (same arithmetic, no host round trip per step)
            self.currV = self.ops.power_iterate(B, cur, 1. - self.eps)
            return
        r = self.ops.empty((1,), torch.float64)
        while True:
            cur = self.ops.power_normalize(self.ops.matvec_f64(B, cur), cur, r)
            if abs(float(r.item())) > 1. - self.eps:
                break
        self.currV = cur

    # ---- dist_svd.py:139-145 -------------------------------------------------------------------------------
    @comm_timing()
    def calc_norm(self, vec):
        sq = self.grid_comm.allreduce_(self.ops.sqnorm(D.to_device(vec).contiguous().view(1, -1)))
        return float(np.sqrt(sq.item()))

    # ---- dist_svd.py:147-181 -------------------------------------------------------------------------------
    @comm_timing()
    def svd(self):
        """Returns (singular values [k], U [rows, k], V [k, cols]) like the reference (float64)."""
        rows, cols = self.A.shape
        tall = self.globalm > self.globaln
        Ut = self.ops.empty((self.k, rows), torch.float64)
        Vt = self.ops.empty((self.k, cols), torch.float64)
        sq = self.ops.empty((self.k,), torch.float64)
        sig = self.ops.empty((self.k,), torch.float64)
        self.mat_for_1D = self.A.clone()
        for i in range(self.k):
            if i > 0:        # running deflation == the reference's fresh copy minus the first i terms, in order
                self.ops.rank1_sub(self.mat_for_1D, Ut[i - 1], Vt[i - 1], sig[i - 1:i])
            self.svd1D()
            if tall:
                Vt[i].copy_(self.currV)
                un = self.ops.matvec_f64(self.A, self.currV)
                dst = Ut[i]
            else:
                Ut[i].copy_(self.currV)
                un = self.ops.matvec_f64(self.A, self.currV, trans=True)
                dst = Vt[i]
            s2 = self.grid_comm.allreduce_(self.ops.sqnorm(un.view(1, -1)))
            sq[i:i + 1].copy_(s2)
            self.ops.div_store(un, s2, dst)
            sig[i:i + 1].copy_(torch.sqrt(s2))
            self.svdSoFar.append([sig[i:i + 1], Ut[i], Vt[i]])
        self._sig, self._U, self._V = sig, Ut.t().contiguous(), Vt
        return self._out(sig), self._out(self._U), self._out(self._V)

    # ---- dist_svd.py:183-192 -------------------------------------------------------------------------------
    @comm_timing()
    def rel_error(self, U, S, V):
        """||A - U S V|| / ||A|| over the grid; U [rows, k], S [k, k] diagonal (or identity), V [k, cols]."""
        dt = self.A.dtype
        Ud = D.to_device(U).to(torch.float64).contiguous().clone()
        Sd = torch.diagonal(D.to_device(S).to(torch.float64)).contiguous()
        self.ops.scale_groups(Ud, Sd, (0, 1), mode=0)
        r = self.ops.residual_sqnorm(self.A, Ud.to(dt), D.to_device(V).to(dt).contiguous())
        r = self.grid_comm.allreduce_(r)
        num, den = (float(v) for v in r.cpu().numpy()[:2])
        return np.sqrt(num) / np.sqrt(den)

    # ---- dist_svd.py:194-267 -------------------------------------------------------------------------------
    @comm_timing()
    def nnsvd(self, flag=1, verbose=1):
        self.svd()
        S, U, V = self._sig, self._U, self._V
        k = self.k
        if verbose == 1:
            recon_err_svd = self.rel_error(U, torch.diag(S), V)
            if self.rank == 0:
                print('Reconstruction error for SVD is :', recon_err_svd)
        if flag == 0:
            W = U.clone()
            H = V.clone()
            self.ops.scale_groups(H, S, (1, 0), mode=0)
            self.ops.clamp_min(W, 0.0)
            self.ops.clamp_min(H, 0.0)
        elif flag == 1:
            Vc = V.t().contiguous()                                           # [cols, k]
            un = np.sqrt(self.grid_comm.allreduce_(self.ops.posneg_colsumsq(U)).cpu().numpy())
            vn = np.sqrt(self.ops.posneg_colsumsq(Vc).cpu().numpy())          # local only (dist_svd.py:232-235)
            UP_norm, UN_norm, VP_norm, VN_norm = un[0], un[1], vn[0], vn[1]
            if self.globalm > self.globaln:
                UP_norm, UN_norm = UP_norm / self.p, UN_norm / self.p
            Sh = S.cpu().numpy()
            mp = np.sqrt(UP_norm * VP_norm * Sh)
            mn = np.sqrt(UN_norm * VN_norm * Sh)
            pos = D.to_device((mp > mn).astype(np.int32))
            W = self.ops.nnsvd_pick(U, D.to_device(np.concatenate([mp, UP_norm + self.eps, mn, UN_norm + self.eps])), pos)
            H = self.ops.nnsvd_pick(Vc, D.to_device(np.concatenate([mp, VP_norm + self.eps, mn, VN_norm + self.eps])), pos,
                                    transpose_out=True)
        if verbose == 1:
            recon_err_nnsvd = self.rel_error(W, torch.eye(k, dtype=torch.float64, device=W.device), H)
            if self.rank == 0:
                print('Reconstruction error for nnSVD is :', recon_err_nnsvd)
        W, H = self.normalize_by_W(W, H, self.grid_comm)
        if verbose == 1:
            return (self._out(W), self._out(H)), {'recon_err_svd': recon_err_svd, 'recon_err_nnsvd': recon_err_nnsvd}
        return self._out(W), self._out(H)
