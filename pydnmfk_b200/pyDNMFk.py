"""NMFk perturbation ensemble on top of ``PyNMF``.

Mirrors the hot-path part of ``pyDNMFk/pyDNMFk.py``: ``sample`` (:8-67) and the perturbation loop
/ result stacking / W-fixed regression fit of ``PyNMFk.pynmfk_per_k`` (:218-248).  Clustering,
silhouettes, the Wilcoxon rank selection and results.h5 are the "next" rows N2-N4 of SURVEY.md
section 8f and are not part of this package yet.
"""
import numpy as np
import numpy
import torch

from . import config   # noqa: F401
from . import device as D
from .pyDNMF import *  # noqa: F401,F403
from .pyDNMF import PyNMF
from .utils import comm_timing, var_init, Checkpoint


class sample():
    """Perturbed copy of a data shard (pyDNMFk.py:8-67).

    ``uniform``: X * (1 + nv + 2 nv u), u ~ U[0,1) -- what the reference's code does (its docstring
    says (1-eps, 1+eps); the code wins).  The random field is drawn on the host from the
    process-global legacy numpy stream (seeded with ``seed`` exactly like the reference, so
    ``PyNMF.init_factors`` continues the same stream) and the multiply runs on the
    device, bit-identical to numpy's fp32/fp64 result.
    ``poisson``: host ``np.random.poisson`` (same stream), uploaded.
    """

    @comm_timing()
    def __init__(self, data, noise_var, method, seed=None):
        self.X = data
        self.noise_var = noise_var
        self.seed = seed
        if self.seed != None:  # noqa: E711
            np.random.seed(self.seed)
        self.method = method
        self.X_per = 0

    def _np_dtype(self):
        return np.dtype(str(self.X.dtype).replace('torch.', ''))

    @comm_timing()
    def randM(self):
        ops = D.default_ops()
        dt = self._np_dtype()
        U = np.random.random_sample(tuple(self.X.shape)).astype(dt)
        Xd = D.to_device(self.X)
        self.X_per = ops.perturb_uniform(Xd, D.to_device(U), self.noise_var)

    @comm_timing()
    def poisson(self):
        Xh = self.X.cpu().numpy() if isinstance(self.X, torch.Tensor) else np.asarray(self.X)
        self.X_per = D.to_device(np.random.poisson(Xh).astype(Xh.dtype))

    @comm_timing()
    def fit(self):
        """Returns the perturbed shard: a CUDA tensor if the input was one, else a numpy array."""
        if self.method == 'uniform':
            self.randM()
        elif self.method == 'poisson':
            self.poisson()
        if isinstance(self.X_per, torch.Tensor) and not isinstance(self.X, torch.Tensor):
            return self.X_per.cpu().numpy()
        return self.X_per


class PyNMFk():
    r"""Perturbation ensemble for automatic rank estimation (pyDNMFk.py:126-258), ensemble part.

    ``fit_ensemble(k)`` runs ``perturbations`` independent ``PyNMF`` fits of perturbed copies of the
    resident shard and returns the stacked factors exactly as the reference stacks them before
    clustering: ``Wall (m_loc, k, P)``, ``Hall (k, n_loc, P)``, ``recon_err [P]``.
    ``fit_regression(W, H)`` is the W-fixed fit that follows the clustering (pyDNMFk.py:245-248).
    """

    @comm_timing()
    def __init__(self, A_ij, factors=None, params=None):
        self.A_ij = A_ij
        self.local_m, self.local_n = self.A_ij.shape
        self.params = params
        self.comm1 = self.params.comm1
        self.rank = self.comm1.rank
        self.p_r, self.p_c = self.params.p_r, self.params.p_c
        self.fpath = var_init(self.params, 'fpath', default='data/')
        self.fname = var_init(self.params, 'fname', default='A_')
        self.p = self.p_r * self.p_c
        self.start_k = var_init(self.params, 'start_k', default=1)
        self.end_k = var_init(self.params, 'end_k', default=10)
        self.step_k = var_init(self.params, 'step_k', default=1)
        self.sill_thr = var_init(self.params, 'sill_thr', default=0.9)
        self.verbose = var_init(self.params, 'verbose', default=False)
        self.sampling = var_init(self.params, 'sampling', default='uniform')
        self.perturbations = var_init(self.params, 'perturbations', default=20)
        self.noise_var = var_init(self.params, 'noise_var', default=.03)
        self.params.results_path = var_init(self.params, 'results_path', default='results/')
        self.Hall = 0
        self.Wall = 0
        self.recon_err = 0
        self.AvgH = 0
        self.AvgG = 0
        self.col_err = 0
        self.clusterSilhouetteCoefficients, self.avgSilhouetteCoefficients = 0, 0
        self.L_errDist = 0
        self.avgErr = 0
        self.start_time = 0
        self.end_time = 0
        self.params.checkpoint = var_init(self.params, 'checkpoint', default=False)
        self.params.rank = self.rank
        self.cp = Checkpoint(self.params.checkpoint, self.params)
        # keep the shard on the device across all perturbations
        self._A_dev = D.to_device(self.A_ij)
        self._numpy_in = not isinstance(self.A_ij, torch.Tensor)

    def _spread(self):
        return (bool(getattr(self.params, 'ensemble_parallel', False)) and self.params.comm1.size > 1
                and self.p_r * self.p_c == 1)

    def _enter_solo(self):
        """Swap the communicators on ``params`` for a private size-1 grid (replica mode)."""
        from .dist_comm import Comm, MPI_comm
        world = self.params.comm1
        saved = (self.params.comm1, self.params.comm, self.params.row_comm, self.params.col_comm)
        solo = Comm([world.ranks[world.rank]], None)
        grid = MPI_comm(solo, 1, 1)
        self.params.comm1, self.params.comm = solo, grid
        self.params.row_comm, self.params.col_comm = grid.cart_1d_row(), grid.cart_1d_column()
        return saved

    def _leave_solo(self, saved):
        self.params.comm1, self.params.comm, self.params.row_comm, self.params.col_comm = saved

    def fit_ensemble(self, k):
        """pyDNMFk.py:226-238 for one k.

        The reference runs the perturbations one after another on the whole grid.  With ``params.ensemble_parallel``
        set and a 1x1 factorization grid inside a larger world (every rank holds the full matrix on its own GPU), the
        perturbations are independent replicas: rank r takes perturbations r, r+size, ... on a private size-1
        communicator (no data-path collective) and the factors are exchanged once at the end.  Results are identical to
        the sequential order because every perturbation re-seeds its own RNG stream (seed = 1000 * perturbation).
        """
        self.k = k
        self.params.k = k
        world = self.params.comm1
        spread = self._spread()
        todo = list(range(self.perturbations))
        saved = None
        if spread:
            todo = todo[world.rank::world.size]
            saved = self._enter_solo()
        mine = {}
        for perturbation in todo:
            data = sample(data=self._A_dev, noise_var=self.noise_var, method=self.sampling,
                          seed=perturbation * 1000).fit()
            self.params.W_update = True
            W, H, err = PyNMF(data, factors=None, params=self.params).fit()
            mine[perturbation] = (W.cpu().numpy(), H.cpu().numpy(), err)
        if spread:
            self._leave_solo(saved)
            merged = {}
            for part in world.allgather(mine):
                merged.update(part)
            mine = merged
        results = [mine[p] for p in range(self.perturbations)]
        self.Wall = np.hstack(([results[i][0] for i in range(self.perturbations)]))
        self.Wall = self.Wall.reshape(self.Wall.shape[0], self.k, self.perturbations, order='F')
        self.Hall = np.vstack(([results[i][1] for i in range(self.perturbations)]))
        self.Hall = self.Hall.reshape(self.k, self.Hall.shape[1], self.perturbations)
        self.recon_err = [results[i][2] for i in range(self.perturbations)]
        return self.Wall, self.Hall, self.recon_err

    def fit_regression(self, AvgW, AvgH):
        """W-fixed fit from the cluster medians (pyDNMFk.py:245-248): only the H half-step runs.  In replica mode
        every rank runs the same (deterministic) fit on its own copy."""
        self.params.W_update = False
        saved = self._enter_solo() if self._spread() else None
        reg = PyNMF(self._A_dev, factors=[AvgW, AvgH], params=self.params)
        W, H, err = reg.fit()
        self.col_err = reg.column_err()
        if saved is not None:
            self._leave_solo(saved)
        return W.cpu().numpy(), H.cpu().numpy(), err

    @comm_timing()
    def fit(self):
        raise NotImplementedError(
            'PyNMFk.fit: clustering / silhouettes / p-value rank selection (dist_clustering.py, '
            'pyDNMFk.py:239-299) are rows N2/N4 of the scope table and not built yet; use '
            'fit_ensemble(k) and fit_regression(W, H) for the accelerated parts')
