"""``pyDNMFk_Runner``: the reference's keyword-argument front end (pyDNMFk/runner.py:12-191).

    runner = pyDNMFk_Runner(itr=1000, init='rand', norm='kl', method='mu', process='pyDNMFk', perturbations=20)
    results = runner.run(grid=[p_r, p_c], fpath='data/', fname='wtsi', ftype='mat', results_path='results/',
                         k_range=[2, 10], step_k=1)           # {'nopt': ...}  or  {'W', 'H', 'err'} for process='pyDNMF'

As in the reference the runner object itself is the ``params`` bag handed to ``data_read`` / ``PyNMF`` / ``PyNMFk``.
Launch one process per GPU with ``torchrun`` where the reference uses ``mpirun``.
"""
from . import config
from .data_io import data_read
from .dist_comm import MPI, MPI_comm
from .pyDNMF import PyNMF
from .pyDNMFk import PyNMFk

# constructor keywords and their defaults (runner.py:13-17)
_OPTIONS = (('init', 'rand'), ('itr', 5000), ('norm', 'kl'), ('method', 'mu'), ('verbose', False), ('checkpoint', False),
            ('timing_stats', False), ('prune', False), ('precision', 'float32'), ('perturbations', 20),
            ('noise_var', 0.015), ('sill_thr', 0.6), ('sampling', 'uniform'), ('process', 'pyDNMF'))
# attributes that only exist after run() (runner.py:62-68, :81-84)
_LATE = ('fpath', 'ftype', 'fname', 'results_path', 'k_range', 'step_k', 'p_r', 'p_c', 'start_k', 'end_k')


class pyDNMFk_Runner:
    def __init__(self, *args, **kwargs):
        names = [n for n, _ in _OPTIONS]
        if len(args) > len(names):
            raise TypeError('pyDNMFk_Runner takes at most %d positional arguments' % len(names))
        given = dict(zip(names, args))
        for key, val in kwargs.items():
            if key not in names:
                raise TypeError("pyDNMFk_Runner got an unexpected keyword argument '%s'" % key)
            if key in given:
                raise TypeError("pyDNMFk_Runner got multiple values for argument '%s'" % key)
            given[key] = val
        for name, default in _OPTIONS:
            setattr(self, name, given.get(name, default))
        for name in _LATE:
            setattr(self, name, None)
        if self.process not in ["pyDNMFk", "pyDNMF"]:
            raise ValueError("process should be either pyDNMFk or pyDNMF")
        config.init(0)
        config.flag = self.timing_stats
        self.main_comm = MPI.COMM_WORLD
        self.rank = self.main_comm.rank

    def run(self, grid, fpath="data/", ftype="mat", fname="A_", results_path="results/", k_range=[1, 10], step_k=1, k=4):
        """Read this rank's shard and factorize it (``process='pyDNMF'``) or estimate the rank (``'pyDNMFk'``)."""
        if len(grid) != 2 or len(k_range) != 2:
            raise ValueError("grid and k_range needs to be a list sized 2")
        self.grid = grid
        self.p_r, self.p_c = grid
        self.k_range = k_range
        self.start_k, self.end_k = k_range
        self.fpath, self.ftype, self.fname, self.results_path, self.step_k, self.k = fpath, ftype, fname, results_path, step_k, k
        self.comm = MPI_comm(self.main_comm, self.p_r, self.p_c)
        self.comm1 = self.comm.comm
        self.col_comm = self.comm.cart_1d_column()
        self.row_comm = self.comm.cart_1d_row()
        talk = self.verbose and self.rank == 0
        if talk:
            print("Reading data now")
        A_ij = data_read(self).read()
        if talk:
            print("Reading data complete")
            print('Starting ' + self.process + '...')
        results = dict()
        if self.process == "pyDNMFk":
            results["nopt"] = PyNMFk(A_ij, factors=None, params=self).fit()
        else:
            results["W"], results["H"], results["err"] = PyNMF(A_ij, factors=None, params=self).fit()
        if talk:
            print('Done ' + self.process + '.')
        if self.rank == 0 and self.timing_stats:
            if self.verbose:
                print(config.time)
            with open(self.results_path + 'Timing_stats.csv', 'w') as f:      # one-row table like DataFrame([time]).to_csv
                keys = list(config.time.keys())
                f.write(',' + ','.join(keys) + '\n0,' + ','.join(str(config.time[key]) for key in keys) + '\n')
        return results
