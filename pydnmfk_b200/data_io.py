"""Shard input / factor output -- thin host code (file I/O is outside the
accelerated path; on-disk formats are the "next" row N3).

``data_read`` follows ``pyDNMFk/data_io.py:12-105``: every rank reads the file named by
``fpath + fname + '.' + ftype`` (or its own ``fname<rank>.npy`` for ``ftype='folder'``), keeps the
block ``determine_block_params`` assigns to it and casts to ``precision``.  ``data_write.save_factors``
writes the per-rank ``.npy`` factor files of ``data_io.py:175-196``.
"""
import os

import numpy as np

from .utils import determine_block_params, comm_timing


class data_read():
    @comm_timing()
    def __init__(self, args):
        self.fpath = args.fpath
        self.pgrid = args.grid if ("grid" in vars(args) and args.grid) else [args.p_r, args.p_c]
        self.ftype = args.ftype
        self.fname = args.fname
        self.comm = args.comm1
        self.rank = self.comm.rank
        self.precision = args.precision if getattr(args, 'precision', None) else 'float32'
        self.data = 0
        if self.ftype == 'folder':
            self.file_path = self.fpath + self.fname + str(self.comm.rank) + '.npy'
        else:
            self.file_path = self.fpath + self.fname + '.' + self.ftype

    @comm_timing()
    def read(self):
        return self.read_dat()

    def read_file_npy(self):
        self.data = np.load(self.file_path)

    def read_file_csv(self):
        self.data = np.loadtxt(self.file_path, delimiter=',', ndmin=2)

    def read_file_mat(self):
        from scipy.io import loadmat
        self.data = loadmat(self.file_path)['X']

    def data_partition(self):
        blk = determine_block_params(self.rank, self.pgrid, self.data.shape).determine_block_index_range_asymm()
        self.data = self.data[blk[0][0]:blk[1][0] + 1, blk[0][1]:blk[1][1] + 1]

    @comm_timing()
    def read_dat(self):
        if self.ftype == 'npy':
            self.read_file_npy()
            self.data_partition()
        elif self.ftype in ('csv', 'txt'):
            self.read_file_csv()
            self.data_partition()
        elif self.ftype == 'mat':
            self.read_file_mat()
            self.data_partition()
        if self.ftype == 'folder':
            self.read_file_npy()
        return np.ascontiguousarray(self.data.astype(self.precision))


class data_write():
    """Per-rank factor files ``W_factors/W_<rank>.npy`` / ``H_factors/H_<rank>.npy`` (1-D grids write
    the replicated factor once), or ``*_reg_factors`` for the NMFk regression fit."""

    @comm_timing()
    def __init__(self, args):
        self.p_r, self.p_c = args.p_r, args.p_c
        self.pgrid = [self.p_r, self.p_c]
        self.fpath = args.results_paths if hasattr(args, 'results_paths') else args.results_path
        self.comm = args.comm1
        self.rank = self.comm.rank
        self.params = args

    def create_folder_dir(self, fpath):
        os.makedirs(fpath, exist_ok=True)

    @comm_timing()
    def save_factors(self, factors, reg=False):
        self.create_folder_dir(self.fpath)
        wdir, hdir = ('W_reg_factors/', 'H_reg_factors/') if reg else ('W_factors/', 'H_factors/')
        if self.rank == 0:
            self.create_folder_dir(self.fpath + wdir)
            self.create_folder_dir(self.fpath + hdir)
        self.comm.barrier()
        W, H = np.asarray(factors[0]), np.asarray(factors[1])
        if self.p_r == 1 and self.p_c != 1:
            if self.rank == 0:
                np.save(self.fpath + wdir + 'W', W)
            np.save(self.fpath + hdir + 'H_' + str(self.rank), H)
        elif self.p_c == 1 and self.p_r != 1:
            if self.rank == 0:
                np.save(self.fpath + hdir + 'H', H)
            np.save(self.fpath + wdir + 'W_' + str(self.rank), W)
        else:
            np.save(self.fpath + wdir + 'W_' + str(self.rank), W)
            np.save(self.fpath + hdir + 'H_' + str(self.rank), H)

    @comm_timing()
    def save_cluster_results(self, params):
        """Per-k NMFk statistics on rank 0 (data_io.py:199-209): ``results.h5`` with the reference's dataset names when
        ``h5py`` is importable, else the same names in ``results.npz`` (this image ships no h5py)."""
        if self.rank == 0:
            write_results(self.fpath, {
                'clusterSilhouetteCoefficients': params['clusterSilhouetteCoefficients'],
                'avgSilhouetteCoefficients': params['avgSilhouetteCoefficients'],
                'L_err': params['L_err'], 'L_errDist': params['L_errDist'], 'avgErr': params['avgErr'],
                'ErrTol': params['recon_err'], 'AIC': params['AIC']})


def _h5py():
    try:
        import h5py
        return h5py
    except ImportError:
        return None


def write_results(dirpath, datasets):
    h5 = _h5py()
    if h5 is not None:
        with h5.File(dirpath + 'results.h5', 'w') as hf:
            for name, val in datasets.items():
                hf.create_dataset(name, data=val)
    else:
        np.savez(dirpath + 'results.npz', **{k: np.asarray(v) for k, v in datasets.items()})


def read_results(dirpath):
    """{dataset name: ndarray} of one k's ``results.h5`` / ``results.npz`` (pyDNMFk.py:278, plot_results.py:117)."""
    h5 = _h5py()
    if h5 is not None and os.path.exists(os.path.join(dirpath, 'results.h5')):
        with h5.File(os.path.join(dirpath, 'results.h5'), 'r') as hf:
            return {k: np.array(hf[k]) for k in hf.keys()}
    with np.load(os.path.join(dirpath, 'results.npz')) as z:
        return {k: z[k] for k in z.files}


class read_factors():
    """Reassemble the regression factors written by ``save_factors(reg=True)`` (data_io.py:212-261).  The H blocks of
    a 2-D grid are concatenated in the order the shards really cover the columns, (j, i); the reference's
    ``transform_H_index`` formula is only right for square grids."""

    @comm_timing()
    def __init__(self, factors_path, pgrid):
        self.factors_path = factors_path
        self.W_path = self.factors_path + 'W_reg_factors/*'
        self.H_path = self.factors_path + 'H_reg_factors/*'
        self.p_grid = pgrid
        self.load_factors()

    def custom_read_npy(self, fpath):
        return np.load(fpath)

    def read_factor(self, fpath):
        import glob
        files = glob.glob(fpath)
        if len(files) == 1:
            return self.custom_read_npy(files[0]), 1
        key = lambda f: int(os.path.splitext(os.path.basename(f))[0].split('_')[-1])   # noqa: E731  (W_10 after W_9)
        return [self.custom_read_npy(f) for f in sorted(files, key=key)], len(files)

    @comm_timing()
    def load_factors(self):
        W_data, ct_W = self.read_factor(self.W_path)
        H_data, ct_H = self.read_factor(self.H_path)
        if ct_W > 1:
            W_data = np.vstack(W_data)
        if ct_H > 1:
            if ct_W > 1:
                p_r, p_c = self.p_grid
                H_data = np.hstack([H_data[i * p_c + j] for j in range(p_c) for i in range(p_r)])
            else:
                H_data = np.hstack(H_data)
        self.W, self.H = W_data, H_data
        return W_data, H_data
