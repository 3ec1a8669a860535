"""Shard input / factor output -- thin host code (file I/O is outside the
accelerated path; on-disk formats are the "next" row N3).

``data_read`` follows ``pyDNMFk/data_io.py:12-105``: every rank reads the file named by
``fpath + fname + '.' + ftype`` (or its own ``fname<rank>.npy`` for ``ftype='folder'``), keeps the
block ``determine_block_params`` assigns to it and casts to ``precision``.  ``data_write.save_factors``
writes the per-rank ``.npy`` factor files of ``data_io.py:175-196``.
"""
import os
import struct

import numpy as np

from .utils import determine_block_params, comm_timing, transform_H_index


def _load_npy(path):
    return np.load(path)


def _load_text(path):
    return np.loadtxt(path, delimiter=',', ndmin=2)


def _load_mat(path):
    from scipy.io import loadmat
    return loadmat(path)['X']


# ftype -> (loader of the file at fpath + fname + suffix, whether every rank then keeps only its own block)
_FORMATS = {'npy': (_load_npy, True), 'csv': (_load_text, True), 'txt': (_load_text, True), 'mat': (_load_mat, True),
            'folder': (_load_npy, False)}


class data_read():
    """``data_read(args).read()`` -> this rank's shard as a C-contiguous array of ``args.precision``."""

    @comm_timing()
    def __init__(self, args):
        self.fpath, self.fname, self.ftype = args.fpath, args.fname, args.ftype
        self.pgrid = args.grid if ("grid" in vars(args) and args.grid) else [args.p_r, args.p_c]
        self.comm = args.comm1
        self.rank = self.comm.rank
        self.precision = getattr(args, 'precision', None) or 'float32'
        self.data = 0
        suffix = str(self.rank) + '.npy' if self.ftype == 'folder' else '.' + self.ftype
        self.file_path = self.fpath + self.fname + suffix

    @comm_timing()
    def read(self):
        return self.read_dat()

    # the reference's per-format entry points, kept for callers that use them directly
    def read_file_npy(self):
        self.data = _load_npy(self.file_path)

    def read_file_csv(self):
        self.data = _load_text(self.file_path)

    def read_file_mat(self):
        self.data = _load_mat(self.file_path)

    def data_partition(self):
        (r0, c0), (r1, c1) = determine_block_params(self.rank, self.pgrid, self.data.shape).determine_block_index_range_asymm()
        self.data = self.data[r0:r1 + 1, c0:c1 + 1]

    @comm_timing()
    def read_dat(self):
        if self.ftype in _FORMATS:
            loader, whole_matrix = _FORMATS[self.ftype]
            self.data = loader(self.file_path)
            if whole_matrix:
                self.data_partition()
        return np.ascontiguousarray(self.data.astype(self.precision))


class split_files_save():
    """Cut a global matrix into the p_r x p_c blocks of ``determine_block_params`` and write them as ``A_<rank>.npy``
    under ``fpath``: the layout ``ftype='folder'`` reads back (data_io.py:108-136; the reference's writer stores the whole
    matrix in every file, this one stores each rank's block)."""

    @comm_timing()
    def __init__(self, data, pgrid, fpath):
        self.data = data
        self.pgrid = pgrid
        self.p_r, self.p_c = pgrid[0], pgrid[1]
        self.fpath = fpath
        os.makedirs(self.fpath, exist_ok=True)

    @comm_timing()
    def split_files(self):
        self.split = []
        for rank in range(self.p_r * self.p_c):
            (r0, c0), (r1, c1) = determine_block_params(rank, self.pgrid, self.data.shape).determine_block_index_range_asymm()
            self.split.append(self.data[r0:r1 + 1, c0:c1 + 1])
        return self.split

    @comm_timing()
    def save_data_to_file(self):
        for rank, block in enumerate(self.split_files()):
            np.save(self.fpath + 'A_' + str(rank) + '.npy', block)


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
        tag = '_reg_factors/' if reg else '_factors/'
        dirs = {'W': self.fpath + 'W' + tag, 'H': self.fpath + 'H' + tag}
        if self.rank == 0:
            for d in dirs.values():
                self.create_folder_dir(d)
        self.comm.barrier()
        # a factor that is replicated over a 1-D grid is written once (by rank 0, without a rank suffix)
        replicated = {'W': self.p_r == 1 and self.p_c != 1, 'H': self.p_c == 1 and self.p_r != 1}
        for name, arr in zip(('W', 'H'), factors):
            if not replicated[name]:
                np.save(dirs[name] + name + '_' + str(self.rank), np.asarray(arr))
            elif self.rank == 0:
                np.save(dirs[name] + name, np.asarray(arr))

    @comm_timing()
    def save_cluster_results(self, params):
        """Per-k NMFk statistics on rank 0 (data_io.py:199-209): ``results.h5`` with the reference's dataset names
        (``write_results``: h5py when importable, else the minimal earliest-format writer ``h5min`` plus ``results.npz``)."""
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
    """``results.h5`` with the reference's dataset names (data_io.py:199-209): through ``h5py`` when it is importable, else
    through the minimal writer of ``h5min`` (earliest-format HDF5; this image ships neither h5py nor libhdf5).  Without
    h5py the same datasets also go to ``results.npz`` so that a reader never depends on the minimal writer alone."""
    h5 = _h5py()
    if h5 is not None:
        with h5.File(dirpath + 'results.h5', 'w') as hf:
            for name, val in datasets.items():
                hf.create_dataset(name, data=val)
    else:
        from . import h5min
        np.savez(dirpath + 'results.npz', **{k: np.asarray(v) for k, v in datasets.items()})
        h5min.write(dirpath + 'results.h5', datasets)


def read_results(dirpath):
    """{dataset name: ndarray} of one k's ``results.h5`` / ``results.npz`` (pyDNMFk.py:278, plot_results.py:117)."""
    h5 = _h5py()
    path = os.path.join(dirpath, 'results.h5')
    if os.path.exists(path):
        if h5 is not None:
            with h5.File(path, 'r') as hf:
                return {k: np.array(hf[k]) for k in hf.keys()}
        from . import h5min
        try:
            return h5min.read(path)
        except (ValueError, struct.error):       # a file of a newer format version than the minimal reader covers
            if not os.path.exists(os.path.join(dirpath, 'results.npz')):
                raise
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
                H_data = np.hstack([H_data[r] for r in transform_H_index(self.p_grid).rankidx2blkidx()])
            else:
                H_data = np.hstack(H_data)
        self.W, self.H = W_data, H_data
        return W_data, H_data
