"""Minimal ``h5py`` stand-in: a pickle-backed ``File`` with ``create_dataset`` / item access, enough for the
reference's results.h5 round trip (data_io.py:202-209 writes, pyDNMFk.py:278 and plot_results.py reads).
TEST INFRASTRUCTURE ONLY."""
import os
import pickle

import numpy as np


class File:
    def __init__(self, name, mode='r', **kw):
        self._name, self._mode = name, mode
        self._data = {}
        if mode in ('r', 'r+', 'a') and os.path.exists(name):
            with open(name, 'rb') as f:
                self._data = pickle.load(f)
        elif mode == 'r':
            raise OSError('h5py stand-in: no such file %s' % name)

    def create_dataset(self, name, data=None, **kw):
        self._data[name] = np.asarray(data)
        return self._data[name]

    def __getitem__(self, name):
        return self._data[name]

    def __contains__(self, name):
        return name in self._data

    def keys(self):
        return self._data.keys()

    def close(self):
        if self._mode != 'r':
            with open(self._name, 'wb') as f:
                pickle.dump(self._data, f)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
