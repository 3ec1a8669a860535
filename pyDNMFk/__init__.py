"""Drop-in import alias: ``from pyDNMFk.pyDNMF import *`` etc. resolve to the B200-native package
``pydnmfk_b200`` (same module names as the reference's ``pyDNMFk`` package for the update-loop path)."""
import importlib
import sys

_MODULES = ('config', 'utils', 'dist_comm', 'dist_nmf', 'data_io', 'pyDNMF', 'dist_clustering', 'dist_svd', 'pyDNMFk', 'runner')
for _m in _MODULES:
    _mod = importlib.import_module('pydnmfk_b200.' + _m)
    sys.modules[__name__ + '.' + _m] = _mod
    globals()[_m] = _mod
