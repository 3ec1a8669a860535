"""Call-swallowing ``matplotlib`` stub (plot_results.py:2-3 imports it at module scope).  TEST INFRASTRUCTURE ONLY."""
from unittest.mock import MagicMock

rcParams = MagicMock()


def __getattr__(name):
    if name.startswith('__'):
        raise AttributeError(name)
    return MagicMock()


import importlib  # noqa: E402

pyplot = importlib.import_module('.pyplot', __name__)   # the real stub module, not a mock attribute
