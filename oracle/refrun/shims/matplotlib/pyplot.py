"""Call-swallowing ``matplotlib.pyplot`` (the reference plots the selection curve at the end of PyNMFk.fit,
pyDNMFk.py:210).  TEST INFRASTRUCTURE ONLY."""
from unittest.mock import MagicMock

rcParams = MagicMock()


def subplots(*a, **k):
    return MagicMock(), MagicMock()


def __getattr__(name):
    return MagicMock()
