"""Empty ``h5py`` stub: the reference imports it at module scope
(data_io.py:5, plot_results.py:5) but the NMF update loop never calls it.
TEST INFRASTRUCTURE ONLY."""


class File:  # pragma: no cover - never exercised by the hot path
    def __init__(self, *a, **k):
        raise RuntimeError("h5py stub: results.h5 I/O is outside the hot path")
