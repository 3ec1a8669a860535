"""The reference's own MPI test scenarios, restated through the `pyDNMFk` import alias (drop-in check): same data recipes,
same calls, the reference's own pass thresholds (tests/test_dist_nmf_1d.py:39, test_dist_nmf_2d.py, test_dist_nmf_1d_nnsvd_init.py:40,
test_dist_utils.py:49-50).  2 ranks on cuda:0 over gloo."""
import pytest

from tests import mp_util, workers

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def suite():
    return mp_util.run(2, workers.reference_suite_worker, (), backend='gloo', timeout=1200)


def test_dist_nmf_1d(suite):
    for o in suite:
        for grid in ('nmf_1x2', 'nmf_2x1'):
            for combo, err in o[grid].items():
                assert err < 1e-3, (grid, combo, err)


def test_dist_nmf_2d(suite):
    for o in suite:
        for combo, err in o['nmf2d_2x1'].items():
            assert err < 1e-4, (combo, err)


def test_dist_nmf_1d_nnsvd_init(suite):
    for o in suite:
        for grid in ('nnsvd_2x1', 'nnsvd_1x2'):
            for combo, err in o[grid].items():
                assert err < 1e-1, (grid, combo, err)


def test_dist_prune_unprune_1d(suite):
    for o in suite:
        for grid in ('prune_1x2', 'prune_2x1'):
            r = o[grid]
            assert tuple(r['orig'][0]) == tuple(r['unpruned'][0]) and tuple(r['orig'][1]) == tuple(r['unpruned'][1])
            assert r['pruned'][1][0] < r['orig'][0][0] or r['pruned'][2][1] < r['orig'][1][1]
