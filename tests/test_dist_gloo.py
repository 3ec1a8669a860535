"""N>1 host-side logic on CPU: world_size 2 and 4 over gloo (no GPU, no compute kernels)."""
import numpy as np
import pytest

from oracle import cases as C
from oracle import nmf_oracle as O
from tests import common as T
from tests import mp_util, workers


@pytest.mark.parametrize('grid', [(2, 1), (1, 2), (2, 2)])
def test_collectives_and_subcommunicators(grid):
    p_r, p_c = grid
    world = p_r * p_c
    res = mp_util.run(world, workers.comm_worker, (p_r, p_c))
    vg = O.VGrid(p_r, p_c)
    for r, o in enumerate(res):
        i, j = divmod(r, p_c)
        assert o['coord'] == [i, j]
        # naming trap (dist_comm.py:34,48): "row" comm spans a grid column, "column" comm a grid row
        assert o['row_ranks'] == vg.row[j] and o['col_ranks'] == vg.col[i]
        assert o['row_rank'] == i and o['col_rank'] == j
        assert o['sum_int'] == world * (world + 1) // 2
        assert np.array_equal(o['sum_arr'], np.full((2, 3), world * (world + 1) / 2, dtype=np.float32))
        assert o['sum_arr'].dtype == np.float32
        assert o['row_sum'][0] == sum(10.0 * ii + j for ii in range(p_r))
        assert o['col_sum'][0] == sum(10.0 * i + jj for jj in range(p_c))
        assert o['gather'] == [('r', q) for q in range(world)]
        assert o['bcast'] == {'from': 0}
        assert np.array_equal(o['t_allreduce'], np.full(4, sum(range(world)), dtype=np.float64))
        assert np.array_equal(o['t_gather'], np.concatenate([np.full((2, 3), float(q)) for q in vg.col[i]]))
        assert np.array_equal(o['t_gather_ragged'],
                              np.concatenate([np.full((qi + 1, 2), float(q)) for qi, q in enumerate(vg.row[j])]))
        tot = sum(q + 1 for q in vg.row[j])
        full = np.arange(float(p_r * 2 * 3)).reshape(p_r * 2, 3) * tot
        assert np.array_equal(o['t_rs'], full[2 * i:2 * i + 2])
        rag = [1 + q for q in range(p_c)]
        tot2 = sum(q + 1 for q in vg.col[i])
        full2 = np.arange(float(sum(rag) * 2)).reshape(sum(rag), 2) * tot2
        off = sum(rag[:j])
        assert np.array_equal(o['t_rs_ragged'], full2[off:off + rag[j]])
        assert np.array_equal(o['t_bcast'], np.full(3, 5.0))
        assert np.array_equal(o['Reduce_scatter'], (np.arange(float(world * 3)) * (world * (world + 1) / 2))[3 * r:3 * r + 3])
        assert np.array_equal(o['Bcast'], np.full(4, float(world - 1)))


GEOM_CASES = ['u64x48k4_2x1_fro_mu_i1_32', 'u64x48k4_1x2_fro_mu_i1_32', 'u64x48k4_2x2_fro_mu_i1_32',
              'ragged26x14k3_2x2_fro_mu', 'ragged26x14k3_3x1_fro_mu', 'u64x48k4_4x2_fro_mu_i1_32']


@pytest.mark.parametrize('name', GEOM_CASES)
def test_shard_geometry_and_init_rng_order(name):
    """params side effects of data_operations equal the reference's (golden), and the rand init draws
    equal the oracle's replay of the reference order (rank 0 draws the replicated factor last)."""
    case = C.CASES_BY_NAME[name]
    p_r, p_c = case['grid']
    world = p_r * p_c
    res = mp_util.run(world, workers.dims_worker, (case,))
    gold = T.golden_case(name)
    A, blocks, rngs, _ = T.oracle_inputs(case)
    grid = O.VGrid(p_r, p_c)
    sh = O.compute_dims(blocks, grid, case['k'])
    W0, H0 = O.init_factors_rand(blocks, grid, sh, case['k'], rngs)
    for r in range(world):
        assert res[r]['geom'] == [int(v) for v in gold[r]['geom']]
        assert np.array_equal(res[r]['W'], W0[r]) and np.array_equal(res[r]['H'], H0[r])
