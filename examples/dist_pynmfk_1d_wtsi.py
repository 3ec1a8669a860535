# torchrun --standalone --local-addr 127.0.0.1 --nproc-per-node 4 examples/dist_pynmfk_1d_wtsi.py
# (the reference: mpirun -n 4 python dist_pynmfk_1d_wtsi.py)
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import pyDNMFk.config as config  # noqa: E402

config.init(0)
from pyDNMFk.pyDNMFk import *    # noqa: E402,F401,F403
from pyDNMFk.utils import *      # noqa: E402,F401,F403
from pyDNMFk.dist_comm import *  # noqa: E402,F401,F403
from pyDNMFk.data_io import data_read  # noqa: E402


def dist_nmfk_1d_nnsvd_init_wtsi():
    comm = MPI.COMM_WORLD
    p_r, p_c = 4, 1
    comms = MPI_comm(comm, p_r, p_c)
    args = parse()
    args.size, args.rank, args.comm, args.p_r, args.p_c = comm.size, comm.rank, comms, p_r, p_c
    args.row_comm, args.col_comm, args.comm1 = comms.cart_1d_row(), comms.cart_1d_column(), comms.comm
    args.fpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden') + '/'
    args.fname, args.ftype = 'wtsi_X', 'npy'
    args.start_k, args.end_k, args.step_k, args.sill_thr = 1, 8, 1, 0.6
    args.itr, args.init, args.verbose, args.norm, args.method = 1000, 'nnsvd', True, 'fro', 'mu'
    args.precision, args.checkpoint = np.float32, False
    A_ij = data_read(args).read().astype(args.precision)
    args.results_path = 'results/'
    nopt = PyNMFk(A_ij, factors=None, params=args).fit()
    assert nopt == 4          # the reference's expected rank for this data set


dist_nmfk_1d_nnsvd_init_wtsi()
