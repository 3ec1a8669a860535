"""Where the end-to-end time of PyNMF(A_host).fit() goes (cfg2 shapes, one process per GPU under torchrun):
constructor (H2D of the shard, init, dims), first eager step, graph capture + replayed steps, normalise + error + D2H."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pydnmfk_b200.dist_comm import MPI, MPI_comm  # noqa: E402
from pydnmfk_b200.pyDNMF import PyNMF  # noqa: E402
from pydnmfk_b200.utils import parse  # noqa: E402

rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
m, n, k, itr = 65536, 65536, 32, int(os.environ.get('ITR', '30'))
comm = MPI.COMM_WORLD
comms = MPI_comm(comm, world, 1)
m_i = m // world
host = torch.empty((m_i, n), dtype=torch.float32, pin_memory=True)
host.uniform_(0, 1)
A_host = host.numpy()


def sync():
    torch.cuda.synchronize()
    comm.barrier()


for norm in ('fro', 'kl', 'fro', 'kl'):
    p = parse()
    p.comm1, p.comm, p.row_comm, p.col_comm = comm, comms, comms.cart_1d_row(), comms.cart_1d_column()
    p.p_r, p.p_c, p.k, p.m, p.n, p.itr, p.init, p.verbose = world, 1, k, m, n, itr, 'rand', False
    p.norm, p.method, p.prune, p.W_update = norm, 'mu', False, True
    np.random.seed(7 + rank)
    sync()
    t0 = time.perf_counter()
    nmf = PyNMF(A_host, params=p)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    nmf._run_loop()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    W, H, err = nmf._finish()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    if rank == 0:
        print(json.dumps({'norm': norm, 'world': world, 'itr': itr, 'ctor_s': t1 - t0, 'h2d_GBps_if_all_copy': A_host.nbytes / (t1 - t0) / 1e9,
                          'loop_s': t2 - t1, 'finish_s': t3 - t2, 'total_s': t3 - t0}), flush=True)
comm.barrier()
os._exit(0)
