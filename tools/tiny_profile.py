"""Launch-bound regime (cfg1 / cfg5 shapes): GPU time vs host overhead of one MU iteration.
    python tools/tiny_profile.py --m 96 --n 21 --k 4 --norm kl
    ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/tiny.csv python tools/tiny_profile.py ... --eager-only
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pydnmfk_b200 import device as D  # noqa: E402
from pydnmfk_b200.dist_comm import MPI, MPI_comm  # noqa: E402
from pydnmfk_b200.dist_nmf import nmf_algorithms_1D  # noqa: E402
from pydnmfk_b200.graphs import StepGraphs  # noqa: E402
from pydnmfk_b200.utils import parse  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--m', type=int, default=96)
ap.add_argument('--n', type=int, default=21)
ap.add_argument('--k', type=int, default=4)
ap.add_argument('--norm', default='kl')
ap.add_argument('--eager-only', action='store_true')
a = ap.parse_args()
comm = MPI.COMM_WORLD
comms = MPI_comm(comm, 1, 1)
p = parse()
p.comm1, p.comm, p.row_comm, p.col_comm = comm, comms, comms.cart_1d_row(), comms.cart_1d_column()
p.p_r, p.p_c, p.k, p.m, p.n, p.itr, p.init, p.verbose = 1, 1, a.k, a.m, a.n, 10, 'rand', False
p.norm, p.method, p.prune, p.W_update, p.eps = a.norm, 'mu', False, True, np.finfo(np.float32).eps
A = torch.rand((a.m, a.n), device='cuda')
W = torch.rand((a.m, a.k), device='cuda')
H = torch.rand((a.k, a.n), device='cuda')
alg = nmf_algorithms_1D(A, W, H, params=p)
ops = D.default_ops()
for _ in range(3):
    alg.update()
torch.cuda.synchronize()
if a.eager_only:
    sys.exit(0)
sg = StepGraphs(alg.update, lambda: (ops.clamp_min(H, 1e-7), ops.clamp_min(W, 1e-7)))
sg.plain()
torch.cuda.synchronize()
N = 500
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(N):
    sg.plain()
e1.record()
t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
t_wall = time.perf_counter() - t0
print('%dx%d k=%d %s: graph replay  device %.1f us/it, host issue %.1f us/it, wall %.1f us/it'
      % (a.m, a.n, a.k, a.norm, e0.elapsed_time(e1) * 1e3 / N, t_issue * 1e6 / N, t_wall * 1e6 / N))
t0 = time.perf_counter()
for _ in range(N):
    alg.update()
torch.cuda.synchronize()
print('eager: wall %.1f us/it' % ((time.perf_counter() - t0) * 1e6 / N))
