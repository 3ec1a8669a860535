"""Secondary configurations of BASELINE.json / SURVEY.md section 8d, measured through the public API on ONE GPU
(bench.py stays on the headline configuration, cfg2).  One JSON line per configuration:

  cfg1   1024 x 256  k=4   FRO-MU and KL-MU, itr iterations (swim-sized; the whole fit runs on one thread-block cluster)
  cfg4   131072 x 65536 k=16 FRO-HALS and FRO-BCD (32 GiB shard; 2 resp. 3 A passes per iteration)
  cfg5   96 x 21 (wtsi-sized) NMFk ensemble: 20 perturbations x KL-MU itr=1000 for k = 2..10, wall time

Per-iteration time = (T(itr = hi) - T(itr = lo)) / (hi - lo) so the init / normalise / error epilogue cancels.
    python tools/bench_configs.py [--only cfg1,cfg4,cfg5]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pydnmfk_b200.dist_comm import MPI, MPI_comm  # noqa: E402
from pydnmfk_b200.pyDNMF import PyNMF  # noqa: E402
from pydnmfk_b200.pyDNMFk import PyNMFk  # noqa: E402
from pydnmfk_b200.utils import parse  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--only', default='cfg1,cfg4,cfg5')
ap.add_argument('--cfg4-rows', type=int, default=131072)
args = ap.parse_args()
only = set(args.only.split(','))
comm = MPI.COMM_WORLD
comms = MPI_comm(comm, 1, 1)
PEAK = 6547.2
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs'])
except Exception:
    pass


def params(k, norm, method, itr, m, n):
    p = parse()
    p.comm1, p.comm, p.row_comm, p.col_comm = comm, comms, comms.cart_1d_row(), comms.cart_1d_column()
    p.p_r, p.p_c, p.k, p.m, p.n, p.itr, p.init, p.verbose = 1, 1, k, m, n, itr, 'rand', False
    p.norm, p.method, p.prune, p.W_update = norm, method, False, True
    return p


def fit_seconds(A, k, norm, method, itr):
    np.random.seed(7)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    W, H, err = PyNMF(A, params=params(k, norm, method, itr, A.shape[0], A.shape[1])).fit()
    torch.cuda.synchronize()
    return time.perf_counter() - t0, float(err)


def per_iteration(A, k, norm, method, lo, hi):
    fit_seconds(A, k, norm, method, hi)              # warm-up on the same code path (kernel attributes, calibration,
                                                     # first CUDA-graph capture / memory pool at this shape)
    t_lo, _ = fit_seconds(A, k, norm, method, lo)
    t_hi, err = fit_seconds(A, k, norm, method, hi)
    return (t_hi - t_lo) / (hi - lo), err


if 'cfg1' in only:
    g = torch.Generator(device='cuda')
    g.manual_seed(1234)
    A = torch.rand((1024, 256), generator=g, device='cuda')
    out = {}
    for norm in ('fro', 'kl'):
        dt, err = per_iteration(A, 4, norm, 'mu', 200, 2200)
        out[norm] = {'us_per_iteration': dt * 1e6, 'iters_per_s': 1.0 / dt, 'recon_err': err}
    print(json.dumps({'config': 'cfg1: 1024x256 fp32 k=4 MU, 1 GPU, PyNMF.fit (whole fit on a 16-CTA thread-block cluster)', 'by_norm': out}), flush=True)

if 'cfg4' in only:
    m, n, k = args.cfg4_rows, 65536, 16
    g = torch.Generator(device='cuda')
    g.manual_seed(1234)
    A = torch.rand((m, n), generator=g, device='cuda')
    out = {}
    for method, passes in (('hals', 2), ('bcd', 3)):
        dt, err = per_iteration(A, k, 'fro', method, 5, 15)
        gbs = passes * m * n * 4 / dt / 1e9
        out[method] = {'ms_per_iteration': dt * 1e3, 'iters_per_s': 1.0 / dt, 'A_passes_per_iteration': passes,
                       'GBps': gbs, 'frac_of_hbm_peak': gbs / PEAK, 'recon_err': err}
    print(json.dumps({'config': 'cfg4: %dx%d fp32 k=%d FRO-HALS / FRO-BCD, 1 GPU, PyNMF.fit' % (m, n, k), 'peak_GBps': PEAK,
                      'by_method': out}), flush=True)
    del A
    torch.cuda.empty_cache()

if 'cfg5' in only:
    import tempfile
    A = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'wtsi_X.npy')).astype(np.float32)
    with tempfile.TemporaryDirectory() as tmp:
        def nmfk(end_k, itr):
            p = params(2, 'kl', 'mu', itr, 96, 21)
            p.perturbations, p.noise_var, p.sampling, p.prune, p.checkpoint = 20, 0.015, 'uniform', True, False
            p.start_k, p.end_k, p.step_k, p.sill_thr, p.fname = 2, end_k, 1, 0.9, 'wtsi_%d_%d' % (end_k, itr)
            p.results_path = tmp + '/'
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            nopt = PyNMFk(A, params=p).fit()
            torch.cuda.synchronize()
            return time.perf_counter() - t0, nopt
        nmfk(2, 50)                                        # warm-up
        dt, nopt = nmfk(10, 1000)
    fits = 9 * 20
    print(json.dumps({'config': 'cfg5: wtsi 96x21 fp32, PyNMFk.fit k=2..10, 20 perturbations, KL-MU itr=1000, rand init, '
                                'sill_thr 0.9, 1 GPU (ensemble + clustering + silhouettes + regression + rank selection)',
                      'nmfk_wall_s': dt, 'nopt': int(nopt), 'perturbation_fits': fits, 'ms_per_fit_incl_clustering': dt / fits * 1e3}),
          flush=True)
