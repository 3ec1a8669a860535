"""Round-2 bring-up check for the opt-in tensor-core residual (tc_kl_kernel<2>, DNMF_TC_RESIDUAL=1), which was written at
the end of round 1 without GPU time left to run it.  NOT part of the test-suite on purpose (run it under `timeout`):

    DNMF_TC_RESIDUAL=1 timeout 120 python tools/check_tc_residual.py

Compares dnmf_residual_sqnorm on the tcgen05 pipeline with float64 numpy and with the CUDA-core kernel, then times both
at the cfg4 shard shape."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
assert os.environ.get('DNMF_TC_RESIDUAL') == '1', 'set DNMF_TC_RESIDUAL=1'
from pydnmfk_b200 import _lib as L  # noqa: E402
from pydnmfk_b200 import device as D  # noqa: E402

ops = D.default_ops()
L.set_tc_min_elems(1)
rs = np.random.RandomState(0)
for m, n, k in ((128, 32, 32), (256, 96, 16), (1000, 1000, 10), (515, 2052, 32), (4100, 300, 7), (2048, 2048, 32)):
    A = rs.rand(m, n).astype(np.float32)
    W = rs.rand(m, k).astype(np.float32)
    H = (rs.rand(k, n) / k).astype(np.float32)
    ref = [np.sum((A.astype(np.float64) - W.astype(np.float64) @ H.astype(np.float64)) ** 2), np.sum(A.astype(np.float64) ** 2)]
    dev = lambda x: torch.from_numpy(x).cuda()   # noqa: E731
    got = ops.residual_sqnorm(dev(A), dev(W), dev(H)).cpu().numpy()
    path = L.last_path()
    L.set_force_generic(True)
    gen = ops.residual_sqnorm(dev(A), dev(W), dev(H)).cpu().numpy()
    L.set_force_generic(False)
    print('%5d x %5d k=%2d path=%d  rel diff vs fp64: residual %.2e norm %.2e   (generic kernel: %.2e %.2e)'
          % (m, n, k, path, abs(got[0] - ref[0]) / ref[0], abs(got[1] - ref[1]) / ref[1], abs(gen[0] - ref[0]) / ref[0],
             abs(gen[1] - ref[1]) / ref[1]), flush=True)
    assert path == 1 and abs(got[0] - ref[0]) <= 1e-5 * ref[0] and abs(got[1] - ref[1]) <= 1e-6 * ref[1]
m, n, k = 65536, 65536, 16
A = torch.rand((m, n), device='cuda')
W = torch.rand((m, k), device='cuda')
H = torch.rand((k, n), device='cuda') / k
for name, force in (('tcgen05', False), ('generic', True)):
    L.set_force_generic(force)
    ops.residual_sqnorm(A, W, H)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ops.residual_sqnorm(A, W, H)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print('%s residual pass at %d x %d k=%d: %.3f ms (%.0f GB/s)' % (name, m, n, k, ms, m * n * 4 / ms / 1e6))
L.set_force_generic(False)
print('ok')
