"""Tiny driver for ncu captures of the A-streaming passes (one GPU, short):
    ncu --set full --clock-control none --import-source on -k regex:tc_pass_kernel -s 4 -c 2 -o gpurun_out/prof \
        python tools/prof_tc.py --m 32768 --n 32768 --k 32
Launch order: calibration (k=16 probe, 1-2 launches), then `--reps` x (ah, wta[, kl_uht, kl_wtu])."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pydnmfk_b200 import device as D  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--m', type=int, default=32768)
ap.add_argument('--n', type=int, default=32768)
ap.add_argument('--k', type=int, default=32)
ap.add_argument('--reps', type=int, default=3)
ap.add_argument('--kl', action='store_true')
ap.add_argument('--roles', action='store_true', help='print the per-warp-role cycle split of the tcgen05 kernel')
ap.add_argument('--dbg', default='0', help='timing-ablation bits (dnmf_set_tc_debug) for the role timers')
a = ap.parse_args()
ops = D.default_ops()


def copy_peak():
    """STREAM-style copy on this box (what MEASURED_PEAKS.json's hbm_gbs is): read + write bytes / time, best of 5."""
    x = torch.empty(1 << 30, dtype=torch.bfloat16, device='cuda')
    y = torch.empty_like(x)
    best = 1e9
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); y.copy_(x); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2 * x.numel() * 2 / best / 1e6


print('copy peak on this box: %.0f GB/s' % copy_peak(), flush=True)
A = torch.rand((a.m, a.n), device='cuda')
H = torch.rand((a.k, a.n), device='cuda')
W = torch.rand((a.m, a.k), device='cuda')
for _ in range(a.reps):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    e[0].record(); ops.ah(A, H); e[1].record(); ops.wta(A, W); e[2].record()
    if a.kl:
        ops.kl_uht(A, W, H, 1.2e-7); e[3].record(); ops.kl_wtu(A, W, H, 1.2e-7); e[4].record()
    torch.cuda.synchronize()
    gb = a.m * a.n * 4 / 1e9
    msg = 'ah %.3f ms (%.0f GB/s)  wta %.3f ms (%.0f GB/s)' % (e[0].elapsed_time(e[1]), gb / e[0].elapsed_time(e[1]) * 1e3,
                                                                e[1].elapsed_time(e[2]), gb / e[1].elapsed_time(e[2]) * 1e3)
    if a.kl:
        msg += '  kl_uht %.3f ms  kl_wtu %.3f ms' % (e[2].elapsed_time(e[3]), e[3].elapsed_time(e[4]))
    print(msg, flush=True)

if a.roles:
    from pydnmfk_b200 import _lib as L
    names = ['A-prod wait a_free', 'A-prod issue', 'MMA wait acce', 'MMA wait t_full', 'MMA wait b_full', 'MMA issue', 'MMA commit',
             'split wait a_full', 'split load+math', 'split wait t_free', 'split store', 'drain wait accf', 'drain work', '', '', 'total']
    for nm, fn in (('ah', lambda: ops.ah(A, H)), ('wta', lambda: ops.wta(A, W))):
        buf = torch.zeros(148 * 16, dtype=torch.int64, device='cuda')
        L.call('dnmf_set_tc_profile', buf.data_ptr())
        fn()
        torch.cuda.synchronize()
        L.call('dnmf_set_tc_profile', None)
        v = buf.cpu().numpy().reshape(148, 16).astype(float)
        tot = v[:, 15].mean()
        print(nm, 'cycles per CTA %.0f' % tot)
        for i, n in enumerate(names):
            if n and i != 15:
                print('   %-22s %5.1f%%' % (n, 100 * v[:, i].mean() / tot))

if a.roles and a.kl:
    from pydnmfk_b200 import _lib as L
    L.call('dnmf_set_tc_debug', int(a.dbg, 0))
    print('role timers with debug flags', a.dbg)
    names = ['MMA gemm1 wait(Fr,S free)', 'MMA gemm1 issue', 'MMA wait acce', 'MMA wait t_full', 'MMA wait b_full', 'MMA gemm2 issue+loop',
             'split wait a_full', 'split load A', 'split wait S', 'split ldtm+div+lo', 'split wait t_free', 'split store', '', '', '', 'total']
    for nm, fn in (('kl_uht', lambda: ops.kl_uht(A, W, H, 1.2e-7)), ('kl_wtu', lambda: ops.kl_wtu(A, W, H, 1.2e-7))):
        buf = torch.zeros(148 * 16, dtype=torch.int64, device='cuda')
        L.call('dnmf_set_tc_profile', buf.data_ptr())
        fn()
        torch.cuda.synchronize()
        L.call('dnmf_set_tc_profile', None)
        v = buf.cpu().numpy().reshape(148, 16).astype(float)
        tot = v[:, 15].mean()
        print(nm, 'cycles per CTA %.0f' % tot)
        for i, n in enumerate(names):
            if n and i != 15:
                print('   %-28s %5.1f%%' % (n, 100 * v[:, i].mean() / tot))
