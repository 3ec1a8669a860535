#!/usr/bin/env python
"""Benchmark of the hot path: MU iterations/s (FRO and KL, k=32, 65536 x 65536 fp32) on N B200s.

    python bench.py --gpus N --steps K --warmup W            # this implementation
    python bench.py --impl reference --steps K --warmup W     # the CPU path of the reference's algorithm

A "step" is one MU iteration = the body of the reference's fit loop (pyDNMF.py:151-172):
``update()`` + the every-10th-iteration clamp.  For each norm W warm-up steps run untimed, then
exactly K steps are timed between barrier + synchronize, with CUDA events on the launching stream;
the maximum over ranks is taken.  ``value`` = total timed iterations / total time over both norms
(the per-norm numbers are in ``by_norm``).  The data shard is far larger than L2 (A = 16 GiB per
pass), so no extra L2 flush is needed between steps.

Workload (BASELINE.json configs[1]): synthetic i.i.d. uniform(0,1) fp32 matrix, 65536 x 65536, k=32,
grid N x 1 (row shards; the matrix is fixed as N grows => strong scaling), prune off, rand init.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'MU iters/s (FRO & KL, k=32, 65536^2 fp32)'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--m', type=int, default=65536)
    ap.add_argument('--n', type=int, default=65536)
    ap.add_argument('--k', type=int, default=32)
    ap.add_argument('--norms', default='fro,kl')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-sample', type=int, default=8192, help='side of the square CPU-baseline sample')
    ap.add_argument('--force-generic', action='store_true', help='disable the tcgen05 path (A/B runs)')
    ap.add_argument('--no-graph', action='store_true', help='launch every step eagerly instead of replaying a CUDA graph')
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port (numpy restatement of the reference's loop) on the host cores
# ------------------------------------------------------------------------------------------------
def _cpu_rank_work(q, barrier, shard_rows, side, k, norms, steps, warmup, seed):
    """One forked 'rank' of a C x 1 row grid: the oracle's update on its own shard, single BLAS thread
    (the reference forces OMP_NUM_THREADS=1 per MPI rank, main.py:3)."""
    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(1)
    except Exception:
        limiter = None
    from oracle import nmf_oracle as O
    rs = np.random.RandomState(seed)
    A = rs.rand(shard_rows, side).astype(np.float32)
    res = {}
    for norm in norms:
        st = O._State()
        st.grid, st.k, st.norm, st.method, st.W_update = O.VGrid(1, 1), k, norm, 'mu', True
        st.A, st.dt, st.eps, st.topo = [A], A.dtype, np.finfo(np.float32).eps, '1d'
        st.W = [rs.rand(shard_rows, k).astype(np.float32)]
        st.H = [rs.rand(k, side).astype(np.float32)]
        for i in range(warmup):
            O.update(st, 1)
        barrier.wait()
        t0 = time.perf_counter()
        for i in range(steps):
            O.update(st, 1)
            if i % 10 == 0:
                st.H = [np.maximum(h, st.eps) for h in st.H]
                st.W = [np.maximum(w, st.eps) for w in st.W]
        res[norm] = time.perf_counter() - t0
        barrier.wait()
    q.put(res)
    del limiter


def cpu_baseline(args, steps, warmup):
    """Times oracle.nmf_oracle (kind "port") the way the reference runs on a CPU: C = host cores forked
    ranks of a C x 1 row grid, one BLAS thread each, every rank updating its own row shard of a bounded
    square sample concurrently (the k x n all-reduce between ranks is not included: <1 % of the
    reference's time).  Time = slowest rank; iterations/s are scaled to the full
    workload by the element ratio (the loop is O(m n k))."""
    import multiprocessing as mp
    side = min(args.cpu_sample, args.m, args.n)
    k = args.k
    cores = os.cpu_count() or 1
    shard = max(1, side // cores)
    side_rows = shard * cores
    scale = (side_rows * side) / float(args.m * args.n)
    norms = args.norms.split(',')
    ctx = mp.get_context('fork')
    q = ctx.Queue()
    barrier = ctx.Barrier(cores)
    procs = [ctx.Process(target=_cpu_rank_work, args=(q, barrier, shard, side, k, norms, steps, warmup, 1234 + r))
             for r in range(cores)]
    for p in procs:
        p.start()
    results = [q.get() for _ in procs]
    for p in procs:
        p.join()
    out = {}
    tot_t = 0.0
    for norm in norms:
        dt = max(r[norm] for r in results)
        out[norm] = steps / dt * scale
        tot_t += dt
    tot_it = steps * len(norms)
    return dict(value=tot_it / tot_t * scale, unit='iters/s', cores=int(cores), kind='port',
                sample='oracle/nmf_oracle.py (numpy %s + OpenBLAS, 1 BLAS thread per rank) as %d forked ranks of a %dx1 '
                       'grid on a %dx%d k=%d fp32 sample (%d rows per rank), %d timed iterations per norm, slowest rank; '
                       'iterations/s scaled by the element ratio %.5f to the %dx%d workload'
                       % (np.__version__, cores, cores, side_rows, side, k, shard, steps, scale, args.m, args.n),
                by_norm=out)


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 10))
    warmup = max(1, min(args.warmup, 2))
    t0 = time.perf_counter()
    cb = cpu_baseline(args, steps, warmup)
    wall = time.perf_counter() - t0
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': cb['value'], 'unit': 'iters/s', 'n_gpus': args.gpus,
        'steps': steps, 'warmup': warmup, 'ms_per_step': 1000.0 / cb['value'] if cb['value'] > 0 else None,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'synthetic %dx%d fp32 non-negative, KL and FRO MU k=%d (CPU sample, see cpu_baseline.sample)'
                               % (args.m, args.n, args.k)},
        'cpu_baseline': cb,
        'e2e': {'value': cb['value'], 'unit': 'iters/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0, 'by_norm': cb['by_norm'], 'wall_s': wall,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                pass

    def summary(self, windows):
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for t, line in self.rows:
            if not any(a <= t <= b for a, b in windows):
                continue
            f = [x.strip() for x in line.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons),
                'samples': len(sm)}


# ------------------------------------------------------------------------------------------------
# this implementation
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from pydnmfk_b200 import _lib as L
    from pydnmfk_b200 import device as D
    from pydnmfk_b200.dist_comm import MPI, MPI_comm
    from pydnmfk_b200.dist_nmf import nmf_algorithms_1D
    from pydnmfk_b200.graphs import StepGraphs, graphs_enabled
    from pydnmfk_b200.pyDNMF import PyNMF
    from pydnmfk_b200.utils import parse, determine_block_params

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    # CPU baseline first (forks workers; done before this process creates a CUDA context)
    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline(args, steps=3, warmup=1)
    assert torch.cuda.is_available(), 'bench.py needs a GPU (no CPU fallback)'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if args.force_generic:
        L.set_force_generic(True)
    comm = MPI.COMM_WORLD          # initialises NCCL from the torchrun environment when world > 1
    assert comm.size == world
    p_r, p_c = world, 1
    comms = MPI_comm(comm, p_r, p_c)
    m, n, k = args.m, args.n, args.k
    blk = determine_block_params(rank, (p_r, p_c), (m, n))
    (r0, _), (r1, _) = blk.determine_block_index_range_asymm()
    m_loc = r1 - r0 + 1
    eps = float(np.finfo(np.float32).eps)

    def make_params(norm, itr):
        p = parse()
        p.comm1, p.comm, p.row_comm, p.col_comm = comm, comms, comms.cart_1d_row(), comms.cart_1d_column()
        p.p_r, p.p_c, p.k, p.m, p.n, p.itr, p.init, p.verbose = p_r, p_c, k, m, n, itr, 'rand', False
        p.norm, p.method, p.prune, p.W_update, p.eps = norm, 'mu', False, True, np.finfo(np.float32).eps
        p.rank = rank
        return p

    # synthetic shard, generated on the device (Philox seed 1234 + rank), strictly positive
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    A = torch.rand((m_loc, n), generator=g, device=dev, dtype=torch.float32)
    g.manual_seed(7)
    W0 = torch.rand((m_loc, k), generator=g, device=dev, dtype=torch.float32)
    g.manual_seed(7)
    H0 = torch.rand((k, n), generator=g, device=dev, dtype=torch.float32)
    ops = D.default_ops()

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    windows = []
    by_norm, pass_stats = {}, {}
    launches = 0
    tot_ms, tot_it = 0.0, 0
    paths = {}
    for norm in args.norms.split(','):
        p = make_params(norm, args.steps)
        W, H = W0.clone(), H0.clone()
        alg = nmf_algorithms_1D(A, W, H, params=p)

        def clamp():
            ops.clamp_min(H, eps)
            ops.clamp_min(W, eps)

        sg = StepGraphs(alg.update, clamp) if (graphs_enabled(comm, 'mu') and not args.no_graph) else None

        def step(i):
            # the product's own loop body (PyNMF.fit): CUDA-graph replay of update() [+ clamp every 10th iteration]
            if sg is not None:
                sg.clamped() if i % 10 == 0 else sg.plain()
            else:
                alg.update()
                if i % 10 == 0:
                    clamp()

        L.launch_count(reset=True)
        alg.update()                       # eager step: sizes the workspace, counts the launches of one step
        clamp()
        launches_per_step = L.launch_count()
        for i in range(args.warmup):
            step(i)
        sync_all()
        L.launch_count(reset=True)
        ops.timers = {} if sg is None else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_a = time.perf_counter()
        e0.record()
        for i in range(args.steps):
            step(i)
        e1.record()
        sync_all()
        t_b = time.perf_counter()
        windows.append((t_a, t_b))
        ms = e0.elapsed_time(e1)
        launches += launches_per_step * args.steps     # same launches per replayed step as in the eager one
        paths[norm] = 'tcgen05' if L.last_path() == 1 else 'generic'
        if sg is not None:
            # per-kernel CUDA events cannot be placed inside a replayed graph: time the same steps once more, eagerly,
            # with an event pair around every A-streaming pass (not part of `value`)
            ops.timers = {}
            for i in range(min(args.steps, 10)):
                alg.update()
                if i % 10 == 0:
                    clamp()
            sync_all()
        pass_stats[norm] = ops.timer_summary()
        ops.timers = None
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        by_norm[norm] = {'iters_per_s': args.steps / (ms / 1000.0), 'ms_per_step': ms / args.steps}
        tot_ms += ms
        tot_it += args.steps
        assert torch.isfinite(W).all() and torch.isfinite(H).all()
        del alg, sg, step, clamp      # captured graphs (with their NCCL nodes) must die before the process group does
        gc.collect()
        torch.cuda.synchronize()

    value = tot_it / (tot_ms / 1000.0)

    # ---- roofline of the dominant kernel (A-streaming pass), measured live with CUDA events ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s'
    pass_bytes = float(m_loc) * n * 4.0
    per_kernel = {}
    dom = None
    for norm, st in pass_stats.items():
        for opn, (cnt, mean_ms) in st.items():
            gbs = pass_bytes / (mean_ms * 1e-3) / 1e9
            per_kernel['%s:%s' % (norm, opn)] = {'launches': cnt, 'mean_ms': mean_ms, 'GBps': gbs, 'frac': gbs / peak}
            if dom is None or cnt * mean_ms > dom[1]:
                dom = ('%s:%s' % (norm, opn), cnt * mean_ms, gbs)
    # DRAM traffic of the dominant kernel from the committed ncu --set full capture (same shard shape only)
    traffic = None
    try:
        cap = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')))
        if dom and (cap['m_loc'], cap['n'], cap['k']) == (m_loc, n, k) and dom[0] in cap['kernels']:
            traffic = cap['kernels'][dom[0]]['dram_bytes_per_launch']
    except Exception:
        pass
    roofline = {'bound': 'hbm', 'achieved': dom[2] if dom else None, 'peak': peak, 'unit': 'GB/s',
                'frac': (dom[2] / peak) if dom else None, 'traffic': traffic,
                'traffic_source': 'profiles/ncu_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum per launch)' if traffic else None, 'kernel': dom[0] if dom else None,
                'peak_source': peak_src, 'algorithmic_bytes_per_launch': pass_bytes, 'per_kernel': per_kernel,
                'iteration_frac_of_A_streaming_roofline': {
                    nm: (2.0 * pass_bytes / (v['ms_per_step'] * 1e-3) / 1e9) / peak for nm, v in by_norm.items()}}

    # ---- end to end through the public API with HOST buffers (PyNMF(A_host).fit()) ----------------
    e2e = None
    if not args.no_e2e:
        try:
            host = torch.empty((m_loc, n), dtype=torch.float32, pin_memory=True)
            host.copy_(A)
            torch.cuda.synchronize()
            A_host = host.numpy()
            del A
            torch.cuda.empty_cache()
            e_t, e_it, h2d, d2h = 0.0, 0, 0, 0
            for norm in args.norms.split(','):
                p = make_params(norm, args.steps)
                np.random.seed(7 + rank)
                sync_all()
                t0 = time.perf_counter()
                Wn, Hn, err = PyNMF(A_host, params=p).fit()
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                if world > 1:
                    t = torch.tensor([dt], device=dev, dtype=torch.float64)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    dt = float(t.item())
                e_t += dt
                e_it += args.steps
                h2d += A_host.nbytes + (m_loc * k + k * n) * 4
                d2h += Wn.nbytes + Hn.nbytes + 8
                assert np.isfinite(float(err))
            e2e = {'value': e_it / e_t, 'unit': 'iters/s', 'h2d_bytes_per_step': h2d * world / e_it,
                   'd2h_bytes_per_step': d2h * world / e_it,
                   'note': 'PyNMF(A_host, params).fit() with itr=%d per norm: pinned host shard -> HBM copy, rand init on '
                           'host, %d iterations, normalise + relative error, factors and error back to host; bytes are '
                           'the whole-fit transfers divided by the iterations' % (args.steps, args.steps)}
        except Exception as ex:  # pinned host memory may be unavailable on a small host
            e2e = {'value': None, 'unit': 'iters/s', 'h2d_bytes_per_step': None, 'd2h_bytes_per_step': None,
                   'note': 'e2e leg failed: %r' % (ex,)}

    clocks = None
    if sampler is not None:
        sampler.stop()
        clocks = sampler.summary(windows)

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': 'iters/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': tot_ms / tot_it, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'synthetic %dx%d fp32 non-negative, KL and FRO MU k=%d on %d B200 (grid %dx1, '
                                   'row shards of %d rows)' % (m, n, k, world, world, m_loc),
                       'l2': 'inputs larger than L2 (one A pass streams %.1f GiB per GPU)' % (pass_bytes / 2 ** 30),
                       'norms': args.norms, 'math_mode': 'fp32-accurate', 'paths': paths,
                       'launch': 'cuda-graph replay of update()+clamp (as PyNMF.fit does)' if not args.no_graph else 'eager',
                       'roofline_timing': 'CUDA events around each A-streaming pass, eager replica of the timed steps' if not args.no_graph else 'CUDA events around each A-streaming pass inside the timed steps'},
            'by_norm': by_norm, 'roofline': roofline, 'cpu_baseline': cb, 'e2e': e2e, 'gpu_launches': int(launches),
            'clocks': clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # Leave without tearing NCCL down: destroy_process_group() can block for minutes once collectives have been
        # captured into CUDA graphs, and the JSON line is already out.  Every rank exits 0 right after a final barrier.
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
