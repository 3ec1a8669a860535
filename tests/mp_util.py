"""Launch world_size ranks of a module-level function with torch.distributed (spawned processes).

Used by the CPU-side gloo tests and by the GPU parity tests that run several ranks on ONE device
(gloo backend, device tensors staged through host memory by pydnmfk_b200.dist_comm)."""
import os
import pickle
import socket
import sys
import tempfile
import traceback

import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _entry(rank, world, port, backend, fn, args, outdir, env=None):
    os.environ['RANK'] = str(rank)
    os.environ['LOCAL_RANK'] = '0'
    os.environ['WORLD_SIZE'] = str(world)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    os.environ['DNMF_BACKEND'] = backend
    os.environ['OMP_NUM_THREADS'] = '1'
    os.environ.setdefault('DNMF_PG_TIMEOUT_S', '240')
    for key, val in (env or {}).items():
        os.environ[key] = str(val)
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    try:
        res = ('ok', fn(rank, world, *args))
    except BaseException:
        res = ('err', traceback.format_exc())
    with open(os.path.join(outdir, 'r%d.pkl' % rank), 'wb') as f:
        pickle.dump(res, f)
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.barrier()
            dist.destroy_process_group()
    except Exception:
        pass


def run(world, fn, args=(), backend='gloo', timeout=600, env=None):
    """Returns [fn(rank, world, *args) for every rank]; raises if any rank failed."""
    outdir = tempfile.mkdtemp(prefix='dnmf_mp_')
    port = _free_port()
    ctx = mp.start_processes(_entry, args=(world, port, backend, fn, args, outdir, env), nprocs=world, join=False,
                             start_method='spawn')
    ctx.join(timeout)
    for p in ctx.processes:
        if p.is_alive():
            p.terminate()
    out = []
    for r in range(world):
        fn_r = os.path.join(outdir, 'r%d.pkl' % r)
        if not os.path.exists(fn_r):
            raise RuntimeError('rank %d produced no result (crash or timeout)' % r)
        with open(fn_r, 'rb') as f:
            tag, val = pickle.load(f)
        if tag == 'err':
            raise RuntimeError('rank %d failed:\n%s' % (r, val))
        out.append(val)
    return out
