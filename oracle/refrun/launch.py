"""Launch the UNMODIFIED reference on R forked ranks (oracle side only).

TEST INFRASTRUCTURE: used by ``oracle/gen_golden.py`` (authoring container,
where ``/root/reference`` is mounted) to produce ``tests/golden/*.npz`` and by
CPU-side tests that cross-check the numpy restatement against the reference
when the reference is present.  Never imported by the product package.

Shims applied before importing the reference:
  * ``mpi4py``  -> oracle/refrun/shims/mpi4py  (fork + pipes)
  * ``h5py`` / ``matplotlib`` -> empty stubs (imported at module scope only)
  * ``np.product = np.prod``   (removed in numpy 2; utils.py:33)
"""
import multiprocessing as mp
import os
import sys
import traceback

HERE = os.path.dirname(os.path.abspath(__file__))
SHIMS = os.path.join(HERE, 'shims')
REFERENCE = os.environ.get('DNMF_REFERENCE', '/root/reference')


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE, 'pyDNMFk'))


def _child(rank, size, conns, result_conn, target, args):
    try:
        os.environ["OMP_NUM_THREADS"] = "1"
        # the reference and the shims must win over anything of the same name in the parent process: the repo ships a
        # `pyDNMFk` import alias of the product, which must never answer for the reference here
        for p in (REFERENCE, SHIMS):
            while p in sys.path:
                sys.path.remove(p)
            sys.path.insert(0, p)
        for name in [n for n in sys.modules if n == 'pyDNMFk' or n.startswith('pyDNMFk.')
                     or n.split('.')[0] in ('mpi4py', 'h5py', 'matplotlib')]:
            del sys.modules[name]
        import pyDNMFk as _ref_pkg
        if not os.path.abspath(_ref_pkg.__file__).startswith(os.path.abspath(REFERENCE) + os.sep):
            raise RuntimeError('reference harness imported %s instead of the reference' % _ref_pkg.__file__)
        import numpy as np
        if not hasattr(np, 'product'):
            np.product = np.prod
        import mpi4py
        mpi4py._Fabric.rank = rank
        mpi4py._Fabric.size = size
        mpi4py._Fabric.conns = conns
        out = target(rank, size, *args)
        result_conn.send(('ok', out))
    except BaseException:
        result_conn.send(('err', traceback.format_exc()))
    finally:
        result_conn.close()


def run_ranks(size, target, args=(), timeout=None):
    """Run ``target(rank, size, *args)`` on ``size`` forked ranks; returns the
    list of per-rank return values (rank order)."""
    ctx = mp.get_context('fork')
    mesh = [dict() for _ in range(size)]
    for i in range(size):
        for j in range(i + 1, size):
            a, b = ctx.Pipe(duplex=True)
            mesh[i][j] = a
            mesh[j][i] = b
    procs, rconns = [], []
    for r in range(size):
        pr, pc = ctx.Pipe(duplex=False)
        p = ctx.Process(target=_child, args=(r, size, mesh[r], pc, target, args))
        p.start()
        pc.close()
        procs.append(p)
        rconns.append(pr)
    results = []
    err = None
    for r in range(size):
        if rconns[r].poll(timeout):
            tag, val = rconns[r].recv()
        else:
            tag, val = 'err', 'rank %d timed out' % r
        if tag == 'err' and err is None:
            err = val
        results.append(val)
    for p in procs:
        p.join(timeout=5)
        if p.is_alive():
            p.terminate()
    if err is not None:
        raise RuntimeError('reference rank failed:\n' + err)
    return results
