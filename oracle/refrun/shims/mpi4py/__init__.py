"""Fork-based stand-in for the small ``mpi4py`` surface the reference uses.

TEST INFRASTRUCTURE ONLY (oracle side).  It exists so that the *unmodified*
reference (``/root/reference``) can be executed in the authoring container,
which has no MPI, to generate the golden vectors under ``tests/golden/``.
Nothing in the product path imports it.

Semantics follow SURVEY.md Appendix B: every rank is a forked OS process (so
each has its own global ``np.random`` stream), lowercase collectives move
pickled objects, ``Reduce_scatter`` sums flat buffers and hands each rank a
contiguous chunk whose size is that rank's ``recvbuf.size``.  Reductions are
evaluated on the communicator's first member in rank order.
"""
import time as _time

import numpy as _np


class _Op:
    def __init__(self, name):
        self.name = name


class _Fabric:
    """Full mesh of duplex pipes between forked ranks (filled by launcher)."""
    rank = 0
    size = 1
    conns = {}          # peer world rank -> Connection


def _send(dst, obj):
    _Fabric.conns[dst].send(obj)


def _recv(src):
    return _Fabric.conns[src].recv()


def _combine(vals, op):
    acc = vals[0]
    for v in vals[1:]:
        acc = acc + v
    return acc


class Comm:
    def __init__(self, members, dims=None):
        self._members = list(members)
        self._dims = list(dims) if dims is not None else None
        self._me = self._members.index(_Fabric.rank)

    # -- identity -----------------------------------------------------------
    @property
    def rank(self):
        return self._me

    @property
    def size(self):
        return len(self._members)

    def Get_rank(self):
        return self._me

    def Get_size(self):
        return len(self._members)

    # -- helpers ------------------------------------------------------------
    def _gather_to_root(self, obj):
        root = self._members[0]
        if _Fabric.rank == root:
            vals = [obj]
            for w in self._members[1:]:
                vals.append(_recv(w))
            return vals
        _send(root, obj)
        return None

    def _from_root(self, make):
        """root computes make(i)->payload for each member index i."""
        root = self._members[0]
        if _Fabric.rank == root:
            for i, w in enumerate(self._members[1:], start=1):
                _send(w, make(i))
            return make(0)
        return _recv(root)

    # -- object collectives -------------------------------------------------
    def allreduce(self, obj, op=None):
        vals = self._gather_to_root(obj)
        res = _combine(vals, op) if vals is not None else None
        return self._from_root(lambda i: res)

    def allgather(self, obj):
        vals = self._gather_to_root(obj)
        return self._from_root(lambda i: vals)

    def bcast(self, obj, root=0):
        wroot = self._members[root]
        if _Fabric.rank == wroot:
            for w in self._members:
                if w != wroot:
                    _send(w, obj)
            return obj
        return _recv(wroot)

    def scatter(self, objs, root=0):
        wroot = self._members[root]
        if _Fabric.rank == wroot:
            for i, w in enumerate(self._members):
                if w != wroot:
                    _send(w, objs[i])
            return objs[root]
        return _recv(wroot)

    def barrier(self):
        self.allreduce(0)

    Barrier = barrier

    # -- buffer collectives -------------------------------------------------
    def Bcast(self, buf, root=0):
        got = self.bcast(_np.asarray(buf) if self._me == root else None, root=root)
        if self._me != root:
            buf[...] = got.reshape(buf.shape)

    def Reduce_scatter(self, sendbuf, recvbuf, op=None):
        vals = self._gather_to_root((_np.ascontiguousarray(sendbuf).ravel(), recvbuf.size))
        if vals is not None:
            total = _combine([v[0] for v in vals], op)
            sizes = [v[1] for v in vals]
            offs = _np.concatenate(([0], _np.cumsum(sizes)))
            chunk = self._from_root(lambda i: total[offs[i]:offs[i + 1]])
        else:
            chunk = self._from_root(None)
        recvbuf.ravel()[...] = chunk

    # -- topology -----------------------------------------------------------
    def Create_cart(self, dims, periods=None, reorder=False):
        assert int(_np.prod(dims)) == self.size
        return Comm(self._members, dims=dims)

    def Get_coords(self, rank):
        return [int(c) for c in _np.unravel_index(rank, self._dims)]

    def Sub(self, remain_dims):
        mine = self.Get_coords(self._me)
        keep = [d for d, r in enumerate(remain_dims) if r]
        members = []
        for r in range(self.size):
            c = self.Get_coords(r)
            if all(c[d] == mine[d] for d in range(len(self._dims)) if d not in keep):
                members.append(self._members[r])
        return Comm(members, dims=[self._dims[d] for d in keep])

    def Free(self):
        pass


class _MPI:
    SUM = _Op('sum')
    Comm = Comm

    @staticmethod
    def Wtime():
        return _time.time()

    @property
    def COMM_WORLD(self):
        return Comm(range(_Fabric.size))


MPI = _MPI()
