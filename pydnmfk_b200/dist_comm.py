"""Communicators of the p_r x p_c processor grid, over ``torch.distributed``.

Mirrors ``pyDNMFk/dist_comm.py:2-56`` (``MPI_comm``) and the slice of the
``mpi4py`` communicator API the update loop uses, so
code written against the reference reads the same here::

    comm  = MPI.COMM_WORLD
    comms = MPI_comm(comm, p_r, p_c)
    args.comm1, args.row_comm, args.col_comm = comms.comm, comms.cart_1d_row(), comms.cart_1d_column()

One process per GPU.  With the ``nccl`` backend device tensors are reduced in
place over NVLink/NVSwitch; with ``gloo`` (CPU-side tests, or several ranks
sharing one GPU) device tensors are staged through host memory.  A single
process without an initialised process group gets size-1 communicators whose
collectives are no-ops.

Naming trap kept from the reference (dist_comm.py:34,48): ``cart_1d_row()``
keeps grid dimension 0, i.e. it spans the p_r ranks of one grid *column*;
``cart_1d_column()`` spans the p_c ranks of one grid *row*.
"""
import os
import time

import ctypes as _C
import socket

import numpy as np
import torch
import torch.distributed as dist

from . import _lib as L

_NCCL_DT = {torch.float32: L.F32, torch.float64: L.F64, torch.int64: L.I64}
_comm_cache = {}       # (member tuple, dims) -> Comm (keeps its library communicator across fits)
_group_cache = {}      # member tuple -> torch process group (new_group is world-collective: create each one once)


def _use_library_nccl():
    """Device collectives go through the NCCL communicators owned by libdnmf.so (include/dnmf.h, dnmf_comm_*)
    unless DNMF_TORCH_COLLECTIVES=1 keeps them on torch.distributed's own NCCL process group."""
    return os.environ.get('DNMF_TORCH_COLLECTIVES', '0') != '1'


def _ensure_world():
    """Initialise torch.distributed from the launcher's environment if needed."""
    if dist.is_available() and dist.is_initialized():
        return True
    ws = int(os.environ.get('WORLD_SIZE', '1'))
    if ws <= 1:
        return False
    backend = os.environ.get('DNMF_BACKEND')
    if backend is None:
        backend = 'nccl' if torch.cuda.is_available() and torch.cuda.device_count() >= int(
            os.environ.get('LOCAL_WORLD_SIZE', ws)) else 'gloo'
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', '29511')
    if torch.cuda.is_available():
        ndev = torch.cuda.device_count()
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', os.environ.get('RANK', '0'))) % ndev)
    kw = {}
    if os.environ.get('DNMF_PG_TIMEOUT_S'):
        import datetime
        kw['timeout'] = datetime.timedelta(seconds=int(os.environ['DNMF_PG_TIMEOUT_S']))
    dist.init_process_group(backend=backend, rank=int(os.environ['RANK']), world_size=ws, **kw)
    return True


class _Sum:
    name = 'SUM'


class Comm:
    """A communicator: an ordered list of world ranks plus (if size > 1) a process group."""

    def __init__(self, ranks, group=None, dims=None):
        self._ranks = list(ranks)
        self._group = group
        self._dims = list(dims) if dims is not None else None
        me = dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0
        self._me = self._ranks.index(me)
        self._subs = {}
        self._lib_comm = None          # NCCL communicator owned by libdnmf.so, created on first use

    # ---- NCCL communicator inside libdnmf.so (the C-ABI's dnmf_comm_*) ---------------------------
    def _library_comm(self):
        """Collective over this communicator on first use: rank 0 creates the 128-byte NCCL id, torch.distributed
        (plumbing) hands it to the members, every member joins with dnmf_comm_init_rank on its current device."""
        parent = getattr(self, '_parent', None)
        if parent is not None:                 # a Cartesian view of the same members
            return parent._library_comm()
        if self._lib_comm is None:
            ident = _C.create_string_buffer(128)
            if self._me == 0:
                L.call('dnmf_comm_unique_id', ident)
            box = [ident.raw]
            dist.broadcast_object_list(box, src=self._ranks[0], group=self._group, device=self._dev())
            out = _C.c_void_p()
            L.call('dnmf_comm_init_rank', _C.create_string_buffer(box[0], 128), self.size, self._me, _C.byref(out))
            self._lib_comm = out
        return self._lib_comm

    def _lib_ok(self, t):
        return (self.backend == 'nccl' and t.is_cuda and t.dtype in _NCCL_DT and t.is_contiguous()
                and _use_library_nccl())

    @staticmethod
    def _stream():
        return torch.cuda.current_stream().cuda_stream

    # ---- identity (mpi4py spelling) -------------------------------------------------------
    @property
    def rank(self):
        return self._me

    @property
    def size(self):
        return len(self._ranks)

    def Get_rank(self):
        return self._me

    def Get_size(self):
        return len(self._ranks)

    @property
    def ranks(self):
        return list(self._ranks)

    @property
    def backend(self):
        if self.size == 1 or not dist.is_initialized():
            return None
        return dist.get_backend(self._group)

    # ---- tensor collectives (device-resident hot path) -------------------------------------
    def _staged(self, t):
        return self.backend == 'gloo' and t.is_cuda

    def allreduce_(self, t):
        """In-place SUM all-reduce of a torch tensor; returns it."""
        if self.size == 1:
            return t
        if self._staged(t):
            h = t.detach().cpu()
            dist.all_reduce(h, group=self._group)
            t.copy_(h)
        elif self._lib_ok(t):
            L.call('dnmf_allreduce', self._library_comm(), t.data_ptr(), t.numel(), _NCCL_DT[t.dtype], self._stream())
        else:
            dist.all_reduce(t, group=self._group)
        return t

    def allgather_cat(self, t, sizes=None):
        """Concatenate every member's tensor along dim 0 (communicator-rank order).
        ``sizes`` = dim-0 length on every member when they differ (ragged shards)."""
        if self.size == 1:
            return t
        t = t.contiguous()
        if sizes is None or len(set(sizes)) == 1:
            out = torch.empty((self.size * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            if self._staged(t):
                ho = out.cpu()
                dist.all_gather_into_tensor(ho, t.cpu(), group=self._group)
                out.copy_(ho)
            elif self._lib_ok(t):
                L.call('dnmf_allgather', self._library_comm(), t.data_ptr(), out.data_ptr(), t.numel(), _NCCL_DT[t.dtype],
                       self._stream())
            else:
                dist.all_gather_into_tensor(out, t, group=self._group)
            return out
        mx = max(sizes)
        pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[:t.shape[0]] = t
        full = self.allgather_cat(pad)
        parts = [full[q * mx:q * mx + sizes[q]] for q in range(self.size)]
        return torch.cat(parts, dim=0)

    def reduce_scatter_rows(self, t, sizes=None):
        """SUM-reduce ``t`` (same shape on every member) and return this member's block of rows
        (contiguous chunk of the flattened buffer, as MPI Reduce_scatter, dist_nmf.py:169,202)."""
        if self.size == 1:
            return t
        t = t.contiguous()
        if sizes is None:
            assert t.shape[0] % self.size == 0
            sizes = [t.shape[0] // self.size] * self.size
        if len(set(sizes)) == 1 and not self._staged(t) and self.backend == 'nccl':
            out = torch.empty((sizes[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            if self._lib_ok(t):
                L.call('dnmf_reduce_scatter', self._library_comm(), t.data_ptr(), out.data_ptr(), out.numel(),
                       _NCCL_DT[t.dtype], self._stream())
            else:
                dist.reduce_scatter_tensor(out, t, group=self._group)
            return out
        # ragged shards or gloo: all-reduce, then slice
        full = self.allreduce_(t.clone())
        off = sum(sizes[:self._me])
        return full[off:off + sizes[self._me]].clone()

    def bcast_(self, t, root=0):
        if self.size == 1:
            return t
        src = self._ranks[root]
        if self._staged(t):
            h = t.detach().cpu()
            dist.broadcast(h, src=src, group=self._group)
            t.copy_(h)
        elif self._lib_ok(t):
            L.call('dnmf_bcast', self._library_comm(), t.data_ptr(), t.numel(), _NCCL_DT[t.dtype], int(root), self._stream())
        else:
            dist.broadcast(t, src=src, group=self._group)
        return t

    # ---- mpi4py-style object collectives (numpy arrays, scalars, tensors) -----------------------
    def _dev(self):
        if self.backend == 'nccl':
            return torch.device('cuda', torch.cuda.current_device())
        return torch.device('cpu')

    def allreduce(self, obj, op=None):
        if isinstance(obj, torch.Tensor):
            return self.allreduce_(obj.clone())
        if self.size == 1:
            return obj
        arr = np.asarray(obj)
        t = torch.from_numpy(np.ascontiguousarray(arr).reshape(-1).copy()).to(self._dev())
        dist.all_reduce(t, group=self._group)
        res = t.cpu().numpy().reshape(arr.shape)
        if isinstance(obj, np.ndarray):
            return res.astype(arr.dtype, copy=False)
        if isinstance(obj, (int, np.integer)):
            return int(res)
        if isinstance(obj, np.floating):
            return type(obj)(res)
        return float(res)

    def allgather(self, obj):
        if self.size == 1:
            return [obj]
        out = [None] * self.size
        dist.all_gather_object(out, obj, group=self._group)
        return out

    def bcast(self, obj, root=0):
        if self.size == 1:
            return obj
        if isinstance(obj, torch.Tensor):
            return self.bcast_(obj, root)
        box = [obj]
        dist.broadcast_object_list(box, src=self._ranks[root], group=self._group, device=self._dev())
        return box[0]

    def barrier(self):
        if self.size > 1:
            dist.barrier(group=self._group)

    Barrier = barrier

    def Bcast(self, buf, root=0):
        if self.size == 1:
            return
        if isinstance(buf, torch.Tensor):
            self.bcast_(buf, root)
            return
        t = torch.from_numpy(buf).to(self._dev())
        dist.broadcast(t, src=self._ranks[root], group=self._group)
        buf[...] = t.cpu().numpy()

    def Reduce_scatter(self, sendbuf, recvbuf, op=None):
        """Flat SUM reduce-scatter; this rank receives ``recvbuf.size`` elements."""
        if isinstance(sendbuf, torch.Tensor):
            flat = sendbuf.reshape(-1)
            n_here = recvbuf.numel()
        else:
            flat = torch.from_numpy(np.ascontiguousarray(sendbuf).reshape(-1)).to(self._dev())
            n_here = recvbuf.size
        if self.size == 1:
            res = flat
        else:
            sizes = [int(s) for s in self.allgather(int(n_here))]
            res = self.reduce_scatter_rows(flat, sizes)
        if isinstance(recvbuf, torch.Tensor):
            recvbuf.reshape(-1).copy_(res)
        else:
            recvbuf.reshape(-1)[...] = res.cpu().numpy()

    # ---- topology --------------------------------------------------------------------------
    def Create_cart(self, dims, periods=None, reorder=False):
        assert int(np.prod(dims)) == self.size, 'grid %s does not match %d ranks' % (dims, self.size)
        c = Comm(self._ranks, self._group, dims=dims)
        c._parent = self                      # same members, same order: share the library communicator
        return c

    def Get_coords(self, rank):
        return [int(c) for c in np.unravel_index(rank, self._dims)]

    def Sub(self, remain_dims):
        """MPI_Cart_sub: ranks that agree on every dropped coordinate, ordered by the kept ones.
        Collective over the whole grid (every rank creates every sub-group, in the same order)."""
        key = tuple(bool(r) for r in remain_dims)
        if key in self._subs:
            return self._subs[key]
        keep = [d for d, r in enumerate(remain_dims) if r]
        drop = [d for d in range(len(self._dims)) if d not in keep]
        buckets = {}
        for r in range(self.size):
            c = self.Get_coords(r)
            buckets.setdefault(tuple(c[d] for d in drop), []).append(r)
        mine = tuple(self.Get_coords(self._me)[d] for d in drop)
        sub = None
        for bkey in sorted(buckets):
            members = [self._ranks[r] for r in buckets[bkey]]
            group = None
            if len(members) > 1 and dist.is_available() and dist.is_initialized():
                # torch.distributed.new_group is collective over the WHOLE default group: Create_cart / Sub must be
                # reached by every process in the same order (they are: MPI_comm is built by every rank).  Groups are
                # cached by member list, so repeated fits re-use them instead of leaking one per MPI_comm.
                gkey = tuple(members)
                if gkey not in _group_cache:
                    if members == self._ranks and self._group is not None:
                        _group_cache[gkey] = self._group          # the sub-communicator spans this whole communicator
                    elif self.size == dist.get_world_size():
                        _group_cache[gkey] = dist.new_group(ranks=members)
                    else:
                        # a grid inside a replica group (NMFk ensemble spread over groups of ranks): only the members
                        # of this communicator get here, so the creation must not wait for the rest of the world
                        _group_cache[gkey] = dist.new_group(ranks=members, use_local_synchronization=True)
                group = _group_cache[gkey]
            if bkey == mine:
                ckey = (tuple(members), tuple(self._dims[d] for d in keep))
                if ckey not in _comm_cache:
                    _comm_cache[ckey] = Comm(members, group, dims=[self._dims[d] for d in keep])
                sub = _comm_cache[ckey]
        self._subs[key] = sub
        return sub

    def Free(self):
        pass


def split_into_groups(world, group_size):
    """Consecutive blocks of ``group_size`` world ranks as communicators (collective over ``world``: every rank creates
    every group, in the same order); returns (this rank's group communicator, group index, number of groups)."""
    assert world.size % group_size == 0
    n_groups = world.size // group_size
    mine = None
    for g in range(n_groups):
        members = [world.ranks[g * group_size + j] for j in range(group_size)]
        gkey = tuple(members)
        if group_size > 1 and dist.is_available() and dist.is_initialized() and gkey not in _group_cache:
            _group_cache[gkey] = dist.new_group(ranks=members)
        if g == world.rank // group_size:
            ckey = (gkey, None)
            if ckey not in _comm_cache:
                _comm_cache[ckey] = Comm(members, _group_cache.get(gkey))
            mine = _comm_cache[ckey]
    return mine, world.rank // group_size, n_groups


class _MPI:
    """Tiny stand-in namespace so reference-style call sites (``MPI.COMM_WORLD``, ``MPI.SUM``) work."""
    SUM = _Sum()

    def __init__(self):
        self._world = None

    @property
    def COMM_WORLD(self):
        if self._world is None:
            if _ensure_world():
                self._world = Comm(range(dist.get_world_size()), None)
            else:
                self._world = Comm([0], None)
        return self._world

    def _reset(self):
        self._world = None
        _comm_cache.clear()
        _group_cache.clear()

    @staticmethod
    def Wtime():
        return time.time()


MPI = _MPI()


class MPI_comm():
    """Cartesian p_r x p_c topology + row/column sub-communicators (dist_comm.py:16-56)."""

    def __init__(self, comm, p_r, p_c):
        self.comm = comm
        self.rank = self.comm.Get_rank()
        self.size = self.comm.Get_size()
        self.p_r = p_r
        self.p_c = p_c
        self.cartesian2d = self.comm.Create_cart(dims=[self.p_r, self.p_c], periods=[False, False], reorder=False)
        self.coord2d = self.cartesian2d.Get_coords(self.rank)

    def cart_1d_row(self):
        """Sub-communicator that keeps grid dimension 0 (p_r members)."""
        self.cartesian1d_row = self.cartesian2d.Sub(remain_dims=[True, False])
        self.rank1d_row = self.cartesian1d_row.Get_rank()
        self.coord1d_row = self.cartesian1d_row.Get_coords(self.rank1d_row)
        return self.cartesian1d_row

    def cart_1d_column(self):
        """Sub-communicator that keeps grid dimension 1 (p_c members)."""
        self.cartesian1d_column = self.cartesian2d.Sub(remain_dims=[False, True])
        self.rank1d_column = self.cartesian1d_column.Get_rank()
        self.coord1d_column = self.cartesian1d_column.Get_coords(self.rank1d_column)
        return self.cartesian1d_column

    def Free(self):
        """The reference re-derives and frees the two sub-communicators (dist_comm.py:53-56);
        process groups are cached here, so this only keeps the call legal."""
        self.cart_1d_row().Free()
        self.cart_1d_column().Free()
