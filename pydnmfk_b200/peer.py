"""Peer-memory exchange of the row grid's H half-step (include/dnmf.h: dnmf_symm_*, dnmf_xchg_*).

On a P x 1 grid every rank holds a replica of H and a partial ``(W_i^T A_i)^T``.  The reference all-reduces the
partial and the k x k Gram and lets every rank repeat the same update (dist_nmf.py:679-681, :705-708, :750-751).
Here each rank owns n / P columns of H: partials are written straight into the owner's memory over NVLink, the owner
sums them in rank order, updates its columns and writes them into every replica -- three kernel launches of libdnmf,
no NCCL call, 1 / P of the update arithmetic per rank, bit-identical replicas.

Host side this module only allocates the exchange regions (``dnmf_symm_alloc``), trades their CUDA IPC handles through
the communicator's object all-gather (plumbing) and keeps the table of mapped peer pointers."""
import ctypes as C
import os
import socket

import torch

from . import _lib as L

_DT = {torch.float32: L.F32, torch.float64: L.F64}
_cache = {}


def enabled_for(comm):
    """Policy: on for NCCL runs (one GPU per rank), off when DNMF_PEER_EXCHANGE=0; DNMF_PEER_EXCHANGE=1 forces it on for
    any backend (several ranks sharing one GPU in the test-suite)."""
    flag = os.environ.get('DNMF_PEER_EXCHANGE')
    if flag == '0' or comm.size < 2 or comm.size > 16:
        return False
    return flag == '1' or comm.backend == 'nccl'


class PeerExchange:
    """Exchange regions of one communicator for a k x n replicated factor.  Construction is collective."""

    def __init__(self, comm, n, k, tdtype):
        self.comm, self.n, self.k, self.tdtype = comm, int(n), int(k), tdtype
        self.P, self.me = comm.size, comm.rank
        self.dt = _DT[tdtype]
        nbytes = L.call('dnmf_xchg_bytes', self.P, self.n, self.k, self.dt)
        if nbytes <= 0:
            raise L.DnmfError('dnmf_xchg_bytes', -1, L.raw().dnmf_last_error().decode())
        local = C.c_void_p()
        handle = C.create_string_buffer(64)
        L.call('dnmf_symm_alloc', nbytes, C.byref(local), handle)
        self._local = local
        infos = comm.allgather((socket.gethostname(), handle.raw))
        if len({h for h, _ in infos}) != 1:
            L.call('dnmf_symm_free', local)
            raise RuntimeError('peer exchange needs all ranks of the communicator on one node')
        self._opened = []
        bases = []
        for q, (_, raw) in enumerate(infos):
            if q == self.me:
                bases.append(local.value)
            else:
                out = C.c_void_p()
                L.call('dnmf_symm_open', C.create_string_buffer(raw, 64), C.byref(out))
                self._opened.append(out)
                bases.append(out.value)
        self._bases = (C.c_void_p * self.P)(*bases)
        self.nbytes = int(nbytes)
        comm.barrier()              # every rank has mapped every region before the first kernel touches one

    def update_h(self, mode, H, Yt, aux, eps, clamp=False):
        """H <- update(H, sum_q Yt_q, sum_q aux_q) on every replica; mode 0 FRO-MU, 2 FRO-HALS, 3 KL-MU."""
        k, n = H.shape
        assert (k, n) == (self.k, self.n) and H.dtype == self.tdtype and Yt.shape == (n, k)
        assert H.stride(1) == 1 and Yt.is_contiguous() and aux.is_contiguous()
        L.call('dnmf_xchg_update_h', self._bases, self.P, self.me, int(mode), H.data_ptr(), H.stride(0), Yt.data_ptr(),
               Yt.stride(0), aux.data_ptr(), n, k, float(eps), 1 if clamp else 0, self.dt,
               torch.cuda.current_stream().cuda_stream)

    def update_h_p(self, mode, H, view, aux, eps, clamp=False):
        """update_h with this rank's partial given as the split-K partial view of `DeviceOps.wta_p / kl_wtu_p`."""
        k, n = H.shape
        assert (k, n) == (self.k, self.n) and H.dtype == self.tdtype and H.stride(1) == 1 and aux.is_contiguous()
        L.call('dnmf_xchg_update_h_p', self._bases, self.P, self.me, int(mode), H.data_ptr(), H.stride(0), view,
               aux.data_ptr(), n, k, float(eps), 1 if clamp else 0, self.dt, torch.cuda.current_stream().cuda_stream)

    def check(self):
        """Raise if a wait for a peer timed out (the peer died or never launched its half of an exchange)."""
        err = C.c_int(0)
        L.call('dnmf_xchg_error', self._local, C.byref(err), torch.cuda.current_stream().cuda_stream)
        if err.value:
            raise RuntimeError('peer exchange: a rank timed out waiting for its peers; results are invalid')


def get(comm, n, k, tdtype):
    """Cached exchange of (communicator members, n, k, dtype); collective on first use.  None when the policy or the
    placement (ranks on several nodes) rules it out -- decided identically on every rank."""
    if not enabled_for(comm) or tdtype not in _DT or k > L.MAX_K:
        return None
    key = (tuple(comm.ranks), int(n), int(k), tdtype)
    if key not in _cache:
        hosts = comm.allgather(socket.gethostname())
        _cache[key] = PeerExchange(comm, n, k, tdtype) if len(set(hosts)) == 1 else None
    return _cache[key]
