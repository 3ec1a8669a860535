"""The NCCL communicators of libdnmf.so (include/dnmf.h: dnmf_comm_*, dnmf_allreduce, ...) against torch.distributed,
one rank per GPU:   torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/comm_check.py"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from pydnmfk_b200 import _lib as L
    from pydnmfk_b200.dist_comm import MPI, MPI_comm
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', rank)))
    comm = MPI.COMM_WORLD
    assert comm.backend == 'nccl'
    ver = C.c_int(0)
    L.call('dnmf_comm_nccl_version', C.byref(ver))
    bad = 0
    for dt in (torch.float32, torch.float64, torch.int64):
        x = (torch.arange(1000, device='cuda') % 13 + rank).to(dt)
        ref = x.clone()
        dist.all_reduce(ref)
        os.environ['DNMF_TORCH_COLLECTIVES'] = '0'
        got = comm.allreduce_(x.clone())
        bad += int(not torch.equal(got, ref))
    t = torch.full((3, 5), float(rank), device='cuda')
    g = comm.allgather_cat(t)
    bad += int(not torch.equal(g, torch.cat([torch.full((3, 5), float(q), device='cuda') for q in range(world)])))
    full = torch.arange(float(world * 4 * 3), device='cuda').reshape(world * 4, 3) * (rank + 1)
    rs = comm.reduce_scatter_rows(full)
    want = torch.arange(float(world * 4 * 3), device='cuda').reshape(world * 4, 3) * sum(range(1, world + 1))
    bad += int(not torch.equal(rs, want[rank * 4:(rank + 1) * 4]))
    b = torch.full((7,), float(rank + 5), device='cuda')
    comm.bcast_(b, root=world - 1)
    bad += int(not torch.equal(b, torch.full((7,), float(world + 4), device='cuda')))
    # sub-communicators of a 2-D grid (when the world factors) and ncclCommSplit through the C-ABI
    if world % 2 == 0:
        comms = MPI_comm(comm, world // 2, 2)
        row, col = comms.cart_1d_row(), comms.cart_1d_column()
        v = torch.tensor([float(rank)], device='cuda')
        i, j = divmod(rank, 2)
        bad += int(row.allreduce_(v.clone()).item() != sum(ii * 2 + j for ii in range(world // 2)))
        bad += int(col.allreduce_(v.clone()).item() != sum(i * 2 + jj for jj in range(2)))
        sub = C.c_void_p()
        L.call('dnmf_comm_split', comm._library_comm(), rank % 2, rank, C.byref(sub))
        r, s = C.c_int(), C.c_int()
        L.call('dnmf_comm_rank', sub, C.byref(r), C.byref(s))
        bad += int((r.value, s.value) != (rank // 2, world // 2))
        w = torch.tensor([1.0], device='cuda')
        L.call('dnmf_allreduce', sub, w.data_ptr(), 1, L.F32, torch.cuda.current_stream().cuda_stream)
        bad += int(w.item() != world // 2)
    flag = torch.tensor([bad], device='cuda')
    dist.all_reduce(flag)
    torch.cuda.synchronize()
    if rank == 0:
        print('NCCL version %d; %s' % (ver.value, 'comm check ok' if flag.item() == 0 else 'comm check FAILED (%d)' % flag.item()),
              flush=True)
    dist.barrier()
    sys.stdout.flush()
    os._exit(0 if flag.item() == 0 else 1)


if __name__ == '__main__':
    main()
