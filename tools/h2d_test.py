"""Host -> device copy bandwidth from page-locked memory: one copy vs several chunks on several streams (B200 box: 55.5 GB/s
with a single copy already, which is why the shard upload of PyNMF stays one cudaMemcpyAsync)."""
import torch, time
n = 4 << 30  # 4 Gi floats? too big; use 16 GiB total bytes
nbytes = 16 << 30
host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
dev = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
def run(nchunks, nstreams):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    per = nbytes // nchunks
    for i in range(nchunks):
        with torch.cuda.stream(streams[i % nstreams]):
            dev[i*per:(i+1)*per].copy_(host[i*per:(i+1)*per], non_blocking=True)
    torch.cuda.synchronize()
    return nbytes / (time.perf_counter() - t0) / 1e9
for _ in range(2):
    for nc, ns in ((1,1),(2,2),(4,2),(4,4),(8,4),(16,2)):
        print('chunks %d streams %d: %.1f GB/s' % (nc, ns, run(nc, ns)), flush=True)
