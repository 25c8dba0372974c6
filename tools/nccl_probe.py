"""Probe NCCL point-to-point / all-to-all bandwidth and transport between the GPUs of the box (run under torchrun)."""
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); local = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
if rank == 0:
    print('peer access 0->1:', torch.cuda.can_device_access_peer(0, 1) if torch.cuda.device_count() > 1 else None, flush=True)
for mb in (64, 512):
    n = mb * (1 << 20) // 8
    src = torch.ones(n, dtype=torch.int64, device='cuda')
    dst = torch.empty(n, dtype=torch.int64, device='cuda')
    for _ in range(3):
        dist.all_to_all_single(dst, src)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    iters = 10
    for _ in range(iters):
        dist.all_to_all_single(dst, src)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / iters
    crossing = mb * (world - 1) / world
    if rank == 0:
        print('all_to_all %4d MB per rank: %.3f ms, %.1f GB/s per direction crossing NVLink' % (mb, dt * 1e3, crossing / 1024 / dt), flush=True)
    # plain send/recv ring
    peer = (rank + 1) % world
    t0 = time.perf_counter()
    for _ in range(iters):
        ops = [dist.P2POp(dist.isend, src, peer), dist.P2POp(dist.irecv, dst, (rank - 1) % world)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / iters
    if rank == 0:
        print('send/recv  %4d MB: %.3f ms, %.1f GB/s' % (mb, dt * 1e3, mb / 1024 / dt), flush=True)
dist.destroy_process_group()
