"""Multi-GPU leg of bench.py: hash-sharded Finder graph build, one process per GPU (launched by torchrun)."""
import json
import os
import time

import numpy as np


def main(args, cfg, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import bench
    from nrpcalc_b200 import _native, sharded

    # NCCL may print its version banner to stdout; rank 0 must print exactly one JSON line there
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    k = cfg['Lmax'] + 1
    # default: BASELINE config 4, the named 10M-part batch sharded across the GPUs (strong scaling);
    # --config c2w: one BASELINE config 2 batch per GPU (weak scaling, N x 1M parts)
    weak = bool(getattr(args, 'weak', False))
    n_parts = cfg['n'] * world if weak else cfg['n']
    buf, off = bench.make_workload(cfg, n_parts)
    ids = np.arange(n_parts, dtype=np.uint32)[::-1].copy()
    want_edges = not args.no_edges
    ctx = _native.Context(local_rank)
    ops = sharded.NativeOps(ctx)
    host_ascii = torch.from_numpy(buf.copy()).pin_memory()
    dev_ascii = host_ascii.cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    ops.load(None, off, ids, device_ptr=dev_ascii.data_ptr())

    def step(phase_ms=None, gather_host=False):
        return sharded.build_sharded(ops, off, k, cfg['internal_repeats'], want_edges=want_edges, phase_ms=phase_ms,
                                     gather_host=gather_host)

    sampler = bench.ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                                    # before the warm-up: nvidia-smi needs a moment before its first row
    for _ in range(args.warmup):
        res = step()
    sampler.mark()
    lib = _native.load_library()
    launches0 = int(lib.nrp_launches_total())
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phase_ms = {}
    total_ms = 0.0
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        start.record()
        res = step(phase_ms)
        stop.record()
        torch.cuda.synchronize()
        dt = torch.tensor([(time.perf_counter() - t0) * 1e3], device='cuda')
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)          # slowest rank
        total_ms += float(dt.item())
    launches = torch.tensor([int(lib.nrp_launches_total()) - launches0], device='cuda', dtype=torch.int64)
    dist.all_reduce(launches)                              # kernels of this library on all ranks during the timed steps
    # a timed region of a few milliseconds can end before nvidia-smi has printed a row: extra UNTIMED steps of the same work
    # (collective, so rank 0 decides for everyone)
    t_top = time.perf_counter()
    while True:
        more = torch.tensor([1 if (rank == 0 and sampler.need_more() and time.perf_counter() - t_top < 4.0) else 0], device='cuda')
        dist.broadcast(more, 0)
        if int(more.item()) == 0:
            break
        step()
        sampler.extra_steps += 1
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps

    # end to end: host buffers in (every rank stages all parts), merged host results out
    e2e = []
    for _ in range(max(2, args.steps // 2)):
        flush.fill_(1)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        ops.load(host_ascii.numpy(), off, ids)
        res = step(gather_host=True)
        dt = torch.tensor([(time.perf_counter() - t0) * 1e3], device='cuda')
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e.append(float(dt.item()))
    m = res['stats']['records']
    stage = ctx.timings()
    all_phases = [None] * world
    dist.all_gather_object(all_phases, {name: v / args.steps for name, v in phase_ms.items()})
    if rank == 0:
        peak, peak_src = bench.measured_peak_gbs()
        key_bytes = 8 if k > 16 else 4
        packed = bool(ctx.counts().get('packed', 0))
        rec = 8 if packed else key_bytes + 4           # bytes of a record on the wire and in HBM
        sent = res['stats']['sent']
        a2a_ms = max(sum(p.get(name, 0.0) for name in ('a2a_counts', 'a2a_alloc', 'a2a_keys', 'a2a_vals', 'all_to_all', 'scatter', 'barrier')) for p in all_phases)
        bytes_out = max(sent) * rec * (world - 1) / world
        slowest = max(range(world), key=lambda r: all_phases[r]['group'])
        line = {'metric': 'finder_graph_build_kmers_per_s', 'value': m / (ms_per_step * 1e-3), 'unit': 'k-mers/s', 'n_gpus': world,
                'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak' if weak else 'strong',
                'vs_baseline': None, 'dtype': 'u64' if key_bytes == 8 else 'u32', 'data': 'synthetic',
                'config': {'workload': cfg['label'] + (' -- x{} parts (one config-2 batch per GPU)'.format(world) if weak else ''), 'n_parts': n_parts, 'part_len': cfg['length'], 'Lmax': cfg['Lmax'], 'k': k,
                           'kmers': m, 'edges': want_edges, 'record_bytes': rec, 'sharding': 'parts block-sharded, k-mers hash-partitioned, 1 exchange (' + os.environ.get('NRPCALC_B200_EXCHANGE', 'peer') + ')',
                           'l2': 'explicit 256 MiB flush between steps', 'timing': 'wall clock between barrier+synchronize around the device phase (results resident on the devices), max over ranks'},
                'parts_per_s': n_parts / (ms_per_step * 1e-3),
                'e2e': {'value': m / (np.mean(e2e) * 1e-3), 'unit': 'k-mers/s', 'h2d_bytes_per_step': int(world * (buf.nbytes + off.nbytes + ids.nbytes)),
                        'd2h_bytes_per_step': int(world * (res['flags'].nbytes + res['members'].nbytes + res['indptr'].nbytes + res['last_single'].nbytes)),
                        'ms_per_step': float(np.mean(e2e)), 'path': 'per rank: nrp_load_parts(host, all parts) + sharded build + merged host results'},
                'gpu_launches': int(launches.item()),
                'clocks': clocks,
                'phase_ms_per_rank': all_phases,
                'roofline': {'bound': 'hbm', 'kernel': 'per-rank pipeline (group phase of the slowest rank)', 'unit': 'GB/s', 'peak': peak,
                             'peak_source': peak_src,
                             'achieved': res['stats']['received'][slowest] * rec * 4 / (all_phases[slowest]['group'] * 1e-3) / 1e9,
                             'frac': res['stats']['received'][slowest] * rec * 4 / (all_phases[slowest]['group'] * 1e-3) / 1e9 / peak,
                             'traffic': None,
                             'nvlink': {'bytes_sent_per_gpu': int(bytes_out), 'all_to_all_ms': a2a_ms,
                                        'GB/s_per_direction': bytes_out / (a2a_ms * 1e-3) / 1e9, 'peak_GB/s': 770.0}},
                'stage_ms_rank0_last_job': stage,
                'result': {'records': m, 'groups': int(res['indptr'].size - 1), 'edges': int(res['edges'][0].size) if res['edges'] else -1,
                           'excluded': int((res['flags'] != 0).sum()), 'sent': sent, 'received': res['stats']['received']}}
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    ctx.close()
    dist.destroy_process_group()
