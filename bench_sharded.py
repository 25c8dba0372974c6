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
    host_ascii = torch.from_numpy(buf).pin_memory()       # one page-locked copy; the pageable original is dropped
    buf = host_ascii.numpy()
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

    # end to end: host buffers in, merged host results out.  Every rank stages only ITS shard of the parts (page-locked host
    # buffer, chunked copy; parts of one length need no offsets array) plus the part ids; the results come back through NCCL
    # all-gathers of the arrays and one device-to-host copy per rank.
    uniform = not isinstance(cfg['length'], tuple)
    host_ids = torch.from_numpy(ids.copy()).pin_memory()
    e2e, e2e_load = [], []
    e2e_phase = {}
    for it in range(2 + max(2, args.steps // 2)):
        flush.fill_(1)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        if uniform:
            ops.load_shard(host_ascii.numpy(), None, host_ids.numpy(), k, part_len=cfg['length'])
        else:
            ops.load_shard(host_ascii.numpy(), off, host_ids.numpy(), k)
        t1 = time.perf_counter()
        res = step(e2e_phase if it >= 2 else None, gather_host=True)
        dt = torch.tensor([(time.perf_counter() - t0) * 1e3, (t1 - t0) * 1e3], device='cuda')
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e.append(float(dt[0].item())); e2e_load.append(float(dt[1].item()))
    e2e, e2e_load = e2e[2:], e2e_load[2:]                    # the first two passes allocate and map the peers' buffers
    shard_bytes = int(buf.nbytes // world) if uniform else int(off[res['stats']['bounds'][rank][1]] - off[res['stats']['bounds'][rank][0]])
    m = res['stats']['records']
    stage = ctx.timings()
    owner_levels = 2 if stage.get('split2', 0.0) > 0 else 1        # split levels the owner ran over the received regions
    # ---- parity (outside every timed region): the sharded result against the single-GPU build of the same batch on rank 0
    # (a cut of the batch when it does not fit one device), and -- with --verify -- against oracle/nrp_oracle.c
    parity_block = verify_sharded(args, cfg, ops, ctx, buf, off, ids, k, res, rank, world, want_edges, dev_ascii)
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
        # bytes THIS pipeline has to move per rank: the sender reads its ASCII shard and stores every record once
        # (NVLink or local); the owner reads + writes the records once per split level and reads them once more in the filter
        received = res['stats']['received']
        owner_bytes = received[slowest] * rec * (1 + 2 * owner_levels)
        job_bytes = int(off[-1]) + sum(sent) * rec + sum(received) * rec * (1 + 2 * owner_levels)
        survey_b = (int(off[-1]) / max(m, 1)) + (key_bytes + 4) * (2 * -(-2 * k // 8) + 3) + key_bytes + 2 * (key_bytes + 4)
        line = {'metric': 'finder_graph_build_kmers_per_s', 'value': m / (ms_per_step * 1e-3), 'unit': 'k-mers/s', 'n_gpus': world,
                'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak' if weak else 'strong',
                'vs_baseline': None, 'dtype': 'u64' if key_bytes == 8 else 'u32', 'data': 'synthetic',
                'config': {'workload': cfg['label'] + (' -- x{} parts (one config-2 batch per GPU)'.format(world) if weak else ''), 'n_parts': n_parts, 'part_len': cfg['length'], 'Lmax': cfg['Lmax'], 'k': k,
                           'kmers': m, 'edges': want_edges, 'record_bytes': rec, 'sharding': 'parts block-sharded, k-mers hash-partitioned, 1 exchange (' + os.environ.get('NRPCALC_B200_EXCHANGE', 'peer') + ')',
                           'l2': 'explicit 256 MiB flush between steps', 'timing': 'wall clock between barrier+synchronize around the device phase (results resident on the devices), max over ranks'},
                'parts_per_s': n_parts / (ms_per_step * 1e-3),
                'e2e': {'value': m / (np.mean(e2e) * 1e-3), 'unit': 'k-mers/s',
                        'h2d_bytes_per_step': int(buf.nbytes + world * ids.nbytes + (0 if uniform else world * off.nbytes)),
                        'd2h_bytes_per_step': int(world * (res['flags'].nbytes + res['members'].nbytes + res['first_window'].nbytes + res['keys'].nbytes +
                                                           res['indptr'].nbytes + res['last_single'].nbytes + (res['edges'][0].nbytes * 2 if res['edges'] else 0))),
                        'ms_per_step': float(np.mean(e2e)), 'ms_load': float(np.mean(e2e_load)), 'h2d_bytes_this_rank': shard_bytes + int(ids.nbytes),
                        'phase_ms_rank0': {name: v / max(1, len(e2e)) for name, v in e2e_phase.items()},
                        'path': 'per rank: nrp_load_parts_shard (page-locked host buffer, ONLY the rank\'s shard of the parts travels; ids in full) + sharded build + NCCL all-gather of the result arrays + one D2H copy'},
                'gpu_launches': int(launches.item()),
                'clocks': clocks,
                'phase_ms_per_rank': all_phases,
                'roofline': {'bound': 'hbm', 'kernel': 'owner phase of the slowest rank ({} split level(s) + filter over the received records)'.format(owner_levels),
                             'unit': 'GB/s', 'peak': peak, 'peak_source': peak_src,
                             'achieved': owner_bytes / (all_phases[slowest]['group'] * 1e-3) / 1e9,
                             'frac': owner_bytes / (all_phases[slowest]['group'] * 1e-3) / 1e9 / peak,
                             'algorithmic_bytes_per_launch': int(owner_bytes), 'traffic': None,
                             'traffic_note': 'no ncu capture of a multi-rank run exists (ncu is single-GPU here); on one GPU the same kernels move their algorithmic bytes within 4 % (profiles/traffic.json)',
                             'whole_job': {'algorithmic_bytes': int(job_bytes), 'GB/s_aggregate': job_bytes / (ms_per_step * 1e-3) / 1e9,
                                           'frac_of_aggregate_peak': job_bytes / (ms_per_step * 1e-3) / 1e9 / (world * peak),
                                           'survey_formula_bytes_per_kmer': survey_b,
                                           'survey_8d_rule': 'min(M*B, measured or algorithmic bytes of this pipeline) / T / (G * peak): this pipeline moves fewer bytes than the LSD formulation, so the fraction above is the figure'},
                             'nvlink': {'bytes_sent_per_gpu': int(bytes_out), 'all_to_all_ms': a2a_ms,
                                        'GB/s_per_direction': bytes_out / (a2a_ms * 1e-3) / 1e9, 'peak_GB/s': 900.0,
                                        'peak_source': 'NVLink 5, 18 links x 50 GB/s per direction per GPU (SURVEY.md 8d; B200_PROFILING.md)'}},
                'stage_ms_rank0_last_job': stage,
                'result': {'records': m, 'groups': int(res['indptr'].size - 1), 'edges': int(res['edges'][0].size) if res['edges'] else -1,
                           'excluded': int((res['flags'] != 0).sum()), 'sent': sent, 'received': res['stats']['received']}}
        if parity_block is not None:
            line['parity'] = parity_block
            if parity_block.get('single_gpu_ms') is not None:
                line['single_gpu_same_config_ms'] = parity_block['single_gpu_ms']       # anchor of the strong-scaling curve
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    ok = parity_block is None or (parity_block.get('vs_single_gpu', True) and parity_block.get('vs_oracle', True))
    ctx.close()
    dist.destroy_process_group()
    if not ok:
        raise SystemExit(3)


def verify_sharded(args, cfg, ops, ctx, buf, off, ids, k, res, rank, world, want_edges, dev_ascii):
    """Parity of the sharded path, never inside a timed region.  The batch (or, when it exceeds what one device indexes, its
    first --verify-parts parts) is built once more on the N ranks and once on rank 0 alone; rank 0 compares the canonical
    forms (tests/parity.py): exclusion flags, every shared k-mer group (key, first window, members in order), last unshared
    windows and the distinct edge list.  --verify adds oracle/nrp_oracle.c on the same parts."""
    import torch
    import torch.distributed as dist
    from nrpcalc_b200 import sharded
    import parity
    if getattr(args, 'no_verify', False):
        return None
    n = len(off) - 1
    windows = int(np.maximum(np.diff(off) - k + 1, 0).sum())
    cut = n
    if windows >= 3_500_000_000:                       # one device indexes records with 32 bits
        cut = min(n, int(getattr(args, 'verify_parts', 5_000_000)))
    if cut < n:
        sub_off = np.ascontiguousarray(off[:cut + 1])
        sub_buf = np.ascontiguousarray(buf[:int(sub_off[-1])])
        sub_ids = np.arange(cut, dtype=np.uint32)[::-1].copy()
        ops.load(sub_buf, sub_off, sub_ids)
        ops._shard_plan = None
        res = sharded.build_sharded(ops, sub_off, k, cfg['internal_repeats'], want_edges=True, gather_host=True)
    else:
        sub_off, sub_buf, sub_ids = off, buf, ids
        if not want_edges or res.get('edges') is None:
            ops.load(None, off, ids, device_ptr=dev_ascii.data_ptr())
            res = sharded.build_sharded(ops, off, k, cfg['internal_repeats'], want_edges=True, gather_host=True)
    torch.cuda.synchronize()
    dist.barrier()
    block = None
    if rank == 0:
        got = parity.canonical(res['flags'], res['indptr'], res['members'], res['first_window'], res['keys'], res['last_single'],
                               res['edges'], None, res['stats']['records'])
        if cut < n:
            ctx.load_parts(sub_buf, sub_off, sub_ids)
        else:
            ctx.load_parts(None, off, ids, device_ptr=dev_ascii.data_ptr())
        ctx.build(k, cfg['internal_repeats'], True)                   # warm-up (buffers)
        ctx.build(k, cfg['internal_repeats'], True)
        single_ms = ctx.timings()['total']
        one = ctx.fetch(want_keys=True, copy=True)
        ref = parity.canonical(one.flags, one.indptr, one.members, one.first_window, one.keys, one.last_single, one.edges, None,
                               one.counts['records'])
        cmp = parity.compare(got, ref)
        block = {'vs_single_gpu': cmp['all'], 'components': cmp, 'digest': parity.digest(got), 'single_gpu_digest': parity.digest(ref),
                 'checked': parity.summary(got), 'parts_checked': int(cut), 'parts_total': int(n),
                 'single_gpu_ms': single_ms if cut == n else None,
                 'single_gpu_ms_on_cut': single_ms if cut < n else None}
        if getattr(args, 'verify', False):
            from oracle import c_oracle
            import bench
            t0 = time.perf_counter()
            cores = os.cpu_count() or 1
            slices = max(1, int(np.ceil(int(np.maximum(np.diff(sub_off) - k + 1, 0).sum()) / 25e6)))
            want = c_oracle.build_table(sub_buf, sub_off, k, cfg['internal_repeats'], None, 0, n_threads=cores, want_groups=True, n_slices=slices)
            cmp_o = parity.compare(got, parity.from_c_oracle(want, sub_ids))
            block.update({'vs_oracle': cmp_o['all'], 'oracle_components': cmp_o,
                          'oracle': 'oracle/nrp_oracle.c on {} parts, {} threads, {} key-space passes, {:.1f} s'.format(cut, cores, slices, time.perf_counter() - t0)})
        if not cmp['all']:
            import sys
            sys.stderr.write('PARITY MISMATCH sharded vs single GPU: {}\n'.format(cmp))
    dist.barrier()
    flag = torch.tensor([0 if (block is None or (block['vs_single_gpu'] and block.get('vs_oracle', True))) else 1], device='cuda')
    dist.broadcast(flag, 0)
    if rank != 0 and int(flag.item()) != 0:
        block = {'vs_single_gpu': False}
    return block
