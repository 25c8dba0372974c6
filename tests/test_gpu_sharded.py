"""Multi-GPU parity: the hash-sharded build (one process per GPU, NCCL all-to-all) must give exactly the
single-GPU result.  Spawns world_size processes itself; needs >= 2 visible GPUs (gpurun --gpus 2)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_parts, k, internal, out_dir, mode='direct', opts=None, dups=0):
    import sys
    # exchange flavour: direct (fixed regions, no count pass), counted (count pass + peer stores), nccl (all-to-all)
    os.environ['NRPCALC_B200_DIRECT'] = '1' if mode == 'direct' else '0'
    os.environ['NRPCALC_B200_EXCHANGE'] = 'nccl' if mode == 'nccl' else 'peer'
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [os.path.dirname(here), here]
    import torch
    import torch.distributed as dist
    import cases
    from nrpcalc_b200 import _native, sharded
    from oracle import ref_port
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        parts = cases.planted_parts(n_parts, 1234) + cases.random_parts(n_parts, 100, 9)
        if dups:        # one 30-mer shared by many parts: fixed regions overflow, the exchange falls back to the counted path
            common = cases.random_parts(1, 30, 77)[0]
            parts = [common + p if i % 4 == 0 and i < 4 * dups else p for i, p in enumerate(parts)]
        uniq = [(pid, seq) for pid, seq in ref_port.uniquify(list(parts)) if len(seq) >= k]
        seqs = [s for _, s in uniq]
        ids = np.array([pid for pid, _ in uniq], dtype=np.uint32)
        offsets = np.zeros(len(seqs) + 1, dtype=np.int64)
        np.cumsum([len(s) for s in seqs], out=offsets[1:])
        buf = np.frombuffer(''.join(seqs).encode(), dtype=np.uint8)
        ctx = _native.Context(rank)
        opts = dict(opts or {})
        for name, value in opts.pop('_env', {}).items():      # switches of the host layer (sharded.py)
            os.environ[name] = value
        for name, value in opts.items():
            ctx.set_option(name, value)
        opts = opts or None
        ops = sharded.NativeOps(ctx)
        ops.load(buf, offsets, ids)
        res = sharded.build_sharded(ops, offsets, k, internal, want_edges=True)
        if rank == 0:
            with open(os.path.join(out_dir, 'layout.txt'), 'w') as fh:
                fh.write(repr(ctx.counts()))
        # a second, larger batch through the same context: the receive buffers are re-exported (collective close first)
        if opts is None and mode == 'direct' and k == 17:
            more = cases.random_parts(4 * n_parts, 100, 10)
            off2 = np.arange(len(more) + 1, dtype=np.int64) * 100
            buf2 = np.frombuffer(''.join(more).encode(), dtype=np.uint8)
            ids2 = np.arange(len(more), dtype=np.uint32)[::-1].copy()
            ops.load(buf2, off2, ids2)
            big = sharded.build_sharded(ops, off2, k, internal, want_edges=True)
            if rank == 0:
                ctx.load_parts(buf2, off2, ids2)
                ctx.build(k, internal, True)
                one2 = ctx.fetch()
                assert sorted(zip(big['edges'][0].tolist(), big['edges'][1].tolist())) == sorted(zip(one2.edges[0].tolist(), one2.edges[1].tolist()))
                assert big['last_single'].tolist() == one2.last_single.tolist() and big['indptr'].size == one2.indptr.size
            ops.load(buf, offsets, ids)
            res = sharded.build_sharded(ops, offsets, k, internal, want_edges=True)
        np.savez(os.path.join(out_dir, 'rank%d.npz' % rank), flags=res['flags'], indptr=res['indptr'], members=res['members'],
                 first_window=res['first_window'], last_single=res['last_single'], eu=res['edges'][0], ev=res['edges'][1],
                 records=res['stats']['records'])
        if rank == 0:       # single-GPU result of the same input on the same device
            ctx.set_option('no_window_carry', 0)
            ctx.load_parts(buf, offsets, ids)
            ctx.build(k, internal, True)
            one = ctx.fetch()
            np.savez(os.path.join(out_dir, 'single.npz'), flags=one.flags, indptr=one.indptr, members=one.members,
                     first_window=one.first_window, last_single=one.last_single, eu=one.edges[0], ev=one.edges[1],
                     records=one.counts['records'])
        ctx.close()
    finally:
        dist.destroy_process_group()


def _groups(res):
    return sorted((int(res['first_window'][g]), tuple(res['members'][res['indptr'][g]:res['indptr'][g + 1]].tolist()))
                  for g in range(len(res['indptr']) - 1))


@pytest.mark.parametrize('k,internal,mode,opts,dups', [
    (13, False, 'direct', None, 0), (17, False, 'direct', None, 0), (16, True, 'direct', None, 0), (16, False, 'direct', None, 0),
    (32, False, 'direct', None, 0),                                             # key + value exceed 64 bits: (key, value) arrays
    (13, False, 'direct', {'leaf_target': 20}, 0),                              # senders pre-partition by 5 bits, one owner level
    (17, False, 'direct', {'leaf_target': 20, 'nvlink_digits': 2}, 0),          # scatter by owner only, two owner levels
    (17, True, 'direct', {'leaf_target': 20, 'nvlink_digits': 2, 'packed': 0}, 0),
    (21, False, 'direct', None, 0), (21, True, 'direct', None, 0),              # BASELINE config 4's k: packed word, owner = top hash bits
    (21, False, 'direct', {'no_window_carry': 1}, 0),                           # ... window index NOT carried (the 4/8-GPU layout of config 4)
    (25, False, 'direct', {'no_window_carry': 1}, 0),                           # config 5's k
    (21, False, 'direct', {'no_window_carry': 1, 'leaf_target': 20, 'nvlink_digits': 2}, 0),   # owner-only scatter, two owner levels, no window
    (16, False, 'direct', {'_env': {'NRPCALC_B200_FLAG_CAP': '2'}}, 0),         # sparse flag lists truncated -> dense all-reduce, finish repeated
    (17, False, 'direct', {'_env': {'NRPCALC_B200_DENSE_FLAGS': '1'}}, 0),      # dense flag all-reduce
    (13, True, 'direct', None, 300),                                            # region overflow -> counted exchange
    (13, False, 'counted', None, 0), (17, True, 'counted', None, 0), (17, False, 'nccl', None, 0)])
def test_sharded_equals_single_gpu(tmp_path, k, internal, mode, opts, dups):
    world = min(_n_gpus(), 4)
    if world < 2:
        pytest.skip('needs at least 2 GPUs')
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, _free_port(), 3000, k, internal, str(tmp_path), mode, opts, dups), nprocs=world, join=True)
    single = np.load(str(tmp_path / 'single.npz'))
    for rank in range(world):
        res = np.load(str(tmp_path / ('rank%d.npz' % rank)))
        assert ((res['flags'] != 0) == (single['flags'] != 0)).all()
        assert int(res['records']) == int(single['records'])
        assert _groups(res) == _groups(single)
        assert res['last_single'].tolist() == single['last_single'].tolist()
        assert sorted(zip(res['eu'].tolist(), res['ev'].tolist())) == sorted(zip(single['eu'].tolist(), single['ev'].tolist()))


def _finder_worker(rank, world, port, case_name, out_dir):
    import json
    import random
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [os.path.dirname(here), here]
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    os.environ['LOCAL_RANK'] = str(rank)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        import nrpcalc_b200
        from conftest import GOLDEN_CASES
        case = next(c for c in GOLDEN_CASES if c['name'] == case_name)
        os.chdir(out_dir)
        os.makedirs('rank%d' % rank, exist_ok=True)
        os.chdir('rank%d' % rank)
        found = {}
        for vc in ('nrp2', '2apx'):
            random.seed(7)
            res = nrpcalc_b200.finder(list(case['parts']), case['Lmax'], internal_repeats=case['internal_repeats'], vercov=vc, verbose=False)
            found[vc] = list(res.keys())
        with open(os.path.join(out_dir, 'finder_rank%d.json' % rank), 'w') as fh:
            json.dump(found, fh)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('case_name', ['planted_s0_L15_ir0', 'random_2000x100_L12', 'short_parts_ir1'])
def test_finder_under_torch_distributed_matches_reference(tmp_path, case_name):
    """nrpcalc_b200.finder() called on every rank of an initialised process group: the build is hash-sharded and every
    rank returns the reference's toolbox."""
    import json
    from conftest import GOLDEN_CASES
    world = 2
    if _n_gpus() < world:
        pytest.skip('needs at least 2 GPUs')
    import torch.multiprocessing as mp
    mp.spawn(_finder_worker, args=(world, _free_port(), case_name, str(tmp_path)), nprocs=world, join=True)
    case = next(c for c in GOLDEN_CASES if c['name'] == case_name)
    for rank in range(world):
        found = json.load(open(str(tmp_path / ('finder_rank%d.json' % rank))))
        for vc in ('nrp2', '2apx'):
            assert found[vc] == case['toolbox'][vc]['7'], (rank, vc)
