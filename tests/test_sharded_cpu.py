"""world_size-2/3 gloo tests of the sharded host logic (shard bounds, record all-to-all, flag OR-reduce, merge)
with the device calls replaced by the numpy model in tests/fake_sharded_ops.py.  The merged result must equal the
single-process oracle specification."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
from nrpcalc_b200 import sharded
from oracle import np_oracle, ref_port


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, parts, k, internal, bg_seqs, bg_K, out_dir):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [os.path.dirname(here), here]
    import fake_sharded_ops
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        uniq = [(pid, seq) for pid, seq in ref_port.uniquify(list(parts)) if len(seq) >= k]
        bg_keys = np_oracle.canonical_key_set(bg_seqs, bg_K) if bg_seqs else None
        ops = fake_sharded_ops.OracleOps(bg_keys, bg_K)
        ops.load([s for _, s in uniq], [pid for pid, _ in uniq])
        res = sharded.build_sharded(ops, ops.offsets, k, internal, want_edges=True)
        np.savez(os.path.join(out_dir, 'rank%d.npz' % rank), flags=res['flags'], indptr=res['indptr'], members=res['members'],
                 first_window=res['first_window'], last_single=res['last_single'], eu=res['edges'][0], ev=res['edges'][1],
                 records=res['stats']['records'])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,k,internal,with_bg', [(2, 13, False, False), (2, 16, True, False), (3, 17, False, True), (2, 6, False, False)])
def test_sharded_merge_equals_single_process_spec(tmp_path, world, k, internal, with_bg):
    parts = cases.planted_parts(220, 40 + k) if k > 6 else cases.palindrome_parts(3, n_parts=80, k=6)
    bg_seqs = cases.background_seqs_for(parts, 5) if with_bg else None
    bg_K = 15
    mp.spawn(_worker, args=(world, _free_port(), parts, k, internal, bg_seqs, bg_K, str(tmp_path)), nprocs=world, join=True)
    uniq = [(pid, seq) for pid, seq in ref_port.uniquify(list(parts)) if len(seq) >= k]
    want = np_oracle.spec(uniq, k, internal, np_oracle.canonical_key_set(bg_seqs, bg_K) if bg_seqs else None, bg_K)
    ids = want['ids']
    results = [np.load(str(tmp_path / ('rank%d.npz' % r))) for r in range(world)]
    for res in results:
        assert ((res['flags'] != 0) == (want['flags'] != 0)).all()
        groups = sorted((int(res['members'][res['indptr'][g]]), int(res['first_window'][g]),
                         tuple(int(ids[m]) for m in res['members'][res['indptr'][g]:res['indptr'][g + 1]]))
                        for g in range(len(res['indptr']) - 1))
        assert groups == sorted(want['groups'])
        assert res['last_single'].tolist() == want['last_single'].tolist()
        assert list(zip(res['eu'].tolist(), res['ev'].tolist())) == sorted(want['edges'])
    for other in results[1:]:                      # every rank ends with the same structure
        assert all((results[0][name] == other[name]).all() for name in ('flags', 'last_single', 'eu', 'ev'))


def test_shard_bounds_balance_windows():
    rng = np.random.default_rng(0)
    lengths = rng.integers(10, 300, size=1000)
    for world in (1, 2, 4, 8):
        bounds = sharded.shard_bounds(lengths, 25, world)
        assert bounds[0][0] == 0 and bounds[-1][1] == 1000
        assert all(a[1] == b[0] for a, b in zip(bounds, bounds[1:]))
        windows = np.maximum(lengths - 24, 0)
        loads = [int(windows[lo:hi].sum()) for lo, hi in bounds]
        assert max(loads) - min(loads) <= 2 * 300
    assert sharded.shard_bounds([], 13, 4) == [(0, 0)] * 4
    assert sharded.shard_bounds([5, 5], 13, 2)[-1][1] == 2
