"""Small builds through the warp-private duplicate filter (both versions, both map sizes, both bit positions of the k-mer) for
compute-sanitizer:  compute-sanitizer --tool memcheck|racecheck python tools/sanitize_fw.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import cases                                   # noqa: E402
from nrpcalc_b200 import _native               # noqa: E402
from oracle import np_oracle, ref_port         # noqa: E402

ctx = _native.Context(0)
for version in (1,):
    ctx.set_option('filter_warp', version)
    for n, length, k, target in ((6000, 100, 17, 0), (6000, 100, 21, 0), (5000, 100, 17, 200), (3000, 60, 23, 150)):
        ctx.set_option('warp_leaf_target', target)
        parts = cases.random_parts(n, length, n + k) + cases.planted_parts(300, 7)
        uniq = [(i, s) for i, s in ref_port.uniquify(parts) if len(s) > k]
        seqs = [s for _, s in uniq]
        ids = np.array([pid for pid, _ in uniq], dtype=np.uint32)
        off = np.zeros(len(seqs) + 1, dtype=np.int64)
        np.cumsum([len(s) for s in seqs], out=off[1:])
        ctx.load_parts(np.frombuffer(''.join(seqs).encode(), dtype=np.uint8), off, ids)
        ctx.build(k, True, True)
        res = ctx.fetch()
        want = np_oracle.spec(uniq, k, True)
        assert res.counts['records'] == want['n_records'] and res.counts['groups'] == len(want['groups']), (version, n, k)
        print('ok', version, n, length, k, target, 'filter_warp', res.counts['filter_warp'], 'buckets', res.counts['buckets'])
