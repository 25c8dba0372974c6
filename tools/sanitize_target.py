"""Small builds that touch every kernel added in round 2 (folded flags, general path, device uniquify, overlapped tail), for
compute-sanitizer:  compute-sanitizer --tool memcheck|racecheck python tools/sanitize_target.py"""
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


def load(uniq):
    seqs = [s for _, s in uniq]
    ids = np.array([pid for pid, _ in uniq], dtype=np.uint32)
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum([len(s) for s in seqs], out=off[1:])
    ctx.load_parts(np.frombuffer(''.join(seqs).encode(), dtype=np.uint8), off, ids)


# folded flags: uniform parts, even k, K == k background
parts = cases.random_parts(1500, 80, 1)
for i in range(0, 1500, 3):
    parts[i] = parts[i][:20] + parts[(i * 7) % 1500][10:40] + parts[i][50:]
bg = np_oracle.canonical_key_set([p[5:45] for p in parts[::25]], 16)
uniq = ref_port.uniquify(list(parts))
ctx.set_background(bg, 16)
load(uniq)
ctx.build(16, False, True)
res = ctx.fetch()
want = np_oracle.spec(uniq, 16, False, bg, 16)
assert res.counts['uniform_kernel'] == 2 and ((res.flags != 0) == (want['flags'] != 0)).all() and res.counts['records'] == want['n_records']
ctx.set_background(None, 0)
# 64-bit keys, no fold (k odd)
ctx.build(17, False, True)
assert ctx.fetch().counts['records'] == np_oracle.spec(uniq, 17, False)['n_records']
# general path on mixed alphabets, k = 13 and k = 40
mixed = [(pid, (s.replace('T', 'U') if pid % 3 == 0 else s[:7] + 'N' + s[8:] if pid % 5 == 0 else s)) for pid, s in uniq]
load(mixed)
for k in (13, 40):
    assert ctx.build_generic(k, False, True) == 0
    ctx.fetch()
assert ctx.build_generic(13, True, True) == 0
# device uniquify: uniform and variable lengths, duplicates
blob = np.frombuffer(''.join(parts + parts[:300]).encode(), dtype=np.uint8)
ids, source, new_off, dev, status, longest = ctx.uniquify(blob, None, 80)
assert status == 0 and ids.size == len(set(parts))
var = cases.buffer_to_list(*cases.random_varlen_buffer(1200, 0, 90, 3))
voff = np.zeros(len(var) + 1, dtype=np.int64)
np.cumsum([len(s) for s in var], out=voff[1:])
ids, source, new_off, dev, status, longest = ctx.uniquify(np.frombuffer(''.join(var).encode(), dtype=np.uint8), voff, 0)
assert status == 0 and ids.size == len(set(var))
ctx.load_parts(None, new_off, ids, device_ptr=dev)
print('sanitize target ok')
