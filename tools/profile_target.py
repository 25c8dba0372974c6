"""One warm-up build + one measured build of BASELINE config 2 (1M x 100 bp, Lmax=16) for ncu."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import cases                                   # noqa: E402
from nrpcalc_b200 import _native               # noqa: E402

n = int(os.environ.get('NRP_PROFILE_PARTS', '1000000'))
k = int(os.environ.get('NRP_PROFILE_K', '17'))
length = int(os.environ.get('NRP_PROFILE_LEN', '100'))
buf, off = cases.random_dna_buffer(n, length, 2)
ids = np.arange(n, dtype=np.uint32)[::-1].copy()
ctx = _native.Context(0)
for item in os.environ.get('NRP_PROFILE_OPTS', '').split(','):       # library knobs name=value,...
    if item:
        ctx.set_option(item.split('=')[0], int(item.split('=')[1]))
bg = int(os.environ.get('NRP_PROFILE_BG', '0'))          # bp of a random background genome (K = k)
if bg:
    from nrpcalc_b200 import encoding
    gbuf, goff = encoding.join_parts([cases.random_genome(bg, 33)])
    ctx.bg_build(gbuf, goff, k, install=True)
ctx.load_parts(buf, off, ids)
for _ in range(int(os.environ.get("NRP_PROFILE_BUILDS", "2"))):
    ctx.build(k, False, True)
print(ctx.counts())
print(ctx.timings())
