"""Wall clock of the PUBLIC Python API (nrpcalc_b200.finder) on random 100-bp parts, next to the literal port of the reference
(oracle/ref_port.py, graph build only) on the same list: what a user of nrpcalc.finder sees, host tail included.
    python tests/api_timing.py [n_parts ...]"""
import os
import random
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # test-side script: it times the oracle port next to the product
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import cases                     # noqa: E402
import nrpcalc_b200              # noqa: E402
from oracle import ref_port     # noqa: E402

os.chdir(tempfile.mkdtemp(prefix='nrp_api_'))
for n in [int(a) for a in sys.argv[1:]] or [10_000, 100_000]:
    parts = cases.random_parts(n, 100, 1)
    random.seed(7)
    nrpcalc_b200.finder(parts[:1000], 12, verbose=False)           # warm-up (context, buffers)
    random.seed(7)
    t0 = time.perf_counter()
    found = nrpcalc_b200.finder(list(parts), 12, verbose=False)
    t1 = time.perf_counter()
    line = 'finder(): {} parts, Lmax=12 -> toolbox {} in {:.3f} s'.format(n, len(found), t1 - t0)
    if n <= 20_000:
        t0 = time.perf_counter()
        ref_port.homology_graph(list(parts), None, 13, False)
        line += '; reference port, graph build only: {:.3f} s'.format(time.perf_counter() - t0)
    print(line, flush=True)
