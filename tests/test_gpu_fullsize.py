"""Full-size parity on the BASELINE shapes: the CUDA path's complete result (exclusion flags, every shared k-mer group with
its first window and its members in order, last unshared windows, distinct edges) against oracle/nrp_oracle.c on the SAME
batch -- configs 1 and 2 in full, config 3 in full WITH its 5-Mbp background.  (Config 4, 8e8 k-mers, takes the oracle
minutes: `python bench.py --config c4 --verify` keeps that line under profiles/.)"""
import os
import sys

import numpy as np
import pytest

import parity
from oracle import c_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ctx():
    from nrpcalc_b200 import _native
    context = _native.Context(0)
    yield context
    context.close()


@pytest.mark.parametrize('name', ['c1', 'c2', 'c3', 'dup'])
def test_baseline_config_against_c_oracle(ctx, name):
    import bench
    cfg = bench.CONFIGS[name]
    k = cfg['Lmax'] + 1
    buf, off = bench.make_workload(cfg)
    ids = np.arange(cfg['n'], dtype=np.uint32)[::-1].copy()
    bg_keys, bg_K = bench.background_keys(cfg, ctx)
    ctx.set_background(bg_keys, bg_K)
    try:
        ctx.load_parts(buf, off, ids)
        ctx.build(k, cfg['internal_repeats'], True)
        res = ctx.fetch(want_keys=True, copy=True)
    finally:
        ctx.set_background(None, 0)
    got = parity.canonical(res.flags, res.indptr, res.members, res.first_window, res.keys, res.last_single, res.edges, None,
                           res.counts['records'])
    want = c_oracle.build_table(buf, off, k, cfg['internal_repeats'], bg_keys, bg_K, n_threads=os.cpu_count() or 1, want_groups=True,
                                n_slices=bench.VERIFY_SLICES[name])
    ref = parity.from_c_oracle(want, ids)
    cmp = parity.compare(got, ref)
    assert cmp['all'], cmp
    assert parity.digest(got) == parity.digest(ref)
    if name == 'c3':
        assert int(got['excluded'].sum()) > cfg['n'] // 10          # the background really excludes parts
    if name == 'dup':
        assert res.counts['fallbacks'] >= 1                        # the fixed regions overflowed: exact path + global sort ran
