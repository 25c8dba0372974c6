"""GPU parity of the raw C ABI (libnrpb200.so) against the numpy oracle on seeded inputs."""
import numpy as np
import pytest

import cases
from oracle import np_oracle, ref_port

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ctx():
    from nrpcalc_b200 import _native
    context = _native.Context(0)
    yield context
    context.close()


def _long_parts(parts, k):
    """(processing-ordered [(id, SEQ)]) restricted to parts with len >= k, as the C ABI expects."""
    uniq = ref_port.uniquify(list(parts))
    return [(pid, seq) for pid, seq in uniq if len(seq) >= k]


def _run(ctx, uniq, k, internal, bg_keys=None, bg_K=None):
    seqs = [s for _, s in uniq]
    ids = np.array([pid for pid, _ in uniq], dtype=np.uint32)
    offsets = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum([len(s) for s in seqs], out=offsets[1:])
    ascii_buf = np.frombuffer(''.join(seqs).encode(), dtype=np.uint8)
    ctx.set_background(bg_keys, bg_K)
    ctx.load_parts(ascii_buf, offsets, ids)
    ctx.build(k, internal, want_edges=True)
    return ctx.fetch()


def _compare(res, uniq, k, internal, bg_keys=None, bg_K=None):
    want = np_oracle.spec(uniq, k, internal, bg_keys, bg_K)
    ids = want['ids']
    assert (res.flags != 0).tolist() == (want['flags'] != 0).tolist()
    # the reason bits are a subset: a part already excluded for one reason is not examined for the others
    assert ((res.flags & ~want['flags']) == 0).all()
    assert res.counts['records'] == want['n_records']
    got_groups = sorted((int(res.members[res.indptr[g]]), int(res.first_window[g]),
                         tuple(int(ids[m]) for m in res.members[res.indptr[g]:res.indptr[g + 1]]))
                        for g in range(res.counts['groups']))
    assert got_groups == sorted(want['groups'])
    assert res.last_single.tolist() == want['last_single'].tolist()
    got_edges = sorted(zip(res.edges[0].tolist(), res.edges[1].tolist()))       # distinct pairs, order unspecified
    assert got_edges == sorted(want['edges'])
    assert len(got_edges) == len(set(got_edges))
    assert res.counts['excluded'] == int((want['flags'] != 0).sum())


@pytest.mark.parametrize('seed', range(3))
@pytest.mark.parametrize('k', [6, 13, 16, 17, 25, 32])
@pytest.mark.parametrize('internal', [False, True])
def test_planted_parts(ctx, seed, k, internal):
    uniq = _long_parts(cases.planted_parts(400, seed), k)
    _compare(_run(ctx, uniq, k, internal), uniq, k, internal)


@pytest.mark.parametrize('seed', range(2))
@pytest.mark.parametrize('k,K', [(16, 16), (16, 12), (13, 20), (17, 17), (20, 32)])
def test_background(ctx, seed, k, K):
    parts = cases.planted_parts(400, 20 + seed)
    bg_keys = np_oracle.canonical_key_set(cases.background_seqs_for(parts, seed), K)
    uniq = _long_parts(parts, k)
    _compare(_run(ctx, uniq, k, False, bg_keys, K), uniq, k, False, bg_keys, K)
    ctx.set_background(None, 0)


@pytest.mark.parametrize('internal', [False, True])
def test_palindromes(ctx, internal):
    uniq = _long_parts(cases.palindrome_parts(0, n_parts=200), 6)
    _compare(_run(ctx, uniq, 6, internal), uniq, 6, internal)


@pytest.mark.parametrize('n,length,k', [(10000, 100, 13), (20000, 100, 17), (3000, 200, 16)])
def test_random_uniform(ctx, n, length, k):
    """BASELINE shapes (config 1 in full; configs 2 and 3 scaled down to what the oracle does in seconds)."""
    uniq = ref_port.uniquify(cases.random_parts(n, length, 1))
    res = _run(ctx, uniq, k, False)
    _compare(res, uniq, k, False)


def test_variable_length(ctx):
    uniq = _long_parts(cases.buffer_to_list(*cases.random_varlen_buffer(5000, 50, 300, 5)), 25)
    _compare(_run(ctx, uniq, 25, False), uniq, 25, False)
    uniq = _long_parts(cases.buffer_to_list(*cases.random_varlen_buffer(5000, 20, 90, 6)), 11)
    _compare(_run(ctx, uniq, 11, False), uniq, 11, False)


def test_forced_paths(ctx):
    """Tiny filter capacity -> every leaf bucket takes the oversized path; one level bit -> deep buckets."""
    uniq = _long_parts(cases.planted_parts(600, 77), 13)
    try:
        ctx.set_option('filter_cap', 16)
        _compare(_run(ctx, uniq, 13, False), uniq, 13, False)
        ctx.set_option('filter_cap', 0)
        ctx.set_option('leaf_target', 40)
        _compare(_run(ctx, uniq, 13, True), uniq, 13, True)
        ctx.set_option('max_level_bits', 2)
        _compare(_run(ctx, uniq, 13, True), uniq, 13, True)
        ctx.set_option('max_level_bits', 0)
        ctx.set_option('leaf_target', 0)
        ctx.set_option('pass_records', 9000)          # exact path: several passes over slices of the hash space
        ctx.set_option('direct', 0)
        res = _run(ctx, uniq, 13, False)
        assert res.counts['runs'] > 1
        _compare(res, uniq, 13, False)
        _compare(_run(ctx, uniq, 13, True), uniq, 13, True)
        ctx.set_option('pass_records', 0)
        ctx.set_option('direct', -1)
        ctx.set_option('coop_group', 0)               # grouping as a kernel sequence instead of one cooperative kernel
        _compare(_run(ctx, uniq, 13, False), uniq, 13, False)
        _compare(_run(ctx, uniq, 13, True), uniq, 13, True)
        ctx.set_option('coop_group', -1)
        ctx.set_option('group_walk', 0)               # grouping with nine grid barriers (compaction) instead of the walk variant
        _compare(_run(ctx, uniq, 13, False), uniq, 13, False)
        _compare(_run(ctx, uniq, 13, True), uniq, 13, True)
        ctx.set_option('group_walk', -1)
        ctx.set_option('filter_wide', 0)              # two 512-thread filter CTAs per SM instead of three 256-thread ones
        _compare(_run(ctx, uniq, 13, False), uniq, 13, False)
        big = _long_parts(cases.planted_parts(800, 5) + cases.random_parts(12000, 100, 6), 17)
        _compare(_run(ctx, big, 17, False), big, 17, False)
        ctx.set_option('filter_wide', -1)
        ctx.set_option('uniform_kernel', 0)           # byte-position extraction kernel on a batch of one length
        uni = _long_parts(cases.random_parts(6000, 90, 12) + [p[:90] for p in cases.planted_parts(900, 8, lo=90, hi=120, p_short=0, p_dup=0)], 17)
        res = _run(ctx, uni, 17, False)
        assert res.counts['uniform_kernel'] == 0
        _compare(res, uni, 17, False)
        ctx.set_option('uniform_kernel', -1)
        res = _run(ctx, uni, 17, False)
        assert res.counts['uniform_kernel'] == 1
        _compare(res, uni, 17, False)
        ctx.set_option('filter_tiny', 1)              # 256-thread filter CTAs (leaves that average <= 1400 records)
        _compare(_run(ctx, uniq, 13, False), uniq, 13, False)
        _compare(_run(ctx, uniq, 13, True), uniq, 13, True)
        ctx.set_option('leaf_target', 1300)
        big = _long_parts(cases.planted_parts(800, 5) + cases.random_parts(12000, 100, 6), 17)
        res = _run(ctx, big, 17, False)
        assert res.counts['buckets'] >= 512 and res.counts['packed'] == 1
        _compare(res, big, 17, False)
        ctx.set_option('leaf_target', 0)
        ctx.set_option('filter_tiny', 0)
        ctx.set_option('no_bloom', 1)                 # exact table on every record instead of the two-stage filter
        _compare(_run(ctx, uniq, 13, False), uniq, 13, False)
        ctx.set_option('no_bloom', 0)
        ctx.set_option('no_window_carry', 1)          # ranks from the ASCII scan instead of the record value
        _compare(_run(ctx, uniq, 13, False), uniq, 13, False)
        uniq17 = _long_parts(cases.planted_parts(600, 77), 17)
        _compare(_run(ctx, uniq17, 17, True), uniq17, 17, True)
    finally:
        ctx.set_option('group_walk', -1)
        ctx.set_option('filter_wide', -1)
        ctx.set_option('uniform_kernel', -1)
        ctx.set_option('filter_tiny', 0)
        ctx.set_option('coop_group', -1)
        ctx.set_option('direct', -1)
        ctx.set_option('no_window_carry', 0)
        ctx.set_option('no_bloom', 0)
        ctx.set_option('pass_records', 0)
        ctx.set_option('filter_cap', 0)
        ctx.set_option('leaf_target', 0)
        ctx.set_option('max_level_bits', 0)


def test_direct_and_exact_paths_agree(ctx):
    """Default = fixed regions without a histogram pass; direct=0 = exact histogram path.  Both must match the oracle."""
    for k, internal in ((13, False), (17, True), (25, False)):
        uniq = _long_parts(cases.planted_parts(1500, 5) + cases.random_parts(3000, 100, 3), k)
        try:
            res = _run(ctx, uniq, k, internal)
            assert res.counts['direct'] == 1 and res.counts['fallbacks'] == 0
            # one 64-bit word per record whenever the 2k hash bits minus the level-1 bits, plus 20 value bits, fit 64 bits
            assert res.counts['packed'] == 1
            _compare(res, uniq, k, internal)
            ctx.set_option('packed', 0)               # (key, value) arrays on the direct path
            res = _run(ctx, uniq, k, internal)
            assert res.counts['direct'] == 1 and res.counts['packed'] == 0
            _compare(res, uniq, k, internal)
            ctx.set_option('packed', -1)
            ctx.set_option('direct', 0)
            res = _run(ctx, uniq, k, internal)
            assert res.counts['direct'] == 0
            _compare(res, uniq, k, internal)
        finally:
            ctx.set_option('direct', -1)
            ctx.set_option('packed', -1)


def test_unpacked_when_key_and_value_exceed_64_bits(ctx):
    uniq = _long_parts(cases.planted_parts(800, 11) + cases.random_parts(2000, 100, 6), 32)
    res = _run(ctx, uniq, 32, False)
    assert res.counts['direct'] == 1 and res.counts['packed'] == 0
    _compare(res, uniq, 32, False)


@pytest.mark.parametrize('warp', [1, 0])
def test_warp_private_filter(ctx, warp):
    """The warp-private duplicate filter (filter_warp.cuh: leaves of a few hundred records, one warp each) against the CTA-per-leaf
    variants and the oracle: random parts (both map sizes), planted repeats (leaves with more suspects than a warp takes are passed
    on whole and the global sort runs), 32-bit and 64-bit keys, internal repeats on and off, a background."""
    try:
        ctx.set_option('filter_warp', warp)
        seen = set()
        # (k = 21 / 23: the k-mer bits of a record start below bit 32 of the word -- the kernel's other addressing)
        for n, length, k, internal, target in ((20000, 100, 17, False, 0), (60000, 100, 17, False, 0), (9000, 100, 13, False, 0),
                                               (30000, 60, 16, True, 0), (20000, 100, 21, False, 0), (40000, 100, 17, False, 300),
                                               (15000, 90, 23, True, 250), (2500, 33, 17, False, 100),
                                               (25000, 100, 13, False, 0)):      # 6 % shared records: the variant with room for 192 suspects
            ctx.set_option('warp_leaf_target', target)
            uniq = ref_port.uniquify(cases.random_parts(n, length, n + k))
            res = _run(ctx, uniq, k, internal)
            seen.add(res.counts['filter_warp'])
            assert (res.counts['filter_warp'] != 0) == bool(warp)
            _compare(res, uniq, k, internal)
        ctx.set_option('warp_leaf_target', 0)
        if warp:
            assert seen == {1, 2, 3}              # standard, small-map and many-suspect variants all ran
        planted = _long_parts(cases.planted_parts(3000, 12) + cases.random_parts(12000, 100, 13), 17)
        res = _run(ctx, planted, 17, False)
        _compare(res, planted, 17, False)
        res = _run(ctx, planted, 17, True)
        _compare(res, planted, 17, True)
        bg_keys = np_oracle.canonical_key_set(cases.background_seqs_for([s for _, s in planted[:400]], 3), 17)
        res = _run(ctx, planted, 17, False, bg_keys, 17)
        _compare(res, planted, 17, False, bg_keys, 17)
    finally:
        ctx.set_option('filter_warp', -1)
        ctx.set_option('warp_leaf_target', 0)
        ctx.set_background(None, 0)


def test_direct_overflow_falls_back_to_exact(ctx):
    """A k-mer shared by thousands of parts overflows its fixed region: the batch is repeated on the exact path."""
    import random
    rnd = random.Random(3)
    common = ''.join(rnd.choice('ACGT') for _ in range(14))
    parts = [common + ''.join(rnd.choice('ACGT') for _ in range(36)) for _ in range(1200)]
    uniq = _long_parts(parts, 13)
    try:
        ctx.set_option('max_pairs', 1 << 40)
        res = _run(ctx, uniq, 13, True)
        assert res.counts['fallbacks'] == 1 and res.counts['direct'] == 0
        want = np_oracle.spec(uniq, 13, True)
        assert res.counts['records'] == want['n_records']
        assert res.counts['groups'] == len(want['groups'])
        assert sorted(zip(res.edges[0].tolist(), res.edges[1].tolist())) == sorted(want['edges'])
        assert res.last_single.tolist() == want['last_single'].tolist()
    finally:
        ctx.set_option('max_pairs', 0)


def test_direct_overflow_after_folded_flags(ctx):
    """The same overflow when the extraction kernel also computed the palindrome / background flags (even k, K == k): the exact
    path that repeats the batch must find those flags complete."""
    import random
    rnd = random.Random(4)
    common = ''.join(rnd.choice('ACGT') for _ in range(15))
    parts = [common + ''.join(rnd.choice('ACGT') for _ in range(35)) for _ in range(1200)]
    for i in range(3, 1200, 50):
        half = ''.join(rnd.choice('ACGT') for _ in range(7))
        parts[i] = parts[i][:20] + half + cases.rc(half) + parts[i][34:]
    bg_keys = np_oracle.canonical_key_set([p[30:50] for p in parts[::40]], 14)
    uniq = _long_parts(parts, 14)
    try:
        ctx.set_option('max_pairs', 1 << 40)
        res = _run(ctx, uniq, 14, False, bg_keys, 14)
        assert res.counts['fallbacks'] == 1 and res.counts['direct'] == 0 and res.counts['uniform_kernel'] == 2
        want = np_oracle.spec(uniq, 14, False, bg_keys, 14)
        assert (res.flags != 0).tolist() == (want['flags'] != 0).tolist() and (res.flags & 4).any() and (res.flags & 2).any()
        assert res.counts['records'] == want['n_records']
        assert res.counts['groups'] == len(want['groups'])
        assert sorted(zip(res.edges[0].tolist(), res.edges[1].tolist())) == sorted(want['edges'])
        assert res.last_single.tolist() == want['last_single'].tolist()
    finally:
        ctx.set_option('max_pairs', 0)
        ctx.set_background(None, 0)


def test_deep_direct_levels(ctx):
    """Tiny leaf target: two 9-bit levels on the direct path."""
    uniq = _long_parts(cases.planted_parts(800, 9) + cases.random_parts(4000, 100, 4), 17)
    try:
        ctx.set_option('leaf_target', 1)
        res = _run(ctx, uniq, 17, False)
        assert res.counts['buckets'] == 1 << 18
        _compare(res, uniq, 17, False)
    finally:
        ctx.set_option('leaf_target', 0)


def test_empty_and_tiny(ctx):
    for parts in ([], ['ACGTACGTACGTA'], ['ACGTACGTACGTA', 'ACGTACGTACGTA'.lower(), 'TTTTTTTTTTTTTTTT']):
        uniq = _long_parts(parts, 13)
        _compare(_run(ctx, uniq, 13, True), uniq, 13, True)


def test_streamed_staging(ctx):
    """nrp_load_parts_streamed: the copy is enqueued in pieces and the build extracts each piece as it lands."""
    import torch
    for k, internal, parts in ((13, False, cases.planted_parts(1500, 21) + cases.random_parts(3000, 100, 8)),
                               (16, False, cases.palindrome_parts(3, n_parts=400, k=16) + cases.random_parts(2000, 100, 9)),
                               (25, True, cases.buffer_to_list(*cases.random_varlen_buffer(4000, 50, 300, 12)))):
        uniq = _long_parts(parts, k)
        seqs = [s for _, s in uniq]
        ids = np.array([pid for pid, _ in uniq], dtype=np.uint32)
        offsets = np.zeros(len(seqs) + 1, dtype=np.int64)
        np.cumsum([len(s) for s in seqs], out=offsets[1:])
        pinned = torch.from_numpy(np.frombuffer(''.join(seqs).encode(), dtype=np.uint8).copy()).pin_memory()
        for chunks in (2, 7, 64):
            ctx.set_background(None, 0)
            ctx.load_parts(pinned.numpy(), offsets, ids, chunks=chunks)
            ctx.build(k, internal, want_edges=True)
            _compare(ctx.fetch(), uniq, k, internal)
        ctx.load_parts(pinned.numpy(), offsets, ids, chunks=5)         # streamed, then the exact path and a second build
        ctx.set_option('direct', 0)
        try:
            ctx.build(k, internal, want_edges=True)
            _compare(ctx.fetch(), uniq, k, internal)
        finally:
            ctx.set_option('direct', -1)
        ctx.build(k, internal, want_edges=True)
        _compare(ctx.fetch(), uniq, k, internal)


def test_uniform_streamed_staging(ctx):
    """nrp_load_parts_uniform: parts of one length, no offsets array, ids page-locked or pageable, copies only enqueued."""
    import torch
    from nrpcalc_b200 import _native
    for k, internal, parts in ((13, False, cases.planted_parts(1500, 23) + cases.random_parts(3000, 100, 18)),
                               (16, False, cases.palindrome_parts(5, n_parts=400, k=16) + cases.random_parts(2000, 100, 19)),
                               (17, True, cases.random_parts(5000, 100, 20))):
        uniq = [(pid, s) for pid, s in _long_parts(parts, k) if len(s) == 100]
        seqs = [s for _, s in uniq]
        ids = np.array([pid for pid, _ in uniq], dtype=np.uint32)
        pinned = torch.from_numpy(np.frombuffer(''.join(seqs).encode(), dtype=np.uint8).copy()).pin_memory()
        pinned_ids = torch.from_numpy(ids.copy()).pin_memory()
        for chunks, id_array in ((1, ids), (3, ids), (8, pinned_ids.numpy()), (64, pinned_ids.numpy())):
            ctx.set_background(None, 0)
            ctx.load_parts_uniform(pinned.numpy(), len(seqs), 100, id_array, chunks=chunks)
            ctx.build(k, internal, want_edges=True)
            _compare(ctx.fetch(), uniq, k, internal)
        ctx.load_parts_uniform(pinned.numpy(), len(seqs), 100, pinned_ids.numpy(), chunks=4)     # a second build of the same batch
        ctx.build(k, internal, want_edges=True)
        ctx.build(k, internal, want_edges=True)
        _compare(ctx.fetch(), uniq, k, internal)
    with pytest.raises(_native.NativeError):
        ctx.load_parts_uniform(np.zeros(16, dtype=np.uint8), 4, 0, None, chunks=2)


def _uniform_with_repeats(n, length, seed, k):
    """n parts of ONE length with planted shared / internal / reverse-complement stretches (all stay `length` long)."""
    import random
    rnd = random.Random(seed)
    parts = cases.random_parts(n, length, seed)
    span = min(length, k + rnd.randint(0, 6))
    for i in range(0, n, 3):
        src = parts[rnd.randrange(n)]
        a, b = rnd.randint(0, length - span), rnd.randint(0, length - span)
        piece = src[a:a + span]
        if rnd.random() < 0.4:
            piece = cases.rc(piece)
        parts[i] = parts[i][:b] + piece + parts[i][b + span:]
    return parts


@pytest.mark.parametrize('length,k', [(13, 13), (14, 13), (20, 13), (21, 13), (29, 13), (100, 16), (40, 6), (64, 32), (33, 32), (100, 32),
                                      (300, 25), (57, 17), (2000, 17), (4111, 16), (4200, 17)])
@pytest.mark.parametrize('internal', [False, True])
def test_uniform_kernel_shapes(ctx, length, k, internal):
    """The part-aligned extraction kernel (extract.cuh) at its corners: one window per part, chunks that end inside the last
    thread of a part, 32-bit and 64-bit keys with no free low bits (k = 16, 32), even k with palindromes (flag pass feeds the
    kernel), 4096 windows per part (the largest it takes) and beyond (byte-position kernel)."""
    n = 3000 if length <= 300 else 60
    uniq = _long_parts(_uniform_with_repeats(n, length, 100 + length + k, k), k)
    res = _run(ctx, uniq, k, internal)
    windows = length - k + 1
    folded = 2 if (k % 2 == 0 and not internal) else 1          # 2: the kernel also tested its windows for palindromes
    assert res.counts['uniform_kernel'] == (folded if windows <= 4096 and res.counts['packed'] else 0)
    _compare(res, uniq, k, internal)


def test_uniform_kernel_with_background(ctx):
    parts = _uniform_with_repeats(4000, 80, 5, 16)
    bg_keys = np_oracle.canonical_key_set(cases.background_seqs_for(parts, 3), 14)
    uniq = _long_parts(parts, 16)
    res = _run(ctx, uniq, 16, False, bg_keys, 14)
    assert res.counts['uniform_kernel'] == 2 and (res.flags & 4).any()          # 2: k = 16 is even, the kernel tests for palindromes itself
    _compare(res, uniq, 16, False, bg_keys, 14)
    ctx.set_background(None, 0)


@pytest.mark.parametrize('k,K', [(16, 16), (6, 6), (17, 17), (20, 20), (32, 32), (12, 12), (16, 14), (14, 16)])
@pytest.mark.parametrize('fold', [1, 0])
def test_uniform_kernel_folded_flags(ctx, k, K, fold):
    """Palindrome test and background probe (K == k) inside the part-aligned extraction kernel (FOLD, extract.cuh) against the
    separate flag pass and the oracle: planted palindromic windows (even k), a background cut from the parts themselves (hits at
    the first / last window of a part and in the middle), background lengths that differ from k (probe stays a pass of its own)."""
    import random
    rnd = random.Random(1000 * k + K)
    length = max(k, K) + 37
    parts = _uniform_with_repeats(3000, length, 7 * k + K, k)
    if k % 2 == 0:
        for i in range(5, len(parts), 41):                         # a palindromic k-window at a random position
            half = cases.random_parts(1, k // 2, i)[0]
            at = rnd.choice([0, length - k, rnd.randint(0, length - k)])
            parts[i] = parts[i][:at] + half + cases.rc(half) + parts[i][at + k:]
    bg_seqs = cases.background_seqs_for(parts, 3)
    bg_seqs += [parts[i][:K] for i in range(0, len(parts), 97)] + [parts[i][-K:] for i in range(1, len(parts), 89)]
    bg_keys = np_oracle.canonical_key_set(bg_seqs, K)
    uniq = _long_parts(parts, k)
    try:
        ctx.set_option('fold_flags', fold)
        for internal in (False, True):
            res = _run(ctx, uniq, k, internal, bg_keys, K)
            expect_fold = fold and (K == k or (k % 2 == 0 and not internal))
            # (records of a k-mer too short for the packed word -- k = 6 -- take the byte-position kernel and the separate flag passes)
            assert res.counts['uniform_kernel'] == ((2 if expect_fold else 1) if res.counts['packed'] else 0)
            assert (res.flags & 4).any()
            _compare(res, uniq, k, internal, bg_keys, K)
        res = _run(ctx, uniq, k, False)                                # no background at all
        assert res.counts['uniform_kernel'] == ((2 if fold and k % 2 == 0 else 1) if res.counts['packed'] else 0)
        _compare(res, uniq, k, False)
    finally:
        ctx.set_option('fold_flags', -1)
        ctx.set_background(None, 0)


@pytest.mark.parametrize('internal', [False, True])
def test_long_key_runs_fall_back_from_the_walk_grouping(ctx, internal):
    """One k-mer shared by ~1500 records (more than the walk variant of the grouping stage follows): the stage is repeated with
    the nine-barrier kernel; a second shared stretch of ~200 records stays inside the walk limit."""
    import random
    rnd = random.Random(4)
    parts = cases.random_parts(4000, 60, 31)
    motif, short_motif = cases.random_parts(1, 20, 32)[0], cases.random_parts(1, 18, 33)[0]
    for i in range(0, 1500):
        at = rnd.randint(0, 40)
        parts[i] = parts[i][:at] + motif + parts[i][at + 20:]
    for i in range(2000, 2200):
        parts[i] = parts[i][:5] + short_motif + parts[i][23:]
    uniq = _long_parts(parts, 17)
    res = _run(ctx, uniq, 17, internal)
    _compare(res, uniq, 17, internal)


@pytest.mark.parametrize('k,internal', [(13, False), (17, True), (25, False)])
def test_early_fetch_gives_the_same_host_result(ctx, k, internal):
    """Option early_fetch (nrp_build starts copying what is final after the grouping stage into the page-locked result buffer,
    nrp_result_host moves only the rest): the same arrays as the plain fetch -- also when a build's early fetch is never collected
    (the next load / build waits for it) and when the buffer has to grow for the next batch."""
    def fetched(uniq, early):
        seqs = [s for _, s in uniq]
        ids = np.array([pid for pid, _ in uniq], dtype=np.uint32)
        offsets = np.zeros(len(seqs) + 1, dtype=np.int64)
        np.cumsum([len(s) for s in seqs], out=offsets[1:])
        ctx.load_parts(np.frombuffer(''.join(seqs).encode(), dtype=np.uint8), offsets, ids)
        ctx.build(k, internal, want_edges=True, early_fetch=early)
        return ctx.fetch(want_keys=True, copy=True)

    def groups_of(r):          # (the order of the groups differs from build to build: leaves finish in any order)
        return sorted((int(r.keys[g]), int(r.first_window[g]), tuple(r.members[r.indptr[g]:r.indptr[g + 1]].tolist()))
                      for g in range(r.counts['groups']))

    def same(a, b):
        assert a.counts['groups'] == b.counts['groups'] and a.counts['members'] == b.counts['members']
        assert np.array_equal(a.flags, b.flags) and np.array_equal(a.last_single, b.last_single)
        assert a.indptr[0] == 0 and a.indptr[-1] == a.counts['members'] and a.indptr.size == a.counts['groups'] + 1
        assert groups_of(a) == groups_of(b)
        assert sorted(zip(a.edges[0].tolist(), a.edges[1].tolist())) == sorted(zip(b.edges[0].tolist(), b.edges[1].tolist()))

    try:
        small = _long_parts(cases.planted_parts(900, 5) + cases.random_parts(2000, 100, 6), k)
        large = _long_parts(cases.planted_parts(4000, 7) + cases.random_parts(30000, 100, 8), k)
        plain_small, plain_large = fetched(small, False), fetched(large, False)
        same(fetched(small, True), plain_small)
        _compare(fetched(small, True), small, k, internal)
        same(fetched(large, True), plain_large)                  # larger batch: the result buffer grows
        ctx.build(k, internal, want_edges=True, early_fetch=True)    # an early fetch nobody collects ...
        ctx.build(k, internal, want_edges=True, early_fetch=True)    # ... the next build waits for it
        same(ctx.fetch(want_keys=True, copy=True), plain_large)
        ctx.build(k, internal, want_edges=True, early_fetch=True)    # ... and so does the next load
        same(fetched(small, True), plain_small)
        same(fetched(small, False), plain_small)
    finally:
        ctx.set_option('early_fetch', 0)
        ctx._early_fetch = False
