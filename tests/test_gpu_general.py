"""GPU parity of the general path (nrp_build_generic: any byte alphabet, any k) against the string-level oracle, the fixtures of
the unmodified reference on RNA / other symbols / windows longer than 32 bases, and the 2-bit build on inputs both accept."""
import random

import numpy as np
import pytest

import cases
import fake_device
import nrpcalc_b200
from conftest import GOLDEN_GENERAL
from nrpcalc_b200 import _native, hgraph
from oracle import np_oracle, ref_port

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ctx():
    context = _native.Context(0)
    yield context
    context.close()


def _load(ctx, uniq):
    seqs = [s for _, s in uniq]
    ids = np.array([pid for pid, _ in uniq], dtype=np.uint32)
    offsets = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum([len(s) for s in seqs], out=offsets[1:])
    ctx.load_parts(np.frombuffer(''.join(seqs).encode(), dtype=np.uint8), offsets, ids)
    return ids


def _compare(res, want, uniq):
    ids = [pid for pid, _ in uniq]
    assert (res.flags != 0).tolist() == (want['flags'] != 0).tolist()
    assert res.counts['records'] == want['n_records']
    got = sorted((int(res.members[res.indptr[g]]), int(res.first_window[g]), tuple(int(m) for m in res.members[res.indptr[g]:res.indptr[g + 1]]))
                 for g in range(res.counts['groups']))
    assert got == sorted((p, j, tuple(m)) for p, j, m in want['groups'])
    assert res.last_single.tolist() == want['last_single'].tolist()
    assert sorted(zip(res.edges[0].tolist(), res.edges[1].tolist())) == want['edges']
    assert res.counts['excluded'] == int((want['flags'] != 0).sum())
    assert len(ids) == res.flags.size


def _mixed_parts(n, seed, lo=40, hi=120, frag_lo=18, frag_hi=26):
    rnd = random.Random(seed)
    out = []
    for i, s in enumerate(cases.planted_parts(n, seed, lo=lo, hi=hi, frag_lo=frag_lo, frag_hi=frag_hi)):
        s = s.upper()
        if i % 6 == 0 and s:
            at = rnd.randrange(len(s))
            s = s[:at] + rnd.choice('NRY-') + s[at + 1:]
        if i % 4 == 1:
            s = s.replace('T', 'U')
        out.append(s)
    return out


@pytest.mark.parametrize('k', [6, 13, 16, 17, 33, 40])
@pytest.mark.parametrize('internal', [False, True])
def test_generic_build_against_string_oracle(ctx, k, internal):
    parts = _mixed_parts(500, 100 + k, lo=max(40, k), hi=max(120, 3 * k), frag_lo=k + 2, frag_hi=k + 12)
    uniq = [(pid, s) for pid, s in ref_port.uniquify(list(parts)) if len(s) >= k]
    _load(ctx, uniq)
    assert ctx.build_generic(k, internal, want_edges=True) == 0
    _compare(ctx.fetch(), fake_device.string_spec(uniq, k, internal), uniq)


@pytest.mark.parametrize('k,K', [(16, 16), (13, 20), (17, 12), (40, 16)])
def test_generic_build_with_background_and_preset(ctx, k, K):
    parts = _mixed_parts(400, 7 * k + K, lo=max(50, k), hi=max(130, 3 * k), frag_lo=k + 2, frag_hi=k + 12)
    bg_seqs = [s.replace('U', 'T') for s in cases.background_seqs_for([p for p in parts if set(p) <= set('ACGTU')], 3)]
    bg_keys = np_oracle.canonical_key_set([s for s in bg_seqs if set(s) <= set('ACGT')], K)
    uniq = [(pid, s) for pid, s in ref_port.uniquify(list(parts)) if len(s) >= k]
    preset = np.zeros(len(uniq), dtype=np.uint8)
    preset[3::17] = 4
    _load(ctx, uniq)
    ctx.set_background(bg_keys, K)
    try:
        ctx.build_generic(k, False, want_edges=True, preset_flags=preset)
        want = fake_device.string_spec(uniq, k, False, fake_device._Keys2bit(bg_keys, K), preset)
        res = ctx.fetch()
        assert (res.flags & 4).any()
        _compare(res, want, uniq)
    finally:
        ctx.set_background(None, 0)


@pytest.mark.parametrize('k,internal', [(13, False), (17, True), (24, False)])
def test_generic_equals_2bit_build_on_plain_dna(ctx, k, internal):
    """20k A/C/G/T parts through both builds: flags, groups with ranks, last unshared windows and edges must be the same."""
    parts = cases.random_parts(16000, 100, k) + [p[:100] for p in cases.planted_parts(4000, k, lo=100, hi=130, p_short=0, p_dup=0)]
    uniq = ref_port.uniquify(list(parts))
    _load(ctx, uniq)
    ctx.build(k, internal, want_edges=True)
    a = ctx.fetch(copy=True)
    ctx.build_generic(k, internal, want_edges=True)
    b = ctx.fetch(copy=True)

    def canon(r):
        groups = sorted((int(r.members[r.indptr[g]]), int(r.first_window[g]), tuple(r.members[r.indptr[g]:r.indptr[g + 1]].tolist()))
                        for g in range(r.counts['groups']))
        return ((r.flags != 0).tolist(), groups, r.last_single.tolist(), sorted(zip(r.edges[0].tolist(), r.edges[1].tolist())), r.counts['records'])
    assert canon(a) == canon(b)
    assert a.counts['groups'] > 100 and (internal or k > 20 or (a.flags != 0).any())


def test_hash_collisions_are_detected_not_trusted(ctx):
    """With the hash cut to a few bits different windows collide inside parts and between parts: every attempt must be refused
    (status bits), never answered."""
    uniq = ref_port.uniquify(cases.random_parts(3000, 100, 5))
    _load(ctx, uniq)
    for bits in (6, 16, 22):
        with pytest.raises(_native.NativeError, match='collisions'):
            ctx.build_generic(13, False, hash_bits=bits, attempts=3)
    assert ctx.build_generic(13, False) == 0                     # the full hash: accepted at the first seed
    _compare(ctx.fetch(), fake_device.string_spec(uniq, 13, False), uniq)


def test_limits_are_reported(ctx):
    uniq = [(0, 'ACGU' * 1100), (1, 'ACGGUCA' * 10)]
    _load(ctx, uniq)
    with pytest.raises(_native.NativeError, match='4096 windows'):
        ctx.build_generic(13, False)
    uniq = [(0, 'ACGU' * 1000), (1, 'ACGGUCAU' * 10)]
    _load(ctx, uniq)
    ctx.build_generic(13, True)
    _compare(ctx.fetch(), fake_device.string_spec(uniq, 13, True), uniq)


def _background(case):
    if case['bg_seqs'] is None:
        return None
    bg = nrpcalc_b200.background(path='./bg_store/', Lmax=case['bg_Lmax'], verbose=False)
    bg.multiadd(case['bg_seqs'])
    return bg


@pytest.mark.parametrize('case', GOLDEN_GENERAL, ids=[c['name'] for c in GOLDEN_GENERAL])
def test_general_golden_graph_and_toolbox(case, scratch_cwd):
    bg = _background(case)
    graph = hgraph.get_homology_graph(list(case['parts']), bg, case['Lmax'] + 1, case['internal_repeats'], './seq.txt', False)
    assert list(graph.nodes()) == case['graph_nodes']
    assert [list(graph[n]) for n in graph.nodes()] == case['graph_adj']
    assert hgraph.last_build_info['counts']['generic_builds'] >= 1
    for vc, by_seed in case['toolbox'].items():
        for seed, expected_ids in by_seed.items():
            random.seed(int(seed))
            found = nrpcalc_b200.finder(seq_list=list(case['parts']), Lmax=case['Lmax'], internal_repeats=case['internal_repeats'],
                                        background=bg, vercov=vc, output_file=None, verbose=False)
            assert list(found.keys()) == expected_ids, (vc, seed)
            assert all(found[i] == case['parts'][i].upper() for i in expected_ids)
    if bg is not None:
        bg.drop()
