"""Pin the oracle: oracle/ref_port.py and oracle/np_oracle.py against the reference's known answers,
the golden fixtures produced by the unmodified reference, and (dev container only) the live reference."""
import hashlib
import json
import os

import numpy as np
import pytest

import cases
from conftest import GOLDEN_CASES
from oracle import np_oracle, ref_port

HERE = os.path.dirname(os.path.abspath(__file__))


def _background(case):
    if case['bg_seqs'] is None:
        return None
    bg = ref_port.SetBackground(case['bg_Lmax'] + 1)
    bg.multiadd(case['bg_seqs'])
    return bg


@pytest.mark.parametrize('case', GOLDEN_CASES, ids=[c['name'] for c in GOLDEN_CASES])
def test_ref_port_reproduces_reference_graph(case, tmp_path):
    bg = _background(case)
    if bg is not None:
        assert len(bg) == case['bg_len']
        assert hashlib.sha256('\n'.join(sorted(bg.keys)).encode()).hexdigest() == case['bg_keys_sha256']
    seq_file = str(tmp_path / 'seq_list.txt')
    graph, parts = ref_port.homology_graph(list(case['parts']), bg, case['Lmax'] + 1, case['internal_repeats'], seq_file)
    assert list(graph.nodes()) == case['graph_nodes']
    assert [list(graph[n]) for n in graph.nodes()] == case['graph_adj']
    assert [pid for pid, _ in parts] == case['seq_file_ids']
    lines = open(seq_file).read().splitlines()
    assert lines == ['{},{}'.format(i, case['parts'][i].upper()) for i in case['seq_file_ids']]


@pytest.mark.parametrize('case', GOLDEN_CASES, ids=[c['name'] for c in GOLDEN_CASES])
def test_np_oracle_matches_reference_graph(case):
    k = case['Lmax'] + 1
    bg_keys = bg_K = None
    if case['bg_seqs'] is not None:
        bg_K = case['bg_Lmax'] + 1
        bg_keys = np_oracle.canonical_key_set(case['bg_seqs'], bg_K)
    parts = ref_port.uniquify(list(case['parts']))
    res = np_oracle.spec(parts, k, case['internal_repeats'], bg_keys, bg_K)
    nodes = set(res['ids'][res['flags'] == 0].tolist())
    assert nodes == set(case['graph_nodes'])
    ref_edges = set()
    for u, adj in zip(case['graph_nodes'], case['graph_adj']):
        for v in adj:
            ref_edges.add((min(u, v), max(u, v)))
    assert res['edges'] == ref_edges
    # the ordered stream, replayed through the literal tail, must give the reference's graph order
    graph = ref_port.assemble_graph(ref_port.clique_stream(_table_from_stream(np_oracle.ordered_stream(res))))
    assert list(graph.nodes()) == case['graph_nodes']
    assert [list(graph[n]) for n in graph.nodes()] == case['graph_adj']


def _table_from_stream(stream):
    """A dict whose popitem() order yields the stream (insert in reverse; keys are never read downstream)."""
    return {pos: list(members) for pos, members in enumerate(reversed(stream))}


def test_background_known_answer():
    """Reference docstring example nrpcalc.py:89-167: 12 canonical 16-mers, 10 after remove()."""
    with open(os.path.join(HERE, 'golden', 'background_golden.json')) as fh:
        kat = json.load(fh)
    K = kat['Lmax'] + 1
    doc_keys = ['AAAACTATCAAGCTAA', 'ACAGTAGAGACGAGTA', 'ACCTGACGTAATCGTC', 'ACGATTACGTCAGGTA', 'ATGAGATCGTAGCAAC',
                'ATGCTTAGTGCCATAC', 'CAGTACGAAAAGGCCC', 'CAGTAGAGACGAGTAA', 'CCAGTACGAAAAGGCC', 'GGTATGGCACTAAGCA',
                'GGTTGCTACGATCTCA', 'TAAAACTATCAAGCTA']          # printed at nrpcalc.py:130-141
    assert kat['keys_after_add'] == doc_keys and kat['len_after_add'] == 12 and kat['len_after_remove'] == 10
    bg = ref_port.SetBackground(K)
    bg.add(kat['add'])
    bg.multiadd(kat['multiadd'])
    assert len(bg) == 12
    assert sorted(bg.keys) == doc_keys
    for seq, expected in kat['contains_after_add'].items():
        assert (seq in bg) == expected
    keys = np_oracle.canonical_key_set([kat['add']] + kat['multiadd'], K)
    decoded = sorted(''.join('ACGT'[(int(v) >> (2 * (K - 1 - t))) & 3] for t in range(K)) for v in keys)
    assert decoded == doc_keys


def test_kmersetdb_selftest_known_answer():
    """Reference kmerSetDB.test() (kmerSetDB.py:332-423): 1000 random 100-mers at K=99 give 2000 keys."""
    import random
    rnd = random.Random(3)
    seqs = [''.join(rnd.choice('ATGC') for _ in range(100)) for _ in range(1000)]
    bg = ref_port.SetBackground(99)
    bg.multiadd(seqs)
    assert len(bg) == 2000
    assert all(bg.multicheck(seqs))


@pytest.mark.parametrize('seed', range(4))
@pytest.mark.parametrize('k', [6, 13, 17, 32])
@pytest.mark.parametrize('internal', [False, True])
def test_np_oracle_equals_ref_port(seed, k, internal):
    parts = cases.planted_parts(150, 100 + seed)
    bg_seqs = cases.background_seqs_for(parts, seed) if seed % 2 else None
    bg = bg_keys = None
    K = [12, 16, 20, 32][seed]
    if bg_seqs is not None:
        bg = ref_port.SetBackground(K)
        bg.multiadd(bg_seqs)
        bg_keys = np_oracle.canonical_key_set(bg_seqs, K)
    res = np_oracle.spec(ref_port.uniquify(list(parts)), k, internal, bg_keys, K)
    assert np_oracle.ordered_stream(res) == ref_port.ordered_stream(parts, bg, k, internal)


@pytest.mark.needs_reference
@pytest.mark.parametrize('seed', range(3))
def test_ref_port_against_live_reference(seed, scratch_cwd):
    import refshim
    if not refshim.reference_available():
        pytest.skip('reference tree not present (GPU box)')
    refshim.load_reference()
    from nrpcalc.base import hgraph as ref_hgraph
    parts = cases.planted_parts(200, 500 + seed)
    for k in (13, 17):
        for internal in (False, True):
            theirs = ref_hgraph.get_homology_graph(list(parts), None, k, internal, './seq_list.txt', False)
            mine, _ = ref_port.homology_graph(list(parts), None, k, internal)
            assert list(theirs.nodes()) == list(mine.nodes())
            assert [list(theirs[n]) for n in theirs.nodes()] == [list(mine[n]) for n in mine.nodes()]


@pytest.mark.parametrize('seed', range(3))
@pytest.mark.parametrize('k', [6, 13, 17, 32])
@pytest.mark.parametrize('threads', [1, 3])
def test_c_oracle_equals_np_oracle(seed, k, threads):
    """oracle/nrp_oracle.c (hash table in insertion order, like the reference's dict) against the numpy spec."""
    from oracle import c_oracle
    parts = cases.planted_parts(250, 300 + seed)
    bg_keys = np_oracle.canonical_key_set(cases.background_seqs_for(parts, seed), 14) if seed else None
    for internal in (False, True):
        uniq = [(pid, seq) for pid, seq in ref_port.uniquify(list(parts)) if len(seq) >= k]
        want = np_oracle.spec(uniq, k, internal, bg_keys, 14)
        seqs = [seq for _, seq in uniq]
        offsets = np.zeros(len(seqs) + 1, dtype=np.int64)
        np.cumsum([len(s) for s in seqs], out=offsets[1:])
        buf = np.frombuffer(''.join(seqs).encode(), dtype=np.uint8)
        got = c_oracle.build_table(buf, offsets, k, internal, bg_keys, 14, n_threads=threads)
        assert got['flags'].tolist() == want['flags'].tolist()
        assert got['n_records'] == want['n_records']
        assert got['last_single'].tolist() == want['last_single'].tolist()
        ids = want['ids']
        groups = sorted((int(got['rank_p'][g]), int(got['rank_j'][g]),
                         tuple(int(ids[m]) for m in got['members'][got['indptr'][g]:got['indptr'][g + 1]]))
                        for g in range(got['n_groups']))
        assert groups == sorted(want['groups'])


def test_c_oracle_slices_and_canonical_form():
    """The key-space slicing of the C oracle (memory bound for the full-size configs) changes nothing, and the canonical
    form / digest of tests/parity.py does not depend on the order in which a producer lists the groups."""
    from oracle import c_oracle
    import parity
    parts = cases.planted_parts(400, 4242) + cases.random_parts(300, 60, 3)
    k = 15
    uniq = [(pid, seq) for pid, seq in ref_port.uniquify(list(parts)) if len(seq) >= k]
    ids = np.array([pid for pid, _ in uniq], dtype=np.uint32)
    seqs = [seq for _, seq in uniq]
    offsets = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum([len(s) for s in seqs], out=offsets[1:])
    buf = np.frombuffer(''.join(seqs).encode(), dtype=np.uint8)
    for internal in (False, True):
        one = parity.from_c_oracle(c_oracle.build_table(buf, offsets, k, internal, n_threads=2), ids)
        many = parity.from_c_oracle(c_oracle.build_table(buf, offsets, k, internal, n_threads=3, n_slices=5), ids)
        assert parity.compare(one, many)['all'] and parity.digest(one) == parity.digest(many)
        want = np_oracle.spec(uniq, k, internal)
        assert one['edges'].tolist() == sorted((min(a, b) << 32) | max(a, b) for a, b in want['edges'])
        assert one['records'] == want['n_records'] and one['keys'].size == len(want['groups'])
        # a producer that lists the same groups in another order has the same canonical form
        raw = c_oracle.build_table(buf, offsets, k, internal)
        perm = np.random.default_rng(1).permutation(raw['n_groups'])
        indptr, members = parity.reorder_csr(raw['indptr'], raw['members'], perm)
        shuffled = parity.canonical(raw['flags'], indptr, members, raw['rank_j'][perm], raw['keys'][perm], raw['last_single'], None, ids,
                                    raw['n_records'])
        assert parity.compare(one, shuffled)['all']
        broken = dict(shuffled, last_single=shuffled['last_single'] + 1)
        assert not parity.compare(one, broken)['all']
