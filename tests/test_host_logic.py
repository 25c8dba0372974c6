"""CPU tests of the host layer: stream ordering + clique replay + recovery/vertex cover + finder wrapper,
with the device call replaced by the numpy oracle (tests/fake_device.py).  Expected values are the golden
fixtures produced by the unmodified reference."""
import json
import os
import random

import networkx as nx
import numpy as np
import pytest

import fake_device
import nrpcalc_b200
from conftest import GOLDEN_CASES
from nrpcalc_b200 import hgraph, recovery, vercov

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(autouse=True)
def oracle_device(monkeypatch, scratch_cwd):
    monkeypatch.setattr(hgraph, 'repeat_structure', fake_device.oracle_repeat_structure)


def _background(case):
    if case['bg_seqs'] is None:
        return None
    bg = nrpcalc_b200.background(path='./bg_store/', Lmax=case['bg_Lmax'], verbose=False)
    bg.multiadd(case['bg_seqs'])
    assert len(bg) == case['bg_len']
    return bg


@pytest.mark.parametrize('case', GOLDEN_CASES, ids=[c['name'] for c in GOLDEN_CASES])
def test_graph_order_matches_reference(case):
    bg = _background(case)
    graph = hgraph.get_homology_graph(list(case['parts']), bg, case['Lmax'] + 1, case['internal_repeats'], './seq.txt', False)
    assert list(graph.nodes()) == case['graph_nodes']
    assert [list(graph[n]) for n in graph.nodes()] == case['graph_adj']
    lines = open('./seq.txt').read().splitlines()
    assert lines == ['{},{}'.format(i, case['parts'][i].upper()) for i in case['seq_file_ids']]
    if bg is not None:
        bg.drop()


@pytest.mark.parametrize('case', GOLDEN_CASES, ids=[c['name'] for c in GOLDEN_CASES])
def test_finder_toolbox_matches_reference(case):
    bg = _background(case)
    for vc, by_seed in case['toolbox'].items():
        for seed, expected_ids in by_seed.items():
            random.seed(int(seed))
            parts = list(case['parts'])
            found = nrpcalc_b200.finder(seq_list=parts, Lmax=case['Lmax'], internal_repeats=case['internal_repeats'],
                                        background=bg, vercov=vc, output_file=None, verbose=False)
            assert parts == case['parts']                       # caller's list untouched
            assert list(found.keys()) == expected_ids, (vc, seed)
            assert all(found[i] == case['parts'][i].upper() for i in expected_ids)
    if bg is not None:
        bg.drop()


def test_verbose_run_gives_same_toolbox(capsys):
    case = next(c for c in GOLDEN_CASES if c['name'] == 'planted_s0_L15_ir0')
    random.seed(7)
    found = nrpcalc_b200.finder(case['parts'], case['Lmax'], vercov='nrp2', verbose=True)
    assert list(found) == case['toolbox']['nrp2']['7']
    out = capsys.readouterr().out
    assert 'Finder Mode' in out and 'Non-Repetitive Toolbox Size' in out


def test_fasta_output(tmp_path):
    case = next(c for c in GOLDEN_CASES if c['name'] == 'docstring_finder_example')
    bg = _background(case)
    out = str(tmp_path / 'toolbox.fa')
    found = nrpcalc_b200.finder(case['parts'], 15, background=bg, output_file=out, verbose=False)
    assert found == {1: 'GCAGATAGGGGGTAGTA', 0: 'AGAGCTATGACTGACGT'}          # reference docstring, nrpcalc.py:376
    assert list(found) == [1, 0]
    assert open(out).read() == '>index 1\nGCAGATAGGGGGTAGTA\n>index 0\nAGAGCTATGACTGACGT\n'
    bg.drop()


def test_argument_errors():
    with pytest.raises(RuntimeError, match='Invalid Constraints, or Background'):
        nrpcalc_b200.finder(['ACGT', 5], 12, verbose=False)
    with pytest.raises(RuntimeError):
        nrpcalc_b200.finder(['ACGTACGTACGTACGT'], 4, verbose=False)
    with pytest.raises(RuntimeError):
        nrpcalc_b200.finder(['ACGTACGTACGTACGT'], 12, vercov='best', verbose=False)
    with pytest.raises(RuntimeError):
        nrpcalc_b200.finder(['ACGTACGTACGTACGT'], 12, background='not a background', verbose=False)
    with pytest.raises(RuntimeError):
        nrpcalc_b200.finder(['ACGTACGTACGTACGT'], 12, internal_repeats='yes', verbose=False)
    assert nrpcalc_b200.finder([], 12, verbose=False) == {}


def test_closed_background_raises_like_reference():
    bg = nrpcalc_b200.background(path='./closed_bg/', Lmax=12, verbose=False)
    bg.add('ACGTTGCATGCATGCAAAGT')
    bg.close()
    with pytest.raises(RuntimeError, match='closed or dropped'):
        nrpcalc_b200.finder(['ACGTTGCATGCATGCAAAGTCC'], 12, background=bg, verbose=False)


def test_vertex_cover_routines_cover_every_edge():
    rnd = random.Random(5)
    graph = nx.gnm_random_graph(120, 260, seed=3)
    for name in ('nrp2', 'nrpG', '2apx'):
        random.seed(11)
        routine, _ = vercov.ROUTINES[name]
        cover = routine(nx.Graph(graph), False)
        assert all(u in cover or v in cover for u, v in graph.edges())
    del rnd


def test_recovery_returns_maximal_independent_set():
    graph = nx.gnm_random_graph(150, 300, seed=9)
    for name in ('nrp2', 'nrpG', '2apx'):
        random.seed(3)
        chosen = recovery.get_recovered_non_homologs(nx.Graph(graph), './graph.txt', name, False)
        assert all(not (u in chosen and v in chosen) for u, v in graph.edges())
        for node in graph.nodes():
            if node not in chosen:
                assert any(nb in chosen for nb in graph[node])       # cannot be extended


@pytest.mark.parametrize('n_parts,k,internal', [(1500, 9, True), (2500, 10, False), (800, 7, True)])
def test_direct_graph_build_equals_public_api_and_port(n_parts, k, internal):
    """build_homology_graph fills the graph's dicts directly; node order, adjacency order and the graph object must equal
    what networkx's public API gives (hgraph.py:157-200 literally) and what the literal port of the reference builds,
    on inputs dense enough that most cliques overlap."""
    import cases
    from collections import deque
    from oracle import ref_port
    parts = cases.random_parts(n_parts, 60, 100 + n_parts) + cases.planted_parts(200, n_parts)
    uniq = hgraph.uniquify_seq_list(list(parts))
    structure, ids = fake_device.oracle_repeat_structure(uniq, None, k, internal)
    cliques = [tuple(c) for c in hgraph.replay_cliques(hgraph.ordered_tuples(structure, ids))]
    fast = hgraph.build_homology_graph((deque(c) for c in cliques), False)
    slow = hgraph._build_homology_graph_api((deque(c) for c in cliques), False)
    expect, _ = ref_port.homology_graph(list(parts), None, k, internal)
    assert fast.number_of_edges() > n_parts                      # dense: cliques overlap
    for other in (slow, expect):
        assert list(fast.nodes()) == list(other.nodes())
        assert [list(fast[u]) for u in fast.nodes()] == [list(other[u]) for u in other.nodes()]
        assert nx.utils.graphs_equal(fast, other)
    assert fast.number_of_edges() == expect.number_of_edges() and fast.degree(next(iter(fast))) == expect.degree(next(iter(expect)))
    copy = nx.Graph(fast.subgraph(list(fast.nodes())[::2]))      # the graph behaves like any nx.Graph afterwards
    assert copy.number_of_nodes() == (fast.number_of_nodes() + 1) // 2


@pytest.mark.needs_reference
@pytest.mark.parametrize('mode', ['nrp2', 'nrpG', '2apx'])
def test_recovery_against_live_reference(mode, scratch_cwd):
    """Independent-set recovery (reload parsed once, graph filled through its dicts) against the UNMODIFIED reference's
    recovery on dense graphs that need several expansion rounds, same random.seed."""
    import sys
    sys.path.insert(0, os.path.join(HERE, 'golden'))
    import refshim
    if not refshim.reference_available():
        pytest.skip('reference tree not present (GPU box)')
    refshim.load_reference()
    import cases
    from nrpcalc.base import hgraph as ref_hgraph, recovery as ref_recovery
    rounds = []
    real_subgraph = nx.Graph.subgraph

    def counting_subgraph(self, nodes):
        rounds.append(1)
        return real_subgraph(self, nodes)

    for seed, n_parts, k in ((3, 1200, 9), (4, 2000, 10)):
        parts = cases.random_parts(n_parts, 60, 40 + seed) + cases.planted_parts(150, seed)
        theirs_graph = ref_hgraph.get_homology_graph(list(parts), None, k, True, './seq_ref.txt', False)
        mine_graph = hgraph.get_homology_graph(list(parts), None, k, True, './seq_mine.txt', False)
        assert [list(theirs_graph[n]) for n in theirs_graph.nodes()] == [list(mine_graph[n]) for n in mine_graph.nodes()]
        nx.Graph.subgraph = counting_subgraph                   # rounds past the first one, counted on the reference's side
        try:
            random.seed(11 + seed)
            theirs = ref_recovery.get_recovered_non_homologs(theirs_graph, './graph_ref.txt', mode, False)
        finally:
            nx.Graph.subgraph = real_subgraph
        state_theirs = random.getstate()
        random.seed(11 + seed)
        mine = recovery.get_recovered_non_homologs(mine_graph, './graph_mine.txt', mode, False)      # on arrays for nrp2 / nrpG
        assert mine == theirs and len(mine) > 0
        assert random.getstate() == state_theirs                # the generator was consumed exactly as by the reference
        body = lambda path: [line for line in open(path).read().splitlines() if not line.startswith('#')]   # header = time stamp
        assert body('./graph_ref.txt') == body('./graph_mine.txt')
    assert len(rounds) >= 2                                      # at least one case went past the first reload


def test_c_host_tail_equals_python_loops():
    """csrc/host_tail.c (the tail's container operations through the C API) against the Python loops: same cliques in the
    same order, same graph object -- on a dense case, with duplicate ids inside tuples (internal_repeats=True)."""
    import cases
    from collections import deque
    if hgraph._host_tail is None:
        pytest.skip('nrpcalc_b200/_host_tail.so not built')
    parts = cases.random_parts(3000, 60, 77) + cases.planted_parts(300, 78)
    uniq = hgraph.uniquify_seq_list(list(parts))
    structure, ids = fake_device.oracle_repeat_structure(uniq, None, 10, True)
    tuples = list(hgraph.ordered_tuples(structure, ids))
    native = [tuple(c) for c in hgraph.replay_cliques(iter(tuples))]
    python = [tuple(c) for c in hgraph._replay_cliques_python(iter(tuples))]
    assert native == python and len(native) > 1000
    g_native = hgraph.build_homology_graph((deque(c) for c in native), False)
    saved = hgraph._host_tail
    hgraph._host_tail = None
    try:
        g_python = hgraph.build_homology_graph((deque(c) for c in python), False)
    finally:
        hgraph._host_tail = saved
    assert list(g_native.nodes()) == list(g_python.nodes())
    assert [list(g_native[u]) for u in g_native.nodes()] == [list(g_python[u]) for u in g_python.nodes()]
    assert nx.utils.graphs_equal(g_native, g_python)


@pytest.mark.parametrize('n_parts,k,internal', [(3000, 10, True), (2000, 9, False), (6000, 12, False)])
def test_native_array_tail_equals_python_loops(n_parts, k, internal):
    """csrc/host_replay.cpp (the tail on compact arrays, CPython's set layout modelled) against the Python loops on real
    sets: the same clique stream in the same order and the same graph, on dense inputs incl. duplicate ids inside tuples."""
    import cases
    from nrpcalc_b200 import native_tail
    if not native_tail.usable():
        pytest.skip('libnrptail.so not built: ' + native_tail.status()['why'])
    parts = cases.random_parts(n_parts, 60, 300 + n_parts) + cases.planted_parts(300, n_parts)
    uniq = hgraph.uniquify_seq_list(list(parts))
    structure, ids = fake_device.oracle_repeat_structure(uniq, None, k, internal)
    tuples = list(hgraph.ordered_tuples(structure, ids))
    start, size, members = hgraph.ordered_stream_arrays(structure, ids)
    assert [tuple(members[s:s + z].tolist()) for s, z in zip(start.tolist(), size.tolist())] == tuples
    nodes, indptr, adj, c_off, c_nodes = native_tail.run(start, size, members, want_cliques=True)
    python_cliques = [tuple(c) for c in hgraph._replay_cliques_python(iter(tuples))]
    off, flat = c_off.tolist(), c_nodes.tolist()
    assert [tuple(flat[off[i]:off[i + 1]]) for i in range(len(off) - 1)] == python_cliques and len(python_cliques) > 1000
    from collections import deque
    expect = hgraph._build_homology_graph_python((deque(c) for c in python_cliques), False)
    graph = native_tail.graph_from_arrays(nodes, indptr, adj)
    assert list(graph.nodes()) == list(expect.nodes())
    assert [list(graph[u]) for u in graph.nodes()] == [list(expect[u]) for u in expect.nodes()]
    assert nx.utils.graphs_equal(graph, expect)
    some = list(graph.nodes())[:50]
    assert all(graph[u][v] is graph[v][u] for u in some for v in graph[u])       # one data dict per edge, as add_edge makes
    assembled = hgraph.assemble_graph(structure, ids, False)
    assert list(assembled.nodes()) == list(expect.nodes()) and nx.utils.graphs_equal(assembled, expect)


def test_reload_on_arrays_equals_networkx_round_trip(scratch_cwd):
    """native_tail.reload_arrays against nx.read_adjlist(nx.write_adjlist(G)) on random graphs (isolated nodes, shuffled
    node order, empty graph)."""
    import numpy as np
    from nrpcalc_b200 import native_tail
    if not native_tail.usable():
        pytest.skip('libnrptail.so not built')
    rng = random.Random(5)
    for trial in range(25):
        n = rng.choice([0, 1, 5, 60, 800])
        graph = nx.Graph()
        order = list(range(n))
        rng.shuffle(order)
        for u in order:
            if rng.random() < 0.7:
                graph.add_node(u * 3)
        for _ in range(rng.randrange(0, 3 * n + 1)):
            a, b = rng.randrange(max(n, 1)) * 3, rng.randrange(max(n, 1)) * 3
            if a != b:
                graph.add_edge(a, b)
        nodes = np.array(list(graph.nodes()), dtype=np.int64)
        indptr, flat = [0], []
        for u in graph.nodes():
            flat.extend(graph[u])
            indptr.append(len(flat))
        nx.write_adjlist(graph, './g.txt')
        want = nx.read_adjlist('./g.txt', nodetype=int)
        got = native_tail.graph_from_arrays(*native_tail.reload_arrays(nodes, np.array(indptr, dtype=np.int64), np.array(flat, dtype=np.int64)))
        assert list(want.nodes()) == list(got.nodes())
        assert [list(want[u]) for u in want.nodes()] == [list(got[u]) for u in got.nodes()]
        assert nx.utils.graphs_equal(want, got)


def test_native_array_tail_on_many_small_streams():
    """400 small random streams (tuples of 1-7 ids from a tiny universe: heavy overlap, equal tuples in different orders,
    repeated ids inside a tuple, subset chains) through the array tail and through the Python loops."""
    from nrpcalc_b200 import native_tail
    if not native_tail.usable():
        pytest.skip('libnrptail.so not built')
    rng = random.Random(99)
    for trial in range(400):
        universe = rng.choice([3, 8, 20, 60, 400])
        tuples = []
        for _ in range(rng.randrange(1, 120)):
            k = rng.choice([1, 1, 1, 2, 2, 3, 4, 7])
            base = rng.randrange(universe)
            tuples.append(tuple((base + rng.randrange(0, 4)) % universe if rng.random() < 0.6 else rng.randrange(universe)
                                for _ in range(k)))
            if rng.random() < 0.2 and len(tuples[-1]) > 1:                    # a sub-tuple and a permutation of it
                tuples.append(tuples[-1][:-1])
                tuples.append(tuples[-1][::-1])
        nodes, indptr, adj, c_off, c_nodes = native_tail.run(*native_tail.stream_arrays(tuples), want_cliques=True)
        cliques = [tuple(c) for c in hgraph._replay_cliques_python(iter(tuples))]
        off, flat = c_off.tolist(), c_nodes.tolist()
        assert [tuple(flat[off[i]:off[i + 1]]) for i in range(len(off) - 1)] == cliques, trial
        from collections import deque
        expect = hgraph._build_homology_graph_python((deque(c) for c in cliques), False)
        assert nodes.tolist() == list(expect.nodes()), trial
        bounds, nbrs = indptr.tolist(), adj.tolist()
        assert [nbrs[bounds[i]:bounds[i + 1]] for i in range(len(bounds) - 1)] == [list(expect[u]) for u in expect.nodes()], trial


def test_patch_runs_the_reference_finder_on_our_graph_build(scratch_cwd):
    """nrpcalc_b200.patch(): the UNMODIFIED reference's finder / recovery / vercov run on a graph built behind the seam
    hgraph.get_homology_graph (reference hgraph.py:202-249); toolboxes equal the golden ones, with and without a (reference)
    background object; unpatch() restores the reference's own function."""
    import refshim
    if not refshim.reference_available():
        pytest.skip('no reference tree (baseline/_ref or /root/reference)')
    ref = refshim.load_reference()
    from nrpcalc.base import hgraph as ref_hgraph
    original = ref_hgraph.get_homology_graph
    assert nrpcalc_b200.patch(ref) is ref and ref_hgraph.get_homology_graph is not original
    try:
        for name in ('planted_s0_L15_ir0', 'docstring_finder_example', 'short_parts_ir1'):
            case = next(c for c in GOLDEN_CASES if c['name'] == name)
            bg = None
            if case['bg_seqs'] is not None:
                bg = ref.background(path='./ref_patch_bg/', Lmax=case['bg_Lmax'], verbose=False)
                bg.multiadd(case['bg_seqs'])
            for vc, by_seed in case['toolbox'].items():
                for seed, expected_ids in by_seed.items():
                    random.seed(int(seed))
                    found = ref.finder(seq_list=list(case['parts']), Lmax=case['Lmax'], internal_repeats=case['internal_repeats'],
                                       background=bg, vercov=vc, verbose=False)
                    assert list(found.keys()) == expected_ids, (name, vc, seed)
            if bg is not None:
                bg.drop()
    finally:
        nrpcalc_b200.unpatch()
    assert ref_hgraph.get_homology_graph is original


def test_native_recovery_equals_graph_based_recovery(tmp_path):
    """libnrptail's recovery on arrays (cover routines with the interpreter's Mersenne Twister state, candidate sets and working
    graphs of every round modelled slot for slot) against the graph-based loop on random graphs of several densities: same
    independent set, same random state afterwards; some cases take several rounds, with the working graph ordered by the candidate
    set (2 * |candidates| < |graph|) and by the graph."""
    from nrpcalc_b200 import native_tail
    if not (native_tail.usable() and native_tail.cover_usable()):
        pytest.skip('libnrptail.so not built')
    rng = random.Random(5)
    rounds_seen, shapes = [], set()
    for case in range(80):
        n = rng.choice((40, 300, 2000, 6000))
        m = int(n * rng.choice((0.3, 0.6, 1.0, 1.5, 3.0, 6.0, 12.0)))
        mode = rng.choice(('nrp2', 'nrpG'))
        graph = nx.Graph()
        ids = rng.sample(range(3 * n), n)
        for u in ids:
            if rng.random() < 0.95:
                graph.add_node(u)
        for _ in range(m):
            u, v = rng.choice(ids), rng.choice(ids)
            if u != v:
                graph.add_edge(u, v)
        for _ in range(n // 50):                                    # a few hubs: long tie lists for the shuffles
            hub = rng.choice(ids)
            for v in rng.sample(ids, min(12, n)):
                if v != hub:
                    graph.add_edge(hub, v)
        nodes = np.fromiter(graph.nodes(), dtype=np.int64)
        indptr = np.zeros(nodes.size + 1, dtype=np.int64)
        np.cumsum([len(graph[u]) for u in graph.nodes()], out=indptr[1:])
        adj = np.fromiter((v for u in graph.nodes() for v in graph[u]), dtype=np.int64, count=int(indptr[-1]))
        random.seed(case)
        got, rounds, _ = native_tail.recover_arrays(nodes, indptr, adj, mode)
        state_native = random.getstate()
        random.seed(case)
        expect = recovery._recover_on_graphs(graph, str(tmp_path / 'g.txt'), mode, False)
        assert got == expect, (case, n, m, mode)
        assert random.getstate() == state_native, (case, n, m, mode)
        rounds_seen.append(rounds)
        shapes.add((n, mode))
    assert sum(1 for r in rounds_seen if r >= 2) >= 10 and len(shapes) >= 6, rounds_seen      # rounds >= 2: a re-derived working graph


def test_adjlist_file_from_arrays_equals_networkx(tmp_path):
    from nrpcalc_b200 import native_tail
    if not native_tail.usable():
        pytest.skip('libnrptail.so not built')
    rng = random.Random(9)
    graph = nx.Graph()
    for u in rng.sample(range(5000), 1500):
        graph.add_node(u)
    ids = list(graph.nodes())
    for _ in range(2600):
        u, v = rng.choice(ids), rng.choice(ids)
        if u != v:
            graph.add_edge(u, v)
    nodes = np.fromiter(graph.nodes(), dtype=np.int64)
    indptr = np.zeros(nodes.size + 1, dtype=np.int64)
    np.cumsum([len(graph[u]) for u in graph.nodes()], out=indptr[1:])
    adj = np.fromiter((v for u in graph.nodes() for v in graph[u]), dtype=np.int64, count=int(indptr[-1]))
    nx.write_adjlist(graph, str(tmp_path / 'nx.txt'))
    recovery._write_adjlist_from_arrays((nodes, indptr, adj), str(tmp_path / 'mine.txt'))
    body = lambda path: [line for line in open(path).read().splitlines() if not line.startswith('#')]
    assert body(str(tmp_path / 'mine.txt')) == body(str(tmp_path / 'nx.txt'))
    assert len(open(str(tmp_path / 'mine.txt')).read().splitlines()) == len(open(str(tmp_path / 'nx.txt')).read().splitlines())
    reread = nx.read_adjlist(str(tmp_path / 'mine.txt'), nodetype=int)
    assert sorted(reread.edges()) == sorted(tuple(sorted(e)) for e in graph.edges()) or set(map(frozenset, reread.edges())) == set(map(frozenset, graph.edges()))


def test_self_check_verdicts_are_remembered_by_library_content(tmp_path, monkeypatch):
    """The first-use self-checks of libnrptail.so (array tail, cover routines) cost more than a 100k-part job; a passed check is
    remembered next to the library under a key of interpreter, networkx and the library's CONTENT -- a copied tree keeps it, another
    library (or a stale stamp) is checked again."""
    import shutil
    from nrpcalc_b200 import native_tail
    if not native_tail.usable():
        pytest.skip('libnrptail.so not built')
    lib_copy = tmp_path / 'libnrptail.so'
    shutil.copy(native_tail.LIB_PATH, lib_copy)
    os.utime(lib_copy, (1, 1))                                   # another time stamp: the key must not care
    key_here = native_tail._verdict_key()
    monkeypatch.setattr(native_tail, 'LIB_PATH', str(lib_copy))
    assert native_tail._verdict_key() == key_here
    calls = {'tail': 0, 'cover': 0}
    real_tail, real_cover = native_tail._self_check, native_tail._cover_self_check

    def counted_tail():
        calls['tail'] += 1
        return real_tail()

    def counted_cover():
        calls['cover'] += 1
        return real_cover()

    monkeypatch.setattr(native_tail, '_self_check', counted_tail)
    monkeypatch.setattr(native_tail, '_cover_self_check', counted_cover)

    def fresh_process():
        monkeypatch.setattr(native_tail, '_state', {'checked': False, 'usable': False, 'why': ''})
        monkeypatch.setattr(native_tail, '_cover_state', {'checked': False, 'usable': False, 'why': ''})

    stamp = str(lib_copy) + '.ok'
    fresh_process()                                              # no stamp: both checks run, both verdicts are written
    assert native_tail.usable() and native_tail.cover_usable() and calls == {'tail': 1, 'cover': 1}
    assert open(stamp).read().split('\n') == [key_here, 'cover ok']
    fresh_process()                                              # stamp present: nothing runs
    assert native_tail.usable() and native_tail.cover_usable() and calls == {'tail': 1, 'cover': 1}
    with open(stamp, 'w') as fh:                                 # a round-1 stamp (tail only): the cover check runs once more
        fh.write(key_here)
    fresh_process()
    assert native_tail.usable() and native_tail.cover_usable() and calls == {'tail': 1, 'cover': 2}
    with open(stamp, 'w') as fh:                                 # a stamp of another library: everything is checked again
        fh.write('something else\ncover ok')
    fresh_process()
    assert native_tail.usable() and native_tail.cover_usable() and calls == {'tail': 2, 'cover': 3}
