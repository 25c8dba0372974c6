"""GPU end-to-end parity: nrpcalc_b200.finder() / get_homology_graph() through libnrpb200.so against the golden
fixtures produced by the unmodified reference, and against the literal port on fresh seeded inputs."""
import random

import pytest

import cases
import nrpcalc_b200
from conftest import GOLDEN_CASES
from nrpcalc_b200 import hgraph
from oracle import ref_port

pytestmark = pytest.mark.gpu


def _background(case):
    if case['bg_seqs'] is None:
        return None
    bg = nrpcalc_b200.background(path='./bg_store/', Lmax=case['bg_Lmax'], verbose=False)
    bg.multiadd(case['bg_seqs'])
    return bg


@pytest.mark.parametrize('case', GOLDEN_CASES, ids=[c['name'] for c in GOLDEN_CASES])
def test_golden_graph_and_toolbox(case, scratch_cwd):
    bg = _background(case)
    graph = hgraph.get_homology_graph(list(case['parts']), bg, case['Lmax'] + 1, case['internal_repeats'], './seq.txt', False)
    assert list(graph.nodes()) == case['graph_nodes']
    assert [list(graph[n]) for n in graph.nodes()] == case['graph_adj']
    for vc, by_seed in case['toolbox'].items():
        for seed, expected_ids in by_seed.items():
            random.seed(int(seed))
            found = nrpcalc_b200.finder(seq_list=list(case['parts']), Lmax=case['Lmax'], internal_repeats=case['internal_repeats'],
                                        background=bg, vercov=vc, output_file=None, verbose=False)
            assert list(found.keys()) == expected_ids, (vc, seed)
            assert all(found[i] == case['parts'][i].upper() for i in expected_ids)
    if bg is not None:
        bg.drop()


@pytest.mark.parametrize('seed', range(3))
def test_fresh_inputs_against_reference_port(seed, scratch_cwd):
    parts = cases.planted_parts(1200, 900 + seed, p_short=0.05)
    for k, internal in ((13, False), (16, True), (17, False), (8, False)):
        graph = hgraph.get_homology_graph(list(parts), None, k, internal, './seq.txt', False)
        expect, _ = ref_port.homology_graph(list(parts), None, k, internal)
        assert list(graph.nodes()) == list(expect.nodes())
        assert [list(graph[n]) for n in graph.nodes()] == [list(expect[n]) for n in expect.nodes()]


def test_config1_shape(scratch_cwd):
    """BASELINE config 1: 10k x 100 bp, Lmax=12, no background -- the reference's own CPU-runnable case."""
    parts = cases.random_parts(10000, 100, 1)
    graph = hgraph.get_homology_graph(list(parts), None, 13, False, './seq.txt', False)
    expect, _ = ref_port.homology_graph(list(parts), None, 13, False)
    assert list(graph.nodes()) == list(expect.nodes())
    assert [list(graph[n]) for n in graph.nodes()] == [list(expect[n]) for n in expect.nodes()]
    excluded = 10000 - graph.number_of_nodes()          # internal inverted repeats (hairpins of length k+1 are common)
    assert hgraph.last_build_info['records'] == (10000 - excluded) * 88
    assert hgraph.last_build_info['counts']['runs'] == 1


def test_rejects_other_alphabets(scratch_cwd):
    with pytest.raises(ValueError, match='A, C, G, T'):
        nrpcalc_b200.finder(['ACGUACGUACGUACGUAA', 'ACGTNACGTACGTACGTACGT'], 12, verbose=False)
