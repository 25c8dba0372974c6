"""CPU tests of the general path's HOST side (any alphabet, any k): routing between nrp_build and nrp_build_generic, host-decided
background flags, length classes, the order-exact tail -- with the device context replaced by a host model (fake_device.ModelContext)
that refuses what the 2-bit kernels refuse.  Expected values: fixtures produced by the unmodified reference
(tests/golden/make_golden_general.py)."""
import random

import numpy as np
import pytest

import fake_device
import nrpcalc_b200
from conftest import GOLDEN_CASES, GOLDEN_GENERAL
from nrpcalc_b200 import hgraph


@pytest.fixture
def model(monkeypatch, scratch_cwd):
    ctx = fake_device.ModelContext()
    monkeypatch.setattr(hgraph, '_context', lambda device=None: ctx)
    return ctx


def _background(case):
    if case['bg_seqs'] is None:
        return None
    bg = nrpcalc_b200.background(path='./bg_store/', Lmax=case['bg_Lmax'], verbose=False)
    bg.multiadd(case['bg_seqs'])
    assert len(bg) == case['bg_len']
    return bg


@pytest.mark.parametrize('case', GOLDEN_GENERAL, ids=[c['name'] for c in GOLDEN_GENERAL])
def test_general_graph_and_toolbox(case, model):
    bg = _background(case)
    graph = hgraph.get_homology_graph(list(case['parts']), bg, case['Lmax'] + 1, case['internal_repeats'], './seq.txt', False)
    assert list(graph.nodes()) == case['graph_nodes']
    assert [list(graph[n]) for n in graph.nodes()] == case['graph_adj']
    assert any(name == 'build_generic' for name, _ in model.calls)
    lines = open('./seq.txt').read().splitlines()
    assert lines == ['{},{}'.format(i, case['parts'][i].upper()) for i in case['seq_file_ids']]
    for vc, by_seed in case['toolbox'].items():
        for seed, expected_ids in by_seed.items():
            random.seed(int(seed))
            found = nrpcalc_b200.finder(seq_list=list(case['parts']), Lmax=case['Lmax'], internal_repeats=case['internal_repeats'],
                                        background=bg, vercov=vc, output_file=None, verbose=False)
            assert list(found.keys()) == expected_ids, (vc, seed)
            assert all(found[i] == case['parts'][i].upper() for i in expected_ids)
    if bg is not None:
        bg.drop()


@pytest.mark.parametrize('case', GOLDEN_CASES[:12], ids=[c['name'] for c in GOLDEN_CASES[:12]])
def test_plain_inputs_stay_on_the_2bit_build(case, model):
    """A/C/G/T parts with k <= 32 never take the general path (the model's build() would refuse anything else)."""
    bg = None
    if case['bg_seqs'] is not None:
        bg = nrpcalc_b200.background(path='./bg_store/', Lmax=case['bg_Lmax'], verbose=False)
        bg.multiadd(case['bg_seqs'])
    graph = hgraph.get_homology_graph(list(case['parts']), bg, case['Lmax'] + 1, case['internal_repeats'], './seq.txt', False)
    assert list(graph.nodes()) == case['graph_nodes']
    assert [list(graph[n]) for n in graph.nodes()] == case['graph_adj']
    assert all(name == 'build' for name, _ in model.calls)
    if bg is not None:
        bg.drop()


def test_host_background_flags_cover_only_what_the_device_cannot_decide(model):
    parts = ['ACGTACGTACGTACGTACGTAAGG', 'ACGTNCGTACGTACGTACGTAAGG', 'TTTTTTTTTTTTTTTTTTTTTTTT', 'GGGGNNGGGGGGGGGGGGGGGGGG']
    bg = nrpcalc_b200.background(path='./bg_a/', Lmax=7, verbose=False)
    bg.multiadd(['GGGGNNGG', 'CGTACGTA'])
    uniq = [(i, s) for i, s in enumerate(parts)]
    from nrpcalc_b200 import encoding
    buf, off = encoding.join_parts(parts)
    flags = hgraph._host_background_flags(bg, parts, buf, off)
    assert flags.tolist() == [0, 4, 0, 4]          # part 0 is plain (the device probes it); part 1 holds CGTACGTA, part 3 the N key
    structure, _ = hgraph.repeat_structure(uniq, bg, 9, True)
    assert (structure.flags != 0).tolist() == [True, True, False, True]
    bg.drop()


def test_join_parts_and_alphabet_check():
    """encoding.join_parts / all_acgt (the routing decision between the 2-bit kernels and the general path): the answer by
    bytes.translate on the joined bytes equals the table look-up on a copy, for both cases, other symbols, empty parts, no parts;
    text outside ASCII is refused."""
    from nrpcalc_b200 import encoding
    for seqs, want in ((['ACGT', 'acgt', 'AAAA'], True), (['ACGT', 'ACNT'], False), (['ACGU'], False), (['AC-T'], False),
                       ([], True), ([''], True), (['', 'ACGT', ''], True), (['ACGT' * 1000, 'T' * 33], True), (['ACGT' * 1000 + 'R'], False)):
        buf, off = encoding.join_parts(seqs)
        assert off.tolist() == np.concatenate(([0], np.cumsum([len(s) for s in seqs]))).astype(int).tolist()
        assert bytes(buf) == ''.join(seqs).encode()
        assert encoding.all_acgt(buf) is want                        # a view of the joined bytes: translate
        assert encoding.all_acgt(buf.copy()) is want                 # any other array: the table
        assert encoding.all_acgt(buf[:max(0, buf.size - 1)]) in (True, False)
    with pytest.raises(ValueError):
        encoding.join_parts(['ACGÄ'])
