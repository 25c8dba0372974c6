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


def test_other_alphabets_take_the_general_path(scratch_cwd):
    """RNA and other symbols are keyed as the reference keys them (utils.py:10: U's complement is A, an N is its own); text that
    is not ASCII has no byte form and is refused."""
    parts = ['ACGUACGUACGUACGUAA', 'ACGTNACGTACGTACGTACGT', 'UUACGUACGUACGUACGU', 'ACGTNACGTACGTACGAAAAA']
    random.seed(1)
    found = nrpcalc_b200.finder(list(parts), 12, verbose=False)
    expect, _ = ref_port.homology_graph(list(parts), None, 13, False)
    assert set(found) <= set(expect.nodes()) and hgraph.last_build_info['counts']['generic_builds'] == 1
    graph = hgraph.get_homology_graph(list(parts), None, 13, False, './seq.txt', False)
    assert list(graph.nodes()) == list(expect.nodes()) and [list(graph[n]) for n in graph.nodes()] == [list(expect[n]) for n in expect.nodes()]
    with pytest.raises(ValueError, match='ASCII'):
        nrpcalc_b200.finder(['ACGTACGTACGTACGT\u00e9ACGT', 'ACGTACGTACGTACGTTTTT'], 12, verbose=False)


def test_background_built_on_device_matches_host(scratch_cwd):
    """nrp_bg_build (extract + sort + unique on the GPU) against the storage layer's host code and the reference's
    known answer (nrpcalc.py:89-167: six 17-mers at Lmax=15 -> 12 canonical 16-mers, 10 after a remove)."""
    import numpy as np
    from nrpcalc_b200 import encoding
    docs = ['ATGAGATCGTAGCAACC', 'GACGATTACGTCAGGTA', 'ACAGTAGAGACGAGTAA', 'CCAGTACGAAAAGGCCC', 'TTAGCTTGATAGTTTTA']
    listed = ['AAAACTATCAAGCTAA', 'ACAGTAGAGACGAGTA', 'ACCTGACGTAATCGTC', 'ACGATTACGTCAGGTA', 'ATGAGATCGTAGCAAC', 'ATGCTTAGTGCCATAC',
              'CAGTACGAAAAGGCCC', 'CAGTAGAGACGAGTAA', 'CCAGTACGAAAAGGCC', 'GGTATGGCACTAAGCA', 'GGTTGCTACGATCTCA', 'TAAAACTATCAAGCTA']
    host = nrpcalc_b200.background(path='./bg_host/', Lmax=15, verbose=False)
    dev = nrpcalc_b200.background(path='./bg_dev/', Lmax=15, verbose=False, device=True)
    for bg in (host, dev):
        bg.add('ATGCTTAGTGCCATACC')
        bg.multiadd(docs)
    assert len(dev) == len(host) == 12 and list(dev) == list(host) == listed
    dev.remove('ATGCTTAGTGCCATACC'); host.remove('ATGCTTAGTGCCATACC')
    assert len(dev) == len(host) == 10 and list(dev) == list(host)
    assert 'ATGCTTAGTGCCATACC' not in dev and all(dev.multicheck(docs))
    # a genome-sized batch with lower case, duplicates, short and K-long sequences
    genome = cases.random_genome(300_000, 5)
    seqs = [genome, genome[1000:5000].lower(), 'ACGT', genome[:16], genome[7:23]] + cases.planted_parts(300, 3, p_short=0.2)
    seqs = [s for s in seqs if len(s) >= 16]
    ctx = hgraph._context()
    buf, off = encoding.join_parts(seqs)
    got = ctx.bg_build(buf, off, 16)
    want = np.unique(encoding.canonical_kmers(buf, off, 16))
    assert got.dtype == np.uint64 and got.tolist() == want.tolist()
    got25 = ctx.bg_build(buf, off, 25)
    assert got25.tolist() == np.unique(encoding.canonical_kmers(buf, off, 25)).tolist()
    for bg in (host, dev):
        bg.drop()


def test_finder_with_device_built_background(scratch_cwd):
    case = next(c for c in GOLDEN_CASES if c['bg_seqs'] is not None)
    bg = nrpcalc_b200.background(path='./bg_store_dev/', Lmax=case['bg_Lmax'], verbose=False, device=True)
    bg.multiadd(case['bg_seqs'])
    graph = hgraph.get_homology_graph(list(case['parts']), bg, case['Lmax'] + 1, case['internal_repeats'], './seq.txt', False)
    assert list(graph.nodes()) == case['graph_nodes']
    assert [list(graph[n]) for n in graph.nodes()] == case['graph_adj']
    bg.drop()


@pytest.mark.parametrize('device_bg', [False, True])
def test_background_with_other_characters(scratch_cwd, device_bg):
    """A background chromosome with 'N's and a soft-masked (lower-case) stretch: only the windows that cover such a character
    are lost to the device probe (kmerSetDB classifies windows one by one); result equals the reference port with its string
    keys (hgraph.py:72-79, kmerSetDB.py:177-207)."""
    parts = cases.planted_parts(800, 321)
    rnd = random.Random(5)
    pieces = []
    for seq in parts:
        if len(seq) >= 40 and rnd.random() < 0.3:
            pieces.append(seq[5:35].upper())
    chromosome = 'N'.join(pieces[:len(pieces) // 2]) + 'NNNNN' + ''.join(p if i % 3 else p.lower() for i, p in enumerate(pieces[len(pieces) // 2:]))
    for K, k in ((14, 14), (12, 16), (20, 15)):
        bg = nrpcalc_b200.background(path='./bg_n_%d_%d/' % (K, int(device_bg)), Lmax=K - 1, verbose=False, device=device_bg)
        bg.multiadd([chromosome, cases.random_genome(5000, 3)])
        model = ref_port.SetBackground(K)
        model.multiadd([chromosome, cases.random_genome(5000, 3)])
        assert len(bg) == len(model)
        graph = hgraph.get_homology_graph(list(parts), bg, k, False, './seq.txt', False)
        expect, _ = ref_port.homology_graph(list(parts), model, k, False)
        assert list(graph.nodes()) == list(expect.nodes()) and len(expect.nodes()) < len(set(p.upper() for p in parts))
        assert [list(graph[n]) for n in graph.nodes()] == [list(expect[n]) for n in expect.nodes()]
        bg.drop()


def test_side_files_and_fasta_on_the_device_path(scratch_cwd):
    """Byte-exact artefacts through the CUDA path: seq_list.txt in processing order (hgraph.py:218-223), the adjacency list the
    recovery dumps (recovery.py:26-30) and the FASTA writer (finder.py:184-189)."""
    import networkx as nx
    case = next(c for c in GOLDEN_CASES if c['name'] == 'planted_s0_L15_ir0')
    graph = hgraph.get_homology_graph(list(case['parts']), None, case['Lmax'] + 1, case['internal_repeats'], './seq.txt', False)
    assert open('./seq.txt').read().splitlines() == ['{},{}'.format(i, case['parts'][i].upper()) for i in case['seq_file_ids']]
    expect, _ = ref_port.homology_graph(list(case['parts']), None, case['Lmax'] + 1, case['internal_repeats'])
    nx.write_adjlist(expect, './expect_graph.txt')
    from nrpcalc_b200 import recovery
    random.seed(7)
    recovery.get_recovered_non_homologs(graph, './graph.txt', 'nrp2', False)
    body = lambda path: [line for line in open(path).read().splitlines() if not line.startswith('#')]
    assert body('./graph.txt') == body('./expect_graph.txt')
    random.seed(7)
    found = nrpcalc_b200.finder(case['parts'], case['Lmax'], vercov='nrp2', output_file='./toolbox.fa', verbose=False)
    import textwrap
    want = ''.join('>index {}\n{}\n'.format(i, '\n'.join(textwrap.wrap(case['parts'][i].upper(), 80))) for i in case['toolbox']['nrp2']['7'])
    assert list(found) == case['toolbox']['nrp2']['7'] and open('./toolbox.fa').read() == want


def test_patch_on_the_device_path(scratch_cwd):
    """nrpcalc_b200.patch(): the unmodified reference's finder runs on a graph built by libnrpb200.so behind its own seam."""
    import refshim
    if not refshim.reference_available():
        pytest.skip('no reference tree (baseline/_ref or /root/reference)')
    ref = refshim.load_reference()
    nrpcalc_b200.patch(ref)
    try:
        for name in ('planted_s0_L15_ir0', 'docstring_finder_example'):
            case = next(c for c in GOLDEN_CASES if c['name'] == name)
            bg = None
            if case['bg_seqs'] is not None:
                bg = ref.background(path='./ref_bg_gpu/', Lmax=case['bg_Lmax'], verbose=False)
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


def _py_uniquify(parts):
    """hgraph.py:11-21 + the processing order of 218-223 on strings: (ids, sources) by descending last occurrence"""
    first, last = {}, {}
    for i, s in enumerate(parts):
        s = s.upper()
        first.setdefault(s, i)
        last[s] = i
    order = sorted(last.values(), reverse=True)
    return [first[parts[i].upper()] for i in order], order


@pytest.mark.parametrize('uniform', [True, False])
def test_device_uniquify_against_python(uniform):
    """nrp_uniquify: ids, sources, offsets and the gathered (upper-cased) bytes against the dict-based host version on lists with
    many duplicates, lower-case copies, parts that differ in their last byte only and empty parts."""
    import ctypes
    import numpy as np
    rnd = random.Random(11)
    ctx = hgraph._context()
    for n, dup in ((1, 0.0), (5000, 0.0), (20000, 0.4), (3000, 0.95)):
        base = cases.random_parts(n, 60, n) if uniform else cases.buffer_to_list(*cases.random_varlen_buffer(n, 0 if n > 1 else 5, 90, n))
        parts = []
        for i in range(n):
            r = rnd.random()
            if parts and r < dup:
                s = parts[rnd.randrange(len(parts))]
                s = s.lower() if rnd.random() < 0.3 else s
            elif parts and r < dup + 0.02 and len(parts[-1]) > 1:
                s = parts[-1][:-1] + ('A' if parts[-1][-1].upper() != 'A' else 'C')      # equal but for the last byte
            else:
                s = base[i]
            parts.append(s)
        blob = np.frombuffer(''.join(parts).encode(), dtype=np.uint8)
        offsets = np.zeros(n + 1, dtype=np.int64)
        np.cumsum([len(s) for s in parts], out=offsets[1:])
        if uniform:
            ids, source, new_off, dev, status, longest = ctx.uniquify(blob, None, 60)
        else:
            ids, source, new_off, dev, status, longest = ctx.uniquify(blob, offsets, 0)
        want_ids, want_src = _py_uniquify(parts)
        assert status == 0 and longest == max(len(s) for s in parts)
        assert ids.tolist() == want_ids and source.tolist() == want_src
        assert np.diff(new_off).tolist() == [len(parts[i]) for i in want_src]
        import torch
        gathered = torch.empty(int(new_off[-1]), dtype=torch.uint8, device='cuda:{}'.format(ctx.device))
        if gathered.numel():
            ctypes.CDLL('libcudart.so.12').cudaMemcpy(ctypes.c_void_p(gathered.data_ptr()), ctypes.c_void_p(dev), ctypes.c_size_t(gathered.numel()), 3)
        assert gathered.cpu().numpy().tobytes().decode() == ''.join(parts[i].upper() for i in want_src)
    ids, source, new_off, dev, status, longest = ctx.uniquify(np.frombuffer(b'ACGTNNACGTAC', dtype=np.uint8), None, 6)
    assert status == 1 and ids.tolist() == [1, 0]


@pytest.mark.parametrize('case', [c for c in GOLDEN_CASES if c['parts']], ids=[c['name'] for c in GOLDEN_CASES if c['parts']])
def test_finder_packed_matches_golden(case, scratch_cwd):
    """The packed entry (device uniquify -> build on the gathered buffer -> array tail) against the reference's toolboxes."""
    import numpy as np
    from nrpcalc_b200 import packed
    bg = _background(case)
    blob = ''.join(case['parts']).encode()
    offsets = np.zeros(len(case['parts']) + 1, dtype=np.int64)
    np.cumsum([len(s) for s in case['parts']], out=offsets[1:])
    short = any(len(s) <= case['Lmax'] for s in case['parts'])
    for vc, by_seed in case['toolbox'].items():
        for seed, expected_ids in by_seed.items():
            random.seed(int(seed))
            found = nrpcalc_b200.finder_packed(blob, case['Lmax'], offsets=offsets, internal_repeats=case['internal_repeats'], background=bg, vercov=vc)
            assert list(found.keys()) == expected_ids, (vc, seed)
            assert all(found[i] == case['parts'][i].upper() for i in expected_ids)
            assert packed.last_run_info['path'] == ('strings' if short else 'device')
    if bg is not None:
        bg.drop()


def test_finder_packed_uniform_batch_equals_finder(scratch_cwd):
    from nrpcalc_b200 import packed
    parts = cases.random_parts(30000, 100, 77) + cases.random_parts(2000, 100, 77)       # 2000 exact duplicates
    parts[5] = parts[5].lower()
    random.seed(3)
    want = nrpcalc_b200.finder(list(parts), 12, verbose=False)
    random.seed(3)
    got = nrpcalc_b200.finder_packed(''.join(parts).encode(), 12, part_len=100)
    assert packed.last_run_info['path'] == 'device' and packed.last_run_info['unique'] == 30000
    assert got == want and list(got) == list(want)
