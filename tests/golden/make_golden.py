"""Generate tests/golden/*.json by running the UNMODIFIED reference (dev container only).

    cd /root/repo && python tests/golden/make_golden.py

For every case the fixture stores the inputs verbatim and what the reference produced:
  * the repeat graph returned by hgraph.get_homology_graph (reference hgraph.py:202-249): node order and
    every adjacency list in order (both are observable by the unchanged vertex-cover step),
  * the id order of the seq_list.txt side file (hgraph.py:218-223; each line is 'id,UPPER(parts[id])'),
  * nrpcalc.finder(...) output (nrpcalc.py:226-391) for each vertex-cover routine under random.seed(s),
    as the ordered list of ids (the value of each id is UPPER(parts[id])).
Background cases also store size and sha256 of the sorted canonical key set of the kmerSetDB
(kmerSetDB.py:147-175, 209-217); background_golden.json stores the docstring example's keys verbatim.
"""
import hashlib
import json
import os
import random
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import cases            # noqa: E402
import refshim          # noqa: E402


def run_reference(ref, parts, Lmax, internal_repeats, bg_seqs, bg_Lmax, seeds=(7, 123)):
    from nrpcalc.base import hgraph as ref_hgraph
    out = {}
    background = None
    if bg_seqs is not None:
        background = ref.background(path='./bg_{}/'.format(random.randrange(1 << 30)), Lmax=bg_Lmax, verbose=False)
        background.multiadd(bg_seqs)
        out['bg_keys_sha256'] = hashlib.sha256('\n'.join(sorted(background)).encode()).hexdigest()
        out['bg_len'] = len(background)
    graph = ref_hgraph.get_homology_graph(list(parts), background, Lmax + 1, internal_repeats, './seq_list.txt', False)
    out['graph_nodes'] = list(graph.nodes())
    out['graph_adj'] = [list(graph[n]) for n in graph.nodes()]
    seq_lines = open('./seq_list.txt').read().splitlines()
    out['seq_file_ids'] = [int(line.split(',')[0]) for line in seq_lines]      # sequence = upper(parts[id])
    assert all(line.split(',')[1] == parts[int(line.split(',')[0])].upper() for line in seq_lines)
    out['toolbox'] = {}
    for vercov in ('nrp2', 'nrpG', '2apx'):
        out['toolbox'][vercov] = {}
        for seed in seeds:
            random.seed(seed)
            found = ref.finder(seq_list=list(parts), Lmax=Lmax, internal_repeats=internal_repeats,
                               background=background, vercov=vercov, output_file=None, verbose=False)
            assert all(s == parts[i].upper() for i, s in found.items())
            out['toolbox'][vercov][str(seed)] = [int(i) for i in found]          # ordered ids; seq = upper(parts[id])
    if background is not None:
        background.drop()
    return out


def main():
    ref = refshim.load_reference()
    fixtures = []
    partsets = {}

    def add(name, parts, Lmax, internal_repeats, bg_seqs=None, bg_Lmax=None):
        parts = list(parts)
        set_name = hashlib.sha256('\n'.join(parts).encode()).hexdigest()[:12]
        partsets[set_name] = parts
        case = {'name': name, 'partset': set_name, 'Lmax': Lmax, 'internal_repeats': internal_repeats,
                'bg_seqs': bg_seqs, 'bg_Lmax': bg_Lmax}
        case.update(run_reference(ref, parts, Lmax, internal_repeats, bg_seqs, bg_Lmax))
        fixtures.append(case)
        print('{:40s} nodes {:5d} edges {:5d}'.format(name, len(case['graph_nodes']),
                                                     sum(len(a) for a in case['graph_adj']) // 2))

    # known-answer example of the reference's finder docstring (nrpcalc.py:293-378)
    doc_bg = ['ATGAGATCGTAGCAACC', 'GACGATTACGTCAGGTA', 'ACAGTAGAGACGAGTAA', 'CCAGTACGAAAAGGCCC', 'AAAAAAAAAAAAAAAAA']
    doc_parts = ['AGAGCTATGACTGACGT', 'GCAGATAGGGGGTAGTA', 'TAAAAAAAAAAAAAAAA', 'GAGCTATGACTGACGTC']
    add('docstring_finder_example', doc_parts, 15, False, doc_bg, 15)

    for seed in range(3):
        parts = cases.planted_parts(260, seed)
        for Lmax in (12, 15, 16):
            for internal in (False, True):
                add('planted_s{}_L{}_ir{}'.format(seed, Lmax, int(internal)), parts, Lmax, internal)
        bg = cases.background_seqs_for(parts, seed)
        for Lmax, bgL in ((15, 15), (15, 11), (12, 19), (16, 16)):
            add('planted_bg_s{}_L{}_K{}'.format(seed, Lmax, bgL + 1), parts, Lmax, False, bg, bgL)
    for seed in range(2):
        parts = cases.palindrome_parts(seed)
        for internal in (False, True):
            add('palindromes_s{}_ir{}'.format(seed, int(internal)), parts, 5, internal)
    shorties = ['ACGTAC', 'GTACGT', 'AAAAAA', 'TTTTTT', 'ACGTACG', '', 'acgtac', 'A', 'T', 'AC', 'GT', 'ACGTACGA',
                'TCGTACGT', 'ACGTACGTTT']
    for internal in (False, True):
        add('short_parts_ir{}'.format(int(internal)), shorties, 6, internal)
    add('all_excluded', ['GAATTCGAATTC', 'ACGTACGTACGT'], 5, False)
    add('random_2000x100_L12', cases.random_parts(2000, 100, 1), 12, False)
    add('random_varlen_L24', cases.buffer_to_list(*cases.random_varlen_buffer(400, 50, 300, 5)), 24, False)
    add('longk_planted_L31', cases.planted_parts(200, 11, frag_lo=34, frag_hi=44), 31, False)

    with open(os.path.join(HERE, 'finder_golden.json'), 'w') as fh:
        json.dump({'partsets': partsets, 'cases': fixtures}, fh, separators=(',', ':'))
    print('wrote', len(fixtures), 'cases')

    # known-answer example of the reference's background docstring (nrpcalc.py:89-167)
    bg = ref.background(path='./kat_bg/', Lmax=15, verbose=False)
    doc_list = ['ATGAGATCGTAGCAACC', 'GACGATTACGTCAGGTA', 'ACAGTAGAGACGAGTAA', 'CCAGTACGAAAAGGCCC', 'TTAGCTTGATAGTTTTA']
    single = 'ATGCTTAGTGCCATACC'
    bg.add(single)
    bg.multiadd(doc_list)
    kat = {'Lmax': 15, 'add': single, 'multiadd': doc_list, 'len_after_add': len(bg), 'keys_after_add': sorted(bg),
           'contains_after_add': {s: (s in bg) for s in doc_list + [single, 'TTCGAGATTTTTCGAGT']}}
    bg.remove(single)
    kat['len_after_remove'] = len(bg)
    kat['keys_after_remove'] = sorted(bg)
    kat['contains_after_remove'] = {single: (single in bg)}
    bg.multiremove(doc_list)
    kat['len_after_multiremove'] = len(bg)
    bg.drop()
    with open(os.path.join(HERE, 'background_golden.json'), 'w') as fh:
        json.dump(kat, fh, indent=1)


if __name__ == '__main__':
    workdir = tempfile.mkdtemp(prefix='nrp_golden_')
    os.chdir(workdir)
    main()
