"""Generate tests/golden/finder_golden_general.json by running the UNMODIFIED reference (dev container only) on inputs outside
the 2-bit domain: RNA parts (U), other symbols (N, R, Y, '-'), windows longer than 32 bases, backgrounds with such keys.

    python tests/golden/make_golden_general.py

Same fixture layout as make_golden.py (graph node / adjacency order, seq_list.txt id order, toolboxes of the three vertex-cover
routines under two seeds).
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
from make_golden import run_reference   # noqa: E402


def to_rna(seq):
    return seq.replace('T', 'U').replace('t', 'u')


def sprinkle(parts, seed, symbols='NRY-', every=7):
    """Replace one character of every `every`-th part by another symbol, and turn some parts into RNA."""
    rnd = random.Random(seed)
    out = []
    for i, s in enumerate(parts):
        if s and i % every == 0:
            at = rnd.randrange(len(s))
            s = s[:at] + rnd.choice(symbols) + s[at + 1:]
        if i % 5 == 1:
            s = to_rna(s)
        out.append(s)
    return out


def main():
    ref = refshim.load_reference()
    fixtures = []
    partsets = {}

    def add(name, parts, Lmax, internal_repeats, bg_seqs=None, bg_Lmax=None):
        parts = list(parts)
        set_name = hashlib.sha256('\n'.join(parts).encode()).hexdigest()[:12]
        partsets[set_name] = parts
        case = {'name': name, 'partset': set_name, 'Lmax': Lmax, 'internal_repeats': internal_repeats, 'bg_seqs': bg_seqs, 'bg_Lmax': bg_Lmax}
        case.update(run_reference(ref, parts, Lmax, internal_repeats, bg_seqs, bg_Lmax))
        fixtures.append(case)
        print('{:40s} nodes {:5d} edges {:5d}'.format(name, len(case['graph_nodes']), sum(len(a) for a in case['graph_adj']) // 2))

    for seed in range(2):
        dna = cases.planted_parts(240, 40 + seed)
        rna = [to_rna(s) for s in dna]
        mixed = sprinkle(dna, seed)
        for Lmax in (12, 15):
            for internal in (False, True):
                add('rna_s{}_L{}_ir{}'.format(seed, Lmax, int(internal)), rna, Lmax, internal)
                add('mixed_s{}_L{}_ir{}'.format(seed, Lmax, int(internal)), mixed, Lmax, internal)
        bg = cases.background_seqs_for(dna, seed)
        bg_mixed = [to_rna(s) if i % 2 else s for i, s in enumerate(bg)] + ['ACGTNNACGTACGTTGCANACGTGT', mixed[0][:30]]
        for Lmax, bgL in ((15, 15), (15, 11), (12, 19)):
            add('mixed_bg_s{}_L{}_K{}'.format(seed, Lmax, bgL + 1), mixed, Lmax, False, bg_mixed, bgL)
            add('rna_bg_s{}_L{}_K{}'.format(seed, Lmax, bgL + 1), rna, Lmax, False, bg, bgL)
    # the complement table is not an involution on U: UUUUUU and TTTTTT share the canonical key AAAAAA without being a repeat
    quirk = ['UUUUUUCGCATTTTTTGG', 'GGAAAAAAGGCCATCA', 'CCUUUUUUCCGAGAUC', 'ACGUACGUAC', 'ACGTACGTAC', 'NNNNNNNN', 'NNNNNNNNACGT', 'ACGUAC', 'GUACGU',
             'acguac', 'ACGNAC', '', 'U', 'N', 'AUAUAU', 'ATATAT', 'GAAUUC', 'GAATTC', 'AAUU--AAUU']
    for internal in (False, True):
        add('quirks_ir{}'.format(int(internal)), quirk, 5, internal)
        add('quirks_L7_ir{}'.format(int(internal)), quirk, 7, internal)
    # windows longer than 32 bases (keys wider than 64 bits), DNA and mixed alphabets, with and without a background of the same width
    long_dna = cases.planted_parts(220, 77, lo=70, hi=160, frag_lo=42, frag_hi=70)
    long_mixed = sprinkle(long_dna, 9, every=11)
    for Lmax in (35, 49):
        for internal in (False, True):
            add('longk_dna_L{}_ir{}'.format(Lmax, int(internal)), long_dna, Lmax, internal)
        add('longk_mixed_L{}'.format(Lmax), long_mixed, Lmax, False)
    bg_long = cases.background_seqs_for(long_dna, 5, span=80)
    add('longk_dna_bg_L39_K40', long_dna, 39, False, bg_long, 39)
    add('longk_dna_bg_L20_K40', long_dna, 20, False, bg_long, 39)
    add('longk_mixed_bg_L39_K40', long_mixed, 39, False, bg_long + [long_mixed[0][:60]], 39)

    with open(os.path.join(HERE, 'finder_golden_general.json'), 'w') as fh:
        json.dump({'partsets': partsets, 'cases': fixtures}, fh, separators=(',', ':'))
    print('wrote', len(fixtures), 'cases')


if __name__ == '__main__':
    workdir = tempfile.mkdtemp(prefix='nrp_golden_')
    os.chdir(workdir)
    main()
