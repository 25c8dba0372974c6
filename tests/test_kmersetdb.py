"""kmerSetDB drop-in: the reference's docstring known answers (nrpcalc.py:89-167), its inline self test
(kmerSetDB.py:332-423) and persistence."""
import json
import os
import random

import numpy as np
import pytest

import nrpcalc_b200
from nrpcalc_b200 import encoding, kmerSetDB
from oracle import np_oracle, ref_port

HERE = os.path.dirname(os.path.abspath(__file__))


def test_docstring_example(scratch_cwd):
    with open(os.path.join(HERE, 'golden', 'background_golden.json')) as fh:
        kat = json.load(fh)
    bkg = nrpcalc_b200.background(path='./prj_bkg/', Lmax=kat['Lmax'], verbose=False)
    assert repr(bkg) == 'kmerSetDB stored at ./prj_bkg/ with 0 16-mers'
    bkg.add(kat['add'])
    bkg.multiadd(kat['multiadd'])
    assert kat['add'] in bkg
    assert all(bkg.multicheck(kat['multiadd']))
    assert sorted(bkg) == kat['keys_after_add']
    assert len(bkg) == kat['len_after_add'] == 12
    for seq, expected in kat['contains_after_add'].items():
        assert (seq in bkg) == expected
    bkg.remove(kat['add'])
    assert (kat['add'] in bkg) == kat['contains_after_remove'][kat['add']]
    assert len(bkg) == kat['len_after_remove'] == 10
    assert sorted(bkg) == kat['keys_after_remove']
    bkg.multiremove(kat['multiadd'])
    assert len(bkg) == kat['len_after_multiremove'] == 0
    bkg.clear(Lmax=12)
    assert bkg.K == 13 and len(bkg) == 0 and bkg.ALIVE
    assert bkg.close() is True and bkg.close() is False
    with pytest.raises(RuntimeError, match='closed or dropped'):
        'ACGT' in bkg
    assert bkg.drop() is False


def test_reference_selftest(scratch_cwd):
    """kmerSetDB.test() (kmerSetDB.py:332-423): 1000 random 100-mers at K=99 -> 2000 keys; contains / remove / drop.
    K=99 keys are wider than 64 bits: every window lives in the verbatim side set."""
    rnd = random.Random(3)
    seqs = [''.join(rnd.choice('ATGC') for _ in range(100)) for _ in range(1000)]
    db = kmerSetDB.kmerSetDB(path='./testDB', homology=99, verbose=False)
    db.multiadd(seqs)
    assert len(db) == 2000
    assert db.device_keys()[0].size == 0
    for seq in seqs[:50]:
        assert seq in db
    assert all(db.multicheck(seqs))
    db.remove(seqs[0])
    assert seqs[0] not in db and len(db) == 1998
    db.multiremove(seqs)
    assert len(db) == 0
    assert db.drop() is True
    assert not os.path.exists('./testDB')


def test_selftest_shape_at_K32(scratch_cwd):
    """The same sequence of calls at K=32, where every window is a 2-bit key."""
    rnd = random.Random(3)
    seqs = [''.join(rnd.choice('ATGC') for _ in range(100)) for _ in range(1000)]
    db = kmerSetDB.kmerSetDB(path='./testDB', homology=32, verbose=False)
    for seq in seqs[:10]:
        db.add(seq)                                  # single adds are buffered, LEN stays exact
    assert len(db) == 10 * (100 - 32 + 1)
    db.multiadd(seqs)
    assert len(db) == 1000 * (100 - 32 + 1)
    assert all(db.multicheck(seqs))
    db.remove(seqs[0])
    assert seqs[0] not in db
    db.multiremove(seqs)
    assert len(db) == 0
    assert db.drop() is True


@pytest.mark.parametrize('K', [7, 16, 20])
def test_other_characters_spoil_only_their_windows(scratch_cwd, K):
    """An 'N', a lower-case stretch or a 'U' inside a background sequence: every window is classified on its own, exactly
    as the reference's string keys behave (model: oracle/ref_port.SetBackground, pinned to the reference's kmerSetDB)."""
    rnd = random.Random(K)
    dna = lambda n: ''.join(rnd.choice('ACGT') for _ in range(n))
    chromosome = dna(60) + 'N' + dna(45) + 'NNNN' + dna(K - 1) + 'N' + dna(30) + dna(25).lower() + dna(40) + 'U' + dna(33)
    others = [dna(50), 'ACGTN', dna(K), dna(K - 2), 'acgt' * 8, dna(20) + 'R' + dna(20)]
    model = ref_port.SetBackground(K)
    db = kmerSetDB.kmerSetDB(path='./nDB', homology=K, verbose=False)
    model.add(chromosome); db.add(chromosome)
    model.multiadd(others); db.multiadd(others)
    assert len(db) == len(model)
    assert sorted(db) == sorted(model.keys)
    plain = sorted(key for key in model.keys if len(key) == K and set(key) <= set('ACGT'))
    assert encoding.decode_kmers(db.device_keys()[0], K) == plain          # everything the device can match is uploaded
    probes = [chromosome[5:5 + K + 3], chromosome[58:58 + K + 2], dna(K + 5), 'N' * K, chromosome[-(K + 1):], others[0][3:3 + K],
              others[-1], chromosome[61 - K:61], chromosome.lower()[:K + 4]]
    for probe in probes:
        assert (probe in db) == (probe in model), probe
    assert list(db.multicheck(probes)) == list(model.multicheck(probes))
    db.remove(chromosome); model.remove(chromosome)
    assert len(db) == len(model) and sorted(db) == sorted(model.keys)
    # clear=True decrements once per streamed window, present or not (kmerSetDB.py:267-269)
    db.multiremove([others[0], others[0][:K + 3]], clear=True); model.multiremove([others[0], others[0][:K + 3]], clear=True)
    assert db.LEN == model.LEN and sorted(db) == sorted(model.keys)
    db.close()
    again = kmerSetDB.kmerSetDB(path='./nDB', homology=K, verbose=False)
    assert again.LEN == model.LEN and sorted(again) == sorted(model.keys)
    again.drop()


def test_keys_match_oracle_and_persist(scratch_cwd):
    rnd = random.Random(8)
    seqs = [''.join(rnd.choice('ACGT') for _ in range(rnd.randint(10, 80))) for _ in range(60)] + ['ACGU' * 6, 'acgtacgtacgtacgtacgt', 'ACG']
    db = nrpcalc_b200.background('./store', 15, verbose=False)
    db.multiadd(seqs)
    plain = [s.replace('U', 'T') for s in seqs if s.isupper() and len(s) >= 16]
    keys, version = db.device_keys()
    assert keys.tolist() == np_oracle.canonical_key_set(plain, 16).tolist()
    n_before = len(db)
    model = ref_port.SetBackground(16)             # the oracle's model of the reference's kmerSetDB
    model.multiadd(seqs)
    assert n_before == len(model) and sorted(db) == sorted(model.keys)
    db.close()
    again = nrpcalc_b200.background('./store', 20, verbose=False)      # stored K wins over the argument
    assert again.K == 16 and len(again) == n_before
    assert again.device_keys()[0].tolist() == keys.tolist()
    assert 'ACG' in again and 'ACGU' * 6 in again
    again.drop()


def test_constructor_errors(scratch_cwd):
    with pytest.raises(ValueError, match='Invalid Path or Lmax'):
        nrpcalc_b200.background(5, 12, verbose=False)
    with pytest.raises(ValueError):
        nrpcalc_b200.background('./x', 3, verbose=False)
    with pytest.raises(ValueError):
        kmerSetDB.kmerSetDB('./x', 'a')


def test_encoding_roundtrip():
    rnd = random.Random(1)
    for K in (1, 6, 16, 17, 31, 32):
        seq = ''.join(rnd.choice('ACGT') for _ in range(K + 20))
        buf, off = encoding.join_parts([seq])
        vals = encoding.canonical_kmers(buf, off, K)
        comp = str.maketrans('ACGT', 'TGCA')
        want = [min(seq[i:i + K], seq[i:i + K].translate(comp)[::-1]) for i in range(len(seq) - K + 1)]
        assert encoding.decode_kmers(vals, K) == want


def test_against_the_live_reference_kmersetdb(scratch_cwd):
    """The same calls on the UNMODIFIED reference's kmerSetDB (ShareDB replaced by a dict, tests/golden/refshim.py) and on the
    drop-in: equal key sets, counters and membership answers, including sequences with other characters."""
    import sys
    sys.path.insert(0, os.path.join(HERE, 'golden'))
    import refshim
    if not refshim.reference_available():
        pytest.skip('no reference tree (baseline/_ref or /root/reference)')
    ref = refshim.load_reference()
    rnd = random.Random(11)
    dna = lambda n: ''.join(rnd.choice('ACGT') for _ in range(n))
    seqs = [dna(70) + 'N' + dna(50), dna(40), dna(30).lower() + dna(30), 'AUGC' * 9, dna(16), dna(9), dna(25) + 'NN' + dna(25)]
    theirs = ref.background(path='./ref_bg/', Lmax=15, verbose=False)
    mine = nrpcalc_b200.background(path='./my_bg/', Lmax=15, verbose=False)
    theirs.add(seqs[0]); mine.add(seqs[0])
    theirs.multiadd(seqs[1:]); mine.multiadd(seqs[1:])
    assert len(mine) == len(theirs) and sorted(key for key in mine) == sorted(key for key in theirs)
    probes = [seqs[0][10:40], seqs[0][60:90], dna(40), seqs[2][20:50], 'AUGC' * 5, seqs[5], seqs[6][15:45]]
    assert [p in mine for p in probes] == [p in theirs for p in probes]
    assert list(mine.multicheck(probes)) == list(theirs.multicheck(probes))
    theirs.remove(seqs[1]); mine.remove(seqs[1])
    theirs.multiremove(seqs[2:4]); mine.multiremove(seqs[2:4])
    assert len(mine) == len(theirs) and sorted(key for key in mine) == sorted(key for key in theirs)
    assert repr(mine).split(' with ')[1] == repr(theirs).split(' with ')[1]
    theirs.drop(); mine.drop()
