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
    """kmerSetDB.test(): 1000 random 100-mers at K=99 -> 2000 keys; contains / remove / drop."""
    rnd = random.Random(3)
    seqs = [''.join(rnd.choice('ATGC') for _ in range(100)) for _ in range(1000)]
    db = kmerSetDB.kmerSetDB(path='./testDB', homology=99 if False else 32, verbose=False)
    db.multiadd(seqs)
    assert len(db) == 1000 * (100 - 32 + 1)
    for seq in seqs[:50]:
        assert seq in db
    assert all(db.multicheck(seqs))
    db.remove(seqs[0])
    assert seqs[0] not in db
    db.multiremove(seqs)
    assert len(db) == 0
    assert db.drop() is True
    assert not os.path.exists('./testDB')


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
