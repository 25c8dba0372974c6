"""CPU tests of the packed Finder entry (nrpcalc_b200/packed.py): the host logic around the device calls -- argument checks,
routing between the device entry and the string path, id / source bookkeeping, output dict -- with the two device calls
(nrp_uniquify, the build) replaced by host models.  Expected values: the golden fixtures of the unmodified reference."""
import random

import numpy as np
import pytest

import fake_device
import nrpcalc_b200
from conftest import GOLDEN_CASES
from nrpcalc_b200 import hgraph, packed


class ModelContext(object):
    """Host model of Context.uniquify (include/nrp_b200.h: nrp_uniquify): upper-case, drop duplicates, id = first occurrence,
    processing order = descending last occurrence; status bit 1 = a character outside A/C/G/T."""
    collide = False

    def uniquify(self, ascii_buf, offsets=None, part_len=0):
        buf = np.ascontiguousarray(ascii_buf, dtype=np.uint8)
        if offsets is None:
            n = buf.size // part_len if part_len else 0
            offsets = np.arange(n + 1, dtype=np.int64) * part_len
        n = offsets.size - 1
        parts = [buf[offsets[i]:offsets[i + 1]].tobytes().decode('latin-1').upper() for i in range(n)]
        first, last = {}, {}
        for i, s in enumerate(parts):
            first.setdefault(s, i)
            last[s] = i
        order = sorted(last.values(), reverse=True)
        ids = np.array([first[parts[i]] for i in order], dtype=np.uint32)
        source = np.array(order, dtype=np.uint32)
        new_off = np.zeros(len(order) + 1, dtype=np.int64)
        np.cumsum([len(parts[i]) for i in order], out=new_off[1:])
        status = 1 if any(set(s) - set('ACGT') for s in parts) else 0
        self.gathered = [parts[i] for i in order]
        return ids, source, new_off, 1234, status | (2 if self.collide else 0), max([len(s) for s in parts] or [0])


@pytest.fixture
def model(monkeypatch, scratch_cwd):
    ctx = ModelContext()
    monkeypatch.setattr(hgraph, '_context', lambda device=None: ctx)
    monkeypatch.setattr(hgraph, 'repeat_structure', fake_device.oracle_repeat_structure)

    def device_model(ctx_, device_ptr, offsets, ids, longest, background, homology, internal_repeats, want_edges=False):
        assert device_ptr == 1234 and offsets.size - 1 == len(ctx_.gathered) and longest == max(len(s) for s in ctx_.gathered)
        structure, _ = fake_device.oracle_repeat_structure(list(zip(ids.tolist(), ctx_.gathered)), background, homology, internal_repeats)
        return structure

    monkeypatch.setattr(hgraph, 'repeat_structure_device', device_model)
    return ctx


def _pack(parts):
    offsets = np.zeros(len(parts) + 1, dtype=np.int64)
    np.cumsum([len(s) for s in parts], out=offsets[1:])
    return ''.join(parts).encode('ascii'), offsets


def _background(case):
    if case['bg_seqs'] is None:
        return None
    bg = nrpcalc_b200.background(path='./bg_store/', Lmax=case['bg_Lmax'], verbose=False)
    bg.multiadd(case['bg_seqs'])
    return bg


@pytest.mark.parametrize('case', GOLDEN_CASES, ids=[c['name'] for c in GOLDEN_CASES])
def test_packed_toolbox_matches_reference(case, model):
    bg = _background(case)
    blob, offsets = _pack(case['parts'])
    short = any(len(s) <= case['Lmax'] for s in case['parts'])
    for vc, by_seed in case['toolbox'].items():
        for seed, expected_ids in by_seed.items():
            random.seed(int(seed))
            found = nrpcalc_b200.finder_packed(blob, case['Lmax'], offsets=offsets, internal_repeats=case['internal_repeats'], background=bg, vercov=vc)
            assert list(found.keys()) == expected_ids, (vc, seed)
            assert all(found[i] == case['parts'][i].upper() for i in expected_ids)
            assert packed.last_run_info['path'] == ('strings' if short or not case['parts'] else 'device')
    if bg is not None:
        bg.drop()


def test_uniform_length_entry_and_buffer_types(model):
    import cases
    parts = cases.random_parts(300, 40, 3)
    parts[7] = parts[3].lower()
    parts[100] = parts[3]
    blob = ''.join(parts).encode()
    random.seed(1)
    want = nrpcalc_b200.finder(list(parts), 11, verbose=False)
    for buffer in (blob, bytearray(blob), memoryview(blob), np.frombuffer(blob, dtype=np.uint8)):
        random.seed(1)
        got = nrpcalc_b200.finder_packed(buffer, 11, part_len=40)
        assert got == want and list(got) == list(want)
        assert packed.last_run_info['path'] == 'device' and packed.last_run_info['unique'] == 298
    assert 7 not in want and 100 not in want


def test_routing_to_the_string_path(model):
    import cases
    parts = cases.random_parts(50, 30, 4)
    blob, offsets = _pack(parts)
    random.seed(2)
    want = nrpcalc_b200.finder(list(parts), 9, verbose=False)
    model.collide = True                                       # a reported hash collision: the strings decide
    random.seed(2)
    assert nrpcalc_b200.finder_packed(blob, 9, offsets=offsets) == want
    assert packed.last_run_info == {'path': 'strings', 'reason': 'hash collision between different parts'}
    model.collide = False
    with pytest.raises(ValueError, match='A, C, G, T'):        # other alphabets: what finder() says
        nrpcalc_b200.finder_packed(b'ACGTNACGTACGTACGTACGTAAAA', 9, part_len=25)
    random.seed(2)
    assert nrpcalc_b200.finder_packed(b'', 9, part_len=25) == {}


def test_argument_checks(model):
    with pytest.raises(ValueError, match='exactly one'):
        nrpcalc_b200.finder_packed(b'ACGT', 9)
    with pytest.raises(ValueError, match='exactly one'):
        nrpcalc_b200.finder_packed(b'ACGT', 9, offsets=[0, 4], part_len=4)
    with pytest.raises(ValueError, match='whole number'):
        nrpcalc_b200.finder_packed(b'ACGTA', 9, part_len=4)
    with pytest.raises(ValueError, match='offsets must'):
        nrpcalc_b200.finder_packed(b'ACGTA', 9, offsets=[0, 3, 2])
    with pytest.raises(TypeError):
        nrpcalc_b200.finder_packed('ACGT', 9, part_len=4)
    with pytest.raises(RuntimeError, match='Invalid Constraints'):
        nrpcalc_b200.finder_packed(b'ACGTACGT', 3, part_len=4)
