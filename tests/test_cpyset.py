"""csrc/cpyset.hpp -- the layout-exact model of CPython's set that the native host tail runs on -- against the interpreter
that runs this test: random operation sequences (add / discard / |= / &= / - / copy / difference with a dict, with colliding
keys, dummies and resizes, past the 50000-entry growth rule) must leave every set with the interpreter's iteration order, and
the tuple hash must equal hash()."""
import random

import numpy as np
import pytest

from nrpcalc_b200 import native_tail


@pytest.fixture(scope='module')
def lib():
    handle = native_tail._load()
    if handle is None:
        pytest.skip('nrpcalc_b200/libnrptail.so not built')
    return handle


def test_tuple_hash_equals_interpreter(lib):
    rng = random.Random(1)
    for _ in range(5000):
        n = rng.choice([0, 1, 1, 2, 2, 3, 5, 9])
        t = tuple(rng.choice([rng.randrange(0, 50), rng.randrange(0, 1 << 20), rng.randrange(0, 1 << 32)]) for _ in range(n))
        a = np.array(t, dtype=np.int64)
        assert lib.nrp_tail_tuple_hash(a.ctypes.data, n) == hash(t), t


def _run_script(lib, ops, n_sets, expect):
    arr = np.array(ops, dtype=np.int64)
    out = np.zeros(len(expect) + 16, dtype=np.int64)
    written = lib.nrp_tail_set_script(arr.ctypes.data, len(ops), n_sets, out.ctypes.data, out.size)
    return written == len(expect) and out[:written].tolist() == expect


def _random_case(lib, seed, n_sets=6, n_ops=400, key_range=200, bursts=False):
    rng = random.Random(seed)
    pool = [set() for _ in range(n_sets)]
    ops, expect = [], []

    def key():
        r = rng.random()
        if r < 0.6:
            return rng.randrange(0, key_range)
        if r < 0.8:
            return rng.randrange(0, key_range) * 8            # colliding low bits
        return rng.randrange(0, 1 << 32)

    for _ in range(n_ops):
        r = rng.random()
        a, b, c = rng.randrange(n_sets), rng.randrange(n_sets), rng.randrange(n_sets)
        if r < 0.45:
            for _ in range(rng.randrange(1, 300) if bursts and rng.random() < 0.3 else 1):
                k = key()
                pool[a].add(k)
                ops.append((1, a, k, 0))
        elif r < 0.62:
            k = rng.choice(sorted(pool[a])) if pool[a] and rng.random() < 0.8 else key()
            pool[a].discard(k)
            ops.append((2, a, k, 0))
        elif r < 0.70 and a != b:
            pool[a] |= pool[b]
            ops.append((3, a, b, 0))
        elif r < 0.77 and a != b:
            pool[a] &= pool[b]
            ops.append((4, a, b, 0))
        elif r < 0.84 and b != c:
            pool[a] = pool[b] - pool[c]
            ops.append((5, a, b, c))
        elif r < 0.88 and a != b:
            fresh = set()
            fresh |= pool[b]
            pool[a] = fresh
            ops.append((6, a, b, 0))
        elif r < 0.93 and b != c:
            pool[a] = pool[b].difference(dict.fromkeys(pool[c]))
            ops.append((8, a, b, c))
        elif r < 0.95:
            pool[a] = set()
            ops.append((0, a, 0, 0))
        else:
            ops.append((7, a, 0, 0))
            expect.append(len(pool[a]))
            expect.extend(pool[a])
    for a in range(n_sets):
        ops.append((7, a, 0, 0))
        expect.append(len(pool[a]))
        expect.extend(pool[a])
    return _run_script(lib, ops, n_sets, expect)


def test_random_operation_sequences(lib):
    rng = random.Random(7)
    for seed in range(700):
        assert _random_case(lib, seed, key_range=rng.choice([12, 60, 400, 5000]), n_ops=rng.choice([60, 300, 900]),
                            bursts=(seed % 5 == 0)), seed


def test_large_sets(lib):
    """Past 50000 entries (growth by 2x instead of 4x), bulk operations on large operands, a quarter of dummies."""
    rng = random.Random(3)
    A, B, ops = set(), set(), []
    for k in rng.sample(range(0, 400000), 70000):
        A.add(k)
        ops.append((1, 0, k, 0))
    for k in rng.sample(range(0, 400000), 30000):
        B.add(k)
        ops.append((1, 1, k, 0))
    for k in rng.sample(sorted(A), 25000):
        A.discard(k)
        ops.append((2, 0, k, 0))
    C = A - B
    ops.append((5, 2, 0, 1))
    D = set()
    D |= A
    ops.append((6, 3, 0, 0))
    D &= B
    ops.append((4, 3, 1, 0))
    B |= A
    ops.append((3, 1, 0, 0))
    E = A.difference(dict.fromkeys(D))
    ops.append((8, 4, 0, 3))
    expect = []
    for i, s in enumerate((A, B, C, D, E)):
        ops.append((7, i, 0, 0))
        expect.append(len(s))
        expect.extend(s)
    assert _run_script(lib, ops, 5, expect)
