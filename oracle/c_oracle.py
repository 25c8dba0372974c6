"""ORACLE (test infrastructure, NOT product code): ctypes wrapper of oracle/nrp_oracle.c.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / reference leg may import this.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, '_build', 'libnrporacle.so')
_lib = None


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(_HERE, 'nrp_oracle.c')):
        subprocess.check_call(['make', '-C', _HERE, '-B', '_build/libnrporacle.so'], stdout=subprocess.DEVNULL)
    return LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(LIB_PATH)
        vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
        lib.nrp_oracle_build.restype = vp
        lib.nrp_oracle_build.argtypes = [vp, vp, i64, i32, i32, vp, i64, i32, i32, vp, vp]
        lib.nrp_oracle_build_slice.restype = vp
        lib.nrp_oracle_build_slice.argtypes = [vp, vp, i64, i32, i32, vp, i64, i32, i32, vp, vp, i32, i32, i32]
        lib.nrp_oracle_counts.argtypes = [vp, vp]
        lib.nrp_oracle_groups.argtypes = [vp, vp, vp, vp, vp, vp]
        lib.nrp_oracle_free.argtypes = [vp]
        _lib = lib
    return _lib


def _p(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def build_table(ascii_buf, offsets, k, internal_repeats, bg_keys=None, bg_K=0, n_threads=1, want_groups=True, n_slices=1):
    """Runs the C restatement on parts that all have len >= k (or are simply not windowed if shorter).

    n_slices > 1: the key space is processed in that many passes (one slice of the table in memory at a time; flags come
    from the first pass, last_single is max-accumulated, the groups of the slices are concatenated) -- same result, 1/n of
    the memory, for the full-size BASELINE configs.

    Returns dict(flags, last_single, n_records, n_groups, n_members[, indptr, members, rank_p, rank_j, keys])."""
    lib = _load()
    ascii_buf = np.ascontiguousarray(ascii_buf, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n = offsets.size - 1
    flags = np.zeros(n, dtype=np.uint8)
    last_single = np.zeros(n, dtype=np.int32)
    bg = None if bg_keys is None or len(bg_keys) == 0 else np.ascontiguousarray(np.sort(np.asarray(bg_keys, dtype=np.uint64)))
    n_slices = max(1, int(n_slices))
    out = {'flags': flags, 'last_single': last_single, 'n_records': 0, 'n_groups': 0, 'n_members': 0}
    pieces = []
    for sl in range(n_slices):
        handle = lib.nrp_oracle_build_slice(_p(ascii_buf), _p(offsets), n, int(k), 1 if internal_repeats else 0, _p(bg),
                                            0 if bg is None else bg.size, int(bg_K or 0), int(n_threads), _p(flags), _p(last_single),
                                            n_slices, sl, 1 if sl > 0 else 0)
        try:
            counts = np.zeros(4, dtype=np.int64)
            lib.nrp_oracle_counts(handle, _p(counts))
            out['n_records'] += int(counts[0]); out['n_groups'] += int(counts[1]); out['n_members'] += int(counts[2])
            if want_groups:
                ng, nm = int(counts[1]), int(counts[2])
                indptr = np.zeros(ng + 1, dtype=np.int64)
                members = np.zeros(nm, dtype=np.uint32)
                rank_p = np.zeros(ng, dtype=np.uint32)
                rank_j = np.zeros(ng, dtype=np.uint32)
                keys = np.zeros(ng, dtype=np.uint64)
                lib.nrp_oracle_groups(handle, _p(indptr), _p(members), _p(rank_p), _p(rank_j), _p(keys))
                pieces.append((indptr, members, rank_p, rank_j, keys))
        finally:
            lib.nrp_oracle_free(handle)
    if want_groups:
        base, indptrs = 0, [np.zeros(1, dtype=np.int64)]
        for piece in pieces:
            indptrs.append(piece[0][1:] + base)
            base += int(piece[0][-1])
        out.update({'indptr': np.concatenate(indptrs), 'members': np.concatenate([p[1] for p in pieces]),
                    'rank_p': np.concatenate([p[2] for p in pieces]), 'rank_j': np.concatenate([p[3] for p in pieces]),
                    'keys': np.concatenate([p[4] for p in pieces])})
    return out
