"""Canonical form, digest and comparison of a repeat structure (test infrastructure, shared by tests/ and bench.py).

A repeat structure is what the device hands to the host replay: per-part exclusion flags, the
shared k-mers as groups (key, first window, members in (part, window) order with multiplicity), the last unshared
window of every part, and the distinct part-pair edges.  Different producers emit the groups in different orders
(the GPU by leaf bucket, the C oracle by table slot, the sharded path by owner rank); the canonical form orders them
by k-mer, which is unique per group, so that two structures are equal iff their canonical arrays are equal.

Reference semantics: nrpcalc/base/hgraph.py:23-94 (table), 130-135 (tuples), 157-200 (edges).
"""
import hashlib

import numpy as np


def reorder_csr(indptr, members, order):
    """CSR rows permuted by `order` (rows keep their inner order).  Returns (new_indptr, new_members)."""
    indptr = np.asarray(indptr, dtype=np.int64)
    sizes = np.diff(indptr)[order]
    new_indptr = np.zeros(sizes.size + 1, dtype=np.int64)
    np.cumsum(sizes, out=new_indptr[1:])
    total = int(new_indptr[-1])
    if total == 0:
        return new_indptr, np.zeros(0, dtype=np.asarray(members).dtype)
    shift = np.repeat(indptr[:-1][order] - new_indptr[:-1], sizes)
    return new_indptr, np.asarray(members)[np.arange(total, dtype=np.int64) + shift]


def edges_from_groups(indptr, members, ids):
    """Sorted distinct codes (min(id) << 32 | max(id)) of all pairs of different parts inside a group
    (hgraph.py:157-200: an edge for every two parts that share a canonical k-mer)."""
    indptr = np.asarray(indptr, dtype=np.int64)
    sizes = np.diff(indptr)
    part = np.asarray(ids, dtype=np.uint64)[np.asarray(members, dtype=np.int64)] if len(members) else np.zeros(0, dtype=np.uint64)
    codes = []
    for s in np.unique(sizes):
        s = int(s)
        if s < 2:
            continue
        rows = np.nonzero(sizes == s)[0]
        block = part[(indptr[rows][:, None] + np.arange(s, dtype=np.int64)[None, :]).ravel()].reshape(rows.size, s)
        iu, ju = np.triu_indices(s, 1)
        a, b = block[:, iu].ravel(), block[:, ju].ravel()
        keep = a != b
        a, b = a[keep], b[keep]
        codes.append((np.minimum(a, b) << np.uint64(32)) | np.maximum(a, b))
    if not codes:
        return np.zeros(0, dtype=np.uint64)
    return np.unique(np.concatenate(codes))


def canonical(flags, indptr, members, first_window, keys, last_single, edges=None, ids=None, records=None):
    """Canonical arrays of a repeat structure.  `members` are part indices (processing rank); `edges` is a pair of
    id arrays (u, v) or None, in which case they are derived from the groups through `ids`."""
    keys = np.asarray(keys, dtype=np.uint64)
    order = np.argsort(keys, kind='stable')
    new_indptr, new_members = reorder_csr(indptr, members, order)
    if edges is not None:
        u, v = np.asarray(edges[0], dtype=np.uint64), np.asarray(edges[1], dtype=np.uint64)
        codes = np.sort((np.minimum(u, v) << np.uint64(32)) | np.maximum(u, v))
    elif ids is not None:
        codes = edges_from_groups(new_indptr, new_members, ids)
    else:
        codes = None
    return {'excluded': (np.asarray(flags) != 0).astype(np.uint8), 'keys': keys[order], 'indptr': new_indptr,
            'members': new_members.astype(np.uint32), 'first_window': np.asarray(first_window, dtype=np.uint32)[order],
            'last_single': np.asarray(last_single, dtype=np.int32), 'edges': codes,
            'records': None if records is None else int(records)}


def from_c_oracle(res, ids=None):
    """Canonical form of oracle/c_oracle.build_table's result (groups carry rank_j = first window)."""
    return canonical(res['flags'], res['indptr'], res['members'], res['rank_j'], res['keys'], res['last_single'], None, ids,
                     res.get('n_records'))


COMPONENTS = ('excluded', 'keys', 'indptr', 'members', 'first_window', 'last_single', 'edges')


def compare(a, b):
    """Component-wise equality of two canonical forms: {'excluded': bool, ..., 'all': bool}."""
    out = {}
    for name in COMPONENTS:
        x, y = a.get(name), b.get(name)
        if x is None or y is None:
            continue
        out[name] = bool(x.shape == y.shape and np.array_equal(x, y))
    if a.get('records') is not None and b.get('records') is not None:
        out['records'] = a['records'] == b['records']
    out['all'] = all(out.values())
    return out


def digest(c):
    """sha256 over the canonical arrays (the edge list included when present)."""
    h = hashlib.sha256()
    for name in COMPONENTS:
        x = c.get(name)
        if x is None:
            continue
        h.update(name.encode())
        h.update(np.ascontiguousarray(x).tobytes())
    return h.hexdigest()


def summary(c):
    return {'excluded': int(c['excluded'].sum()), 'groups': int(c['keys'].size), 'members': int(c['members'].size),
            'edges': None if c['edges'] is None else int(c['edges'].size), 'records': c.get('records')}
