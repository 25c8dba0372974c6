"""Test-only stand-in for the device call of nrpcalc_b200.hgraph: fills a RepeatStructure from the numpy
oracle so that the HOST logic (stream ordering, clique replay, recovery, vertex cover, finder wrapper) can be
checked on a machine without a GPU.  Never imported by the product."""
import numpy as np

from nrpcalc_b200 import hgraph
from oracle import np_oracle


def oracle_repeat_structure(parts, background, homology, internal_repeats, ctx=None, want_edges=False):
    n = len(parts)
    ids = np.fromiter((pid for pid, _ in parts), dtype=np.uint32, count=n)
    bg_keys = bg_K = None
    if background is not None:
        if not background.ALIVE:
            raise RuntimeError('kmerSetDB was closed or dropped')
        bg_keys, _ = background.device_keys()
        bg_K = background.K
    spec = np_oracle.spec(parts, homology, internal_repeats, bg_keys, bg_K)
    structure = hgraph.RepeatStructure(n)
    structure.flags[:] = spec['flags']
    structure.records = spec['n_records']
    id_to_p = {int(pid): p for p, pid in enumerate(ids.tolist())}
    if spec['groups']:
        rank = np.array([(p << 32) | j for p, j, _ in spec['groups']], dtype=np.int64)
        sizes = [len(m) for _, _, m in spec['groups']]
        indptr = np.zeros(len(sizes) + 1, dtype=np.int64)
        np.cumsum(sizes, out=indptr[1:])
        members = np.array([id_to_p[x] for _, _, m in spec['groups'] for x in m], dtype=np.int64)
        structure.add_entries(rank, indptr, members)
    own = np.nonzero(spec['last_single'] >= 0)[0]
    if own.size:
        structure.add_entries((own.astype(np.int64) << 32) | spec['last_single'][own], np.arange(own.size + 1, dtype=np.int64), own)
    return structure, ids
