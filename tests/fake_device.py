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


# ---- string-level specification (any alphabet, any k) and a host model of the device context --------------------------------------
def string_spec(uniq, k, internal_repeats, background=None, preset=None):
    """What nrp_build / nrp_build_generic must return for processing-ordered [(id, SEQ)] parts (all at least k long), computed on
    strings with the literal port of the reference (oracle/ref_port.py): any alphabet, any k.  `background` answers `seq in
    background` (kmerSetDB.py:177-207); preset = flags decided elsewhere (non-zero: the part is excluded)."""
    from oracle import ref_port
    n = len(uniq)
    flags = np.zeros(n, dtype=np.uint8)
    table = {}
    n_records = 0
    for p, (_, seq) in enumerate(uniq):
        if preset is not None and preset[p]:
            flags[p] = preset[p]
            continue
        why = ref_port.part_is_excluded(seq, k, internal_repeats, background)
        if why is not None:
            flags[p] = {'direct': 1, 'inverted': 1, 'background': 4}[why]
            continue
        for j, w in enumerate(ref_port.windows(seq, k)):
            table.setdefault(min(w, ref_port.revcomp(w)), []).append((p, j))
            n_records += 1
    groups = [(recs[0][0], recs[0][1], [q for q, _ in recs]) for recs in table.values() if len(recs) > 1]
    last_single = np.full(n, -1, dtype=np.int32)
    for recs in table.values():
        if len(recs) == 1:
            p, j = recs[0]
            last_single[p] = max(last_single[p], j)
    edges = set()
    for _, _, members in groups:
        for a in members:
            for b in members:
                ia, ib = uniq[a][0], uniq[b][0]
                if ia < ib:
                    edges.add((ia, ib))
    return {'flags': flags, 'groups': groups, 'last_single': last_single, 'n_records': n_records, 'edges': sorted(edges)}


class _Keys2bit(object):
    """`seq in background` for a device key set: U -> T, any canonical K-window over A/C/G/T in the set (what the kernels probe)."""

    def __init__(self, keys, K):
        from nrpcalc_b200 import encoding
        self.K = int(K)
        self.keys = set(encoding.decode_kmers(keys, K)) if keys is not None and len(keys) else set()
        self.ALIVE = True

    def __contains__(self, seq):
        from oracle import ref_port
        return any(w in self.keys for w in ref_port.canonical_windows(seq.replace('U', 'T'), self.K))

    def multicheck(self, seqs):
        return (s in self for s in seqs)


class ModelContext(object):
    """Host model of nrpcalc_b200._native.Context for the calls hgraph makes on one device: load_parts, set_background, build,
    build_generic, fetch.  build() refuses what the 2-bit kernels refuse, so the routing of hgraph.repeat_structure is checked."""
    device = 0

    def __init__(self):
        self.bg = None
        self.calls = []

    def set_background(self, keys, K):
        self.bg = None if keys is None or len(keys) == 0 else _Keys2bit(keys, K)

    def load_parts(self, ascii_buf, offsets, part_id=None, device_ptr=None, chunks=1):
        blob = np.ascontiguousarray(ascii_buf, dtype=np.uint8).tobytes().decode('latin-1')
        off = offsets.tolist()
        ids = part_id if part_id is not None else np.arange(len(off) - 1, dtype=np.uint32)
        self.uniq = [(int(ids[i]), blob[off[i]:off[i + 1]]) for i in range(len(off) - 1)]

    def _finish(self, spec, want_edges):
        from nrpcalc_b200 import _native
        groups = spec['groups']
        indptr = np.zeros(len(groups) + 1, dtype=np.int64)
        np.cumsum([len(m) for _, _, m in groups], out=indptr[1:])
        members = np.array([q for _, _, m in groups for q in m], dtype=np.uint32)
        first_window = np.array([j for _, j, _ in groups], dtype=np.uint32)
        counts = {name: 0 for name in _native.COUNT_NAMES}
        counts.update(records=spec['n_records'], groups=len(groups), members=int(members.size), excluded=int((spec['flags'] != 0).sum()), runs=1)
        edges = None
        if want_edges:
            edges = (np.array([a for a, _ in spec['edges']], dtype=np.uint32), np.array([b for _, b in spec['edges']], dtype=np.uint32))
        self.result = _native.BuildResult(counts, {name: 0.0 for name in _native.TIMING_NAMES}, spec['flags'], indptr, members, first_window, None,
                                          spec['last_single'], edges)

    def build(self, k, internal_repeats, want_edges=True, early_fetch=False):
        assert 1 <= k <= 32, 'nrp_build: k outside [1, 32]'
        assert all(set(s) <= set('ACGT') for _, s in self.uniq), 'nrp_build: characters outside A/C/G/T'
        self.calls.append(('build', k))
        self._finish(string_spec(self.uniq, k, internal_repeats, self.bg), want_edges)

    def build_generic(self, k, internal_repeats, want_edges=True, preset_flags=None, **_):
        assert all(s == s.upper() for _, s in self.uniq)
        self.calls.append(('build_generic', k))
        self._finish(string_spec(self.uniq, k, internal_repeats, self.bg, preset_flags), want_edges)
        return 0

    def fetch(self, want_keys=True, copy=False):
        return self.result
