"""Test-only numpy model of the per-rank device calls of nrpcalc_b200.sharded (same five calls as NativeOps),
built from the oracle.  Lets the sharding / all-to-all / flag all-reduce / merge logic run under gloo on CPU."""
import numpy as np
import torch

from oracle import np_oracle


def _owner(keys, world):
    x = keys.astype(np.uint64)
    x = (x ^ (x >> np.uint64(29))) * np.uint64(0xBF58476D1CE4E5B9)
    return ((x >> np.uint64(40)) % np.uint64(world)).astype(np.int64)


class OracleOps(object):
    def __init__(self, bg_keys=None, bg_K=None):
        self.device = torch.device('cpu')
        self.bg_keys, self.bg_K = bg_keys, bg_K

    def load(self, seqs, ids):
        self.seqs = seqs
        self.ids = np.asarray(ids, dtype=np.int64)
        self.codes, self.offsets = np_oracle.encode(seqs)
        self.n_parts = len(seqs)
        self.flag_bytes = np.zeros(self.n_parts, dtype=np.uint8)

    def extract(self, lo, hi, k, internal_repeats, world):
        self.lo, self.hi = lo, hi
        part, j, canon, pal = np_oracle.canonical_records(self.codes, self.offsets, k)
        mine = (part >= lo) & (part < hi)
        part, canon, pal = part[mine], canon[mine], pal[mine]
        self.flag_bytes[:] = 0
        if not internal_repeats:
            self.flag_bytes[np.unique(part[pal])] |= np_oracle.FLAG_PALINDROME
        if self.bg_keys is not None and len(self.bg_keys):
            bpart, _, bcanon, _ = np_oracle.canonical_records(self.codes, self.offsets, int(self.bg_K))
            hit = np.isin(bcanon, self.bg_keys) & (bpart >= lo) & (bpart < hi)
            self.flag_bytes[np.unique(bpart[hit])] |= np_oracle.FLAG_BACKGROUND
        keep = self.flag_bytes[part] == 0
        part, canon = part[keep], canon[keep]
        owner = _owner(canon, world)
        order = np.argsort(owner, kind='stable')
        counts = np.bincount(owner, minlength=world).astype(np.int64)
        keys = torch.from_numpy(canon[order].view(np.int64).copy())
        vals = torch.from_numpy(part[order].astype(np.int32))
        return keys, vals, counts

    def group(self, keys, vals, k, internal_repeats):
        self.rec_keys = keys.numpy().view(np.uint64).copy()
        self.rec_vals = vals.numpy().astype(np.int64)
        order = np.lexsort((self.rec_vals, self.rec_keys))
        self.rec_keys, self.rec_vals = self.rec_keys[order], self.rec_vals[order]
        if not internal_repeats and self.rec_keys.size > 1:
            same = (self.rec_keys[1:] == self.rec_keys[:-1]) & (self.rec_vals[1:] == self.rec_vals[:-1])
            self.flag_bytes[np.unique(self.rec_vals[1:][same])] |= np_oracle.FLAG_REPEAT

    def flags(self):
        return torch.from_numpy(self.flag_bytes.copy())

    def set_flags(self, flags):
        self.flag_bytes = flags.numpy().copy()

    def finish(self, k, internal_repeats, want_edges):
        keep = (self.flag_bytes[self.rec_vals] & np_oracle.FLAG_REPEAT) == 0 if not internal_repeats else np.ones(self.rec_vals.size, bool)
        keys, vals = self.rec_keys[keep], self.rec_vals[keep]
        m = keys.size
        head = np.ones(m, dtype=bool)
        if m > 1:
            head[1:] = keys[1:] != keys[:-1]
        starts = np.nonzero(head)[0]
        sizes = np.diff(np.append(starts, m))
        sel = sizes >= 2
        indptr = np.zeros(int(sel.sum()) + 1, dtype=np.int64)
        np.cumsum(sizes[sel], out=indptr[1:])
        members = np.concatenate([vals[s:s + z] for s, z in zip(starts[sel], sizes[sel])]).astype(np.uint32) if sel.any() else np.zeros(0, np.uint32)
        gkeys = keys[starts[sel]]
        first_window = np.zeros(gkeys.size, dtype=np.uint32)
        for g, (key, s) in enumerate(zip(gkeys.tolist(), starts[sel].tolist())):
            p = int(vals[s])
            _, _, canon, _ = np_oracle.canonical_records(self.codes[self.offsets[p]:self.offsets[p + 1]], np.array([0, self.offsets[p + 1] - self.offsets[p]]), k)
            first_window[g] = int(np.nonzero(canon == np.uint64(key))[0][0])
        self.local = {'indptr': indptr, 'members': members, 'first_window': first_window, 'keys': gkeys.astype(np.uint64),
                      'candidates': int(m), 'oversized': 0}
        eu = ev = None
        if want_edges:
            pairs = set()
            for g in range(gkeys.size):
                uniq = sorted(set(int(self.ids[x]) for x in members[indptr[g]:indptr[g + 1]]))
                pairs.update((a, b) for i, a in enumerate(uniq) for b in uniq[i + 1:])
            pairs = sorted(pairs)
            eu = torch.tensor([a for a, _ in pairs], dtype=torch.int32)
            ev = torch.tensor([b for _, b in pairs], dtype=torch.int32)
        return torch.from_numpy(gkeys.astype(np.uint64).view(np.int64).copy()), eu, ev

    def last_single(self, all_keys, k):
        shared = np.sort(all_keys.numpy().view(np.uint64))
        out = np.full(self.n_parts, -1, dtype=np.int32)
        part, j, canon, _ = np_oracle.canonical_records(self.codes, self.offsets, k)
        ok = (part >= self.lo) & (part < self.hi) & (self.flag_bytes[part] == 0) & ~np.isin(canon, shared)
        np.maximum.at(out, part[ok], j[ok].astype(np.int32))
        self.last = out

    def merge_edges(self, codes):
        codes = np.unique(codes.numpy().view(np.uint64))
        self.edges = ((codes >> np.uint64(32)).astype(np.uint32), (codes & np.uint64(0xffffffff)).astype(np.uint32))

    def fetch_host(self, want_edges):
        out = dict(self.local)
        out.update({'last_single': self.last, 'flags': self.flag_bytes, 'edges': self.edges if want_edges else None})
        return out
