"""ORACLE (test infrastructure, NOT product code): vectorised numpy restatement of the SPECIFICATION
of the reference's Finder-mode repeat-graph construction, usable up to ~1e6 parts.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline leg may import this module.

Where oracle/ref_port.py replays the reference's container operations literally (and is therefore slow),
this module computes the same *results* with sorts:

  * which parts are excluded and why          (reference hgraph.py:50-79, kmerSetDB.py:177-207)
  * the canonical-window -> part-id groups     (reference hgraph.py:82-89), as (rank, member ids)
  * the ordered stream of distinct id tuples   (reference hgraph.py:130-135, dict.popitem order)
  * the undirected part-pair edge set          (reference hgraph.py:157-200, order-free)

Encoding follows SURVEY.md 8a: A=0 C=1 G=2 T=3 packed most-significant-base first, so that integer order
equals Python's string order ('A'<'C'<'G'<'T', utils.py:48-50 uses min() on strings) and the complement is
x ^ 3 (utils.py:10).  Parts no longer than k contribute their whole string as one key (utils.py:37-40);
such keys live in a separate key space per length (a string of length l < k can never equal a k-mer).

Parity status: PINNED through tests/test_oracle.py: equality with oracle/ref_port.py (itself pinned to the
reference's known-answer examples, golden fixtures and the live reference) on every generator in
tests/cases.py, and directly against the golden fixtures.

Domain: upper/lower-case A, C, G, T only.  The reference treats any other character as an opaque symbol;
the 2-bit path rejects it (documented deviation).
"""
import numpy as np

FLAG_REPEAT = 1        # two windows of the part share a canonical k-mer        (hgraph.py:53-69)
FLAG_PALINDROME = 2    # a window equals its own reverse complement             (hgraph.py:60-63, rmer == kmer)
FLAG_BACKGROUND = 4    # a canonical K-window of the part is a background key   (hgraph.py:72-79)

_CODE = np.full(256, 255, dtype=np.uint8)
for _i, _ch in enumerate('ACGT'):
    _CODE[ord(_ch)] = _i
    _CODE[ord(_ch.lower())] = _i


def encode(parts):
    """list[str] -> (codes uint8[total], offsets int64[n+1]); raises ValueError outside ACGT."""
    lengths = np.fromiter((len(s) for s in parts), dtype=np.int64, count=len(parts))
    offsets = np.zeros(len(parts) + 1, dtype=np.int64)
    np.cumsum(lengths, out=offsets[1:])
    raw = np.frombuffer(''.join(parts).encode('ascii'), dtype=np.uint8)
    codes = _CODE[raw]
    if codes.size and codes.max() == 255:
        raise ValueError('only A, C, G, T are supported by the 2-bit oracle')
    return codes, offsets


def _window_values(codes, k):
    """fwd[i], rc[i] = integer value of the k-window starting at base i and of its reverse complement
    (valid for i <= len-k; windows crossing part boundaries are masked out by the caller)."""
    n = codes.size - k + 1
    if n <= 0:
        return np.zeros(0, np.uint64), np.zeros(0, np.uint64)
    fwd = np.zeros(n, dtype=np.uint64)
    rev = np.zeros(n, dtype=np.uint64)
    c64 = codes.astype(np.uint64)
    for t in range(k):
        chunk = c64[t:t + n]
        fwd |= chunk << np.uint64(2 * (k - 1 - t))
        rev |= (np.uint64(3) - chunk) << np.uint64(2 * t)
    return fwd, rev


def canonical_records(codes, offsets, k):
    """Canonical keys of all k-windows of every part with length >= k.

    Returns (part int64[M], j int64[M], canon uint64[M], palin bool[M]) in (part, j) order.
    """
    lengths = np.diff(offsets)
    counts = np.maximum(lengths - k + 1, 0)
    total = int(counts.sum())
    if total == 0:
        z = np.zeros(0, np.int64)
        return z, z.copy(), np.zeros(0, np.uint64), np.zeros(0, bool)
    part = np.repeat(np.arange(len(lengths), dtype=np.int64), counts)
    first = np.zeros(len(lengths) + 1, dtype=np.int64)
    np.cumsum(counts, out=first[1:])
    j = np.arange(total, dtype=np.int64) - first[part]
    start = offsets[part] + j
    fwd, rev = _window_values(codes, k)
    f = fwd[start]
    r = rev[start]
    return part, j, np.minimum(f, r), f == r


def canonical_key_set(seqs, K):
    """Sorted unique canonical K-mers (uint64) of the given sequences after U->T -- the content of a
    kmerSetDB after multiadd (kmerSetDB.py:147-175).  Sequences shorter than K contribute their whole string
    in the reference; those keys can never match a K-window and are ignored here."""
    seqs = [s.replace('U', 'T').replace('u', 't') for s in seqs if s not in ('K', 'LEN')]
    codes, offsets = encode(seqs)
    _, _, canon, _ = canonical_records(codes, offsets, K)
    return np.unique(canon)


def spec(parts, k, internal_repeats, bg_keys=None, bg_K=None):
    """Order-exact specification of the hot path.

    parts: processing-ordered [(id, SEQ)] (oracle.ref_port.uniquify output); processing rank p = position.
    Returns a dict with
      flags       uint8[n]     exclusion bits per part (index = p)
      ids         int64[n]     part ids by p
      groups      list of (rank_p, rank_j, (ids...)) for every canonical key with >= 2 members, members in
                  (p, j) order with multiplicity
      last_single int64[n]     largest window index whose key is owned by that part alone, -1 if none
      edges       set of (min id, max id)
      n_records   number of (window, part) records of surviving parts
    """
    n = len(parts)
    ids = np.fromiter((pid for pid, _ in parts), dtype=np.int64, count=n)
    seqs = [s for _, s in parts]
    codes, offsets = encode(seqs)
    lengths = np.diff(offsets)
    flags = np.zeros(n, dtype=np.uint8)

    # records of every length class: class k for parts with L >= k, class L for shorter parts (whole string)
    rec_part, rec_j, rec_canon, rec_class, rec_pal = [], [], [], [], []
    part, j, canon, pal = canonical_records(codes, offsets, k)
    rec_part.append(part); rec_j.append(j); rec_canon.append(canon); rec_pal.append(pal)
    rec_class.append(np.full(part.size, k, dtype=np.int64))
    for short_len in np.unique(lengths[lengths < k]):
        short_len = int(short_len)
        idx = np.nonzero(lengths == short_len)[0]
        if short_len == 0:
            # the empty string: one window '', equal to its own reverse complement
            rec_part.append(idx.astype(np.int64)); rec_j.append(np.zeros(idx.size, np.int64))
            rec_canon.append(np.zeros(idx.size, np.uint64)); rec_pal.append(np.ones(idx.size, bool))
            rec_class.append(np.zeros(idx.size, np.int64))
            continue
        sub_codes = np.concatenate([codes[offsets[i]:offsets[i + 1]] for i in idx])
        sub_off = np.arange(idx.size + 1, dtype=np.int64) * short_len
        sp, sj, sc, spal = canonical_records(sub_codes, sub_off, short_len)
        rec_part.append(idx[sp].astype(np.int64)); rec_j.append(sj); rec_canon.append(sc); rec_pal.append(spal)
        rec_class.append(np.full(sp.size, short_len, dtype=np.int64))
    part = np.concatenate(rec_part); j = np.concatenate(rec_j); canon = np.concatenate(rec_canon)
    klass = np.concatenate(rec_class); pal = np.concatenate(rec_pal)

    if not internal_repeats:
        flags[np.unique(part[pal])] |= FLAG_PALINDROME
        order = np.lexsort((canon, klass, part))
        sp, sc, sk = part[order], canon[order], klass[order]
        same = (sp[1:] == sp[:-1]) & (sc[1:] == sc[:-1]) & (sk[1:] == sk[:-1])
        flags[np.unique(sp[1:][same])] |= FLAG_REPEAT

    if bg_keys is not None and len(bg_keys):
        bpart, _, bcanon, _ = canonical_records(codes, offsets, int(bg_K))   # only parts with L >= K have windows
        hit = np.isin(bcanon, np.asarray(bg_keys, dtype=np.uint64))
        flags[np.unique(bpart[hit])] |= FLAG_BACKGROUND

    keep = flags[part] == 0
    part, j, canon, klass = part[keep], j[keep], canon[keep], klass[keep]
    order = np.lexsort((j, part, canon, klass))
    part, j, canon, klass = part[order], j[order], canon[order], klass[order]
    m = part.size
    head = np.ones(m, dtype=bool)
    if m > 1:
        head[1:] = (canon[1:] != canon[:-1]) | (klass[1:] != klass[:-1])
    starts = np.nonzero(head)[0]
    sizes = np.diff(np.append(starts, m))

    last_single = np.full(n, -1, dtype=np.int64)
    single = starts[sizes == 1]
    if single.size:
        np.maximum.at(last_single, part[single], j[single])

    groups = []
    edges = set()
    for s, sz in zip(starts[sizes >= 2].tolist(), sizes[sizes >= 2].tolist()):
        members = tuple(int(x) for x in ids[part[s:s + sz]])
        groups.append((int(part[s]), int(j[s]), members))
        uniq = sorted(set(members))
        for a in range(len(uniq)):
            for b in range(a + 1, len(uniq)):
                edges.add((uniq[a], uniq[b]))
    return {'flags': flags, 'ids': ids, 'groups': groups, 'last_single': last_single, 'edges': edges,
            'n_records': int(m)}


def ordered_stream(result):
    """Distinct id tuples in the order the reference first meets them: keys are popped in reverse insertion
    order (hgraph.py:132-135), i.e. by DESCENDING (p, j) of their first member; only the first occurrence of
    each tuple changes the reference's set.  A part's single-member keys all give the tuple (id,), so only its
    last one matters."""
    entries = [(p, j, members) for p, j, members in result['groups']]
    ids = result['ids']
    for p in np.nonzero(result['last_single'] >= 0)[0].tolist():
        entries.append((p, int(result['last_single'][p]), (int(ids[p]),)))
    entries.sort(key=lambda e: (e[0], e[1]), reverse=True)
    seen = set()
    stream = []
    for _, _, members in entries:
        if members not in seen:
            seen.add(members)
            stream.append(members)
    return stream
