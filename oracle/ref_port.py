"""ORACLE (test infrastructure, NOT product code): literal CPU restatement of the reference's
Finder-mode repeat-graph construction, in pure Python, for small cases.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline leg may import this module.
The product package `nrpcalc_b200` never does.

It restates, operation for operation where the order of CPython container operations is
observable downstream, these reference functions (paths relative to /root/reference/nrpcalc/base):

  windows()            <- utils.stream_kmers            utils.py:37-40
  revcomp()            <- utils.get_comp / get_revcomp  utils.py:10,42-46
  canonical_windows()  <- utils.stream_min_kmers        utils.py:48-50
  SetBackground        <- kmerSetDB membership          kmerSetDB.py:119-127,147-175,177-207
  uniquify()           <- hgraph.uniquify_seq_list      hgraph.py:11-21 (+ the reversal at 218-223)
  repeat_table()       <- hgraph.get_repeat_dict        hgraph.py:23-94
  maximal_cliques()    <- hgraph.get_maximal_cliques    hgraph.py:96-115
  clique_stream()      <- hgraph.get_repeat_cliques     hgraph.py:117-155
  assemble_graph()     <- hgraph.build_homology_graph   hgraph.py:157-200
  homology_graph()     <- hgraph.get_homology_graph     hgraph.py:202-249

Parity status: PINNED.  tests/test_oracle.py checks it against (i) the reference's docstring
known-answer examples (nrpcalc.py:89-167 and 293-378), (ii) fixtures under tests/golden/ that
tests/golden/make_golden.py produced by running the unmodified reference in the dev container,
and (iii) -- when /root/reference is present -- the live reference on seeded random inputs.
"""
from collections import defaultdict, deque

import networkx as nx

_COMPLEMENT = str.maketrans('ATGCU', 'TACGA')          # utils.py:10 (U->A but A->T: asymmetric on purpose)


def windows(seq, k):
    """All length-k substrings; the whole string when k >= len(seq) (utils.py:37-40)."""
    n = len(seq)
    if k >= n:
        return [seq]
    return [seq[i:i + k] for i in range(n - k + 1)]


def revcomp(seq):
    """Translate with the reference's table, then reverse (utils.py:42-46)."""
    return seq.translate(_COMPLEMENT)[::-1]


def canonical_windows(seq, k):
    """Lexicographic min of each window and its reverse complement (utils.py:48-50)."""
    return [min(w, revcomp(w)) for w in windows(seq, k)]


class SetBackground(object):
    """In-memory model of the kmerSetDB behaviour that the hot path observes.

    add/multiadd: U->T, then every canonical K-window becomes a key (kmerSetDB.py:129-175).
    contains(seq): U->T, True iff ANY canonical K-window is a key (kmerSetDB.py:177-192; the code says
    "any", the docstring at nrpcalc.py:116 says "all" -- code wins).
    """

    def __init__(self, K):
        self.K = int(K)
        self.ALIVE = True
        self.keys = set()
        self.LEN = 0

    def multiadd(self, seq_list):
        for seq in seq_list:
            if seq in ('K', 'LEN'):                      # metadata keys share the keyspace (kmerSetDB.py:159)
                continue
            for key in canonical_windows(seq.replace('U', 'T'), self.K):
                if key not in self.keys:                 # kmerSetDB.py:168-170
                    self.keys.add(key)
                    self.LEN += 1

    def add(self, seq):
        self.multiadd([seq])

    def multiremove(self, seq_list, clear=False):
        """kmerSetDB.py:246-280: with clear the counter drops once per streamed window, present or not."""
        for seq in seq_list:
            if seq in ('K', 'LEN'):
                continue
            for key in canonical_windows(seq.replace('U', 'T'), self.K):
                if clear:
                    self.keys.discard(key)
                    self.LEN -= 1
                elif key in self.keys:
                    self.keys.remove(key)
                    self.LEN -= 1

    def remove(self, seq):
        self.multiremove([seq])

    def contains(self, seq):
        if not self.ALIVE:
            raise RuntimeError('kmerSetDB was closed or dropped')   # kmerSetDB.py:97-102
        seq = seq.replace('U', 'T')
        return any(w in self.keys for w in canonical_windows(seq, self.K))

    __contains__ = contains

    def multicheck(self, seq_list):
        for seq in seq_list:
            yield self.contains(seq)

    def __len__(self):
        return self.LEN                                  # kmerSetDB.py:219-225

    def __iter__(self):
        return iter(self.keys)


def uniquify(seq_list):
    """Processing order of the reference: [(id, SEQ)], id = FIRST index of the upper-cased string,
    ordered by DESCENDING last-occurrence index.

    hgraph.py:11-21 walks the list backwards into a dict (insertion order = descending last occurrence,
    the stored value ends up as the first occurrence), pops it into a list (ascending), and
    hgraph.py:218-223 pops that list from the end again (descending).  The two reversals cancel.
    """
    first_index = {}
    for idx in range(len(seq_list) - 1, -1, -1):
        first_index[seq_list[idx].upper()] = idx
    return [(idx, seq) for seq, idx in first_index.items()]


def part_is_excluded(seq, k, internal_repeats, background):
    """The two `continue` branches of hgraph.py:50-79.  Returns a reason string or None."""
    fwd = windows(seq, k)
    if not internal_repeats:
        distinct = set(fwd)
        if len(distinct) < len(fwd):
            return 'direct'                              # hgraph.py:54-57
        for w in fwd:
            if revcomp(w) in distinct:                   # fires for a palindromic window too
                return 'inverted'                        # hgraph.py:59-69
    if background is not None and background.K <= len(seq):
        if background.K == k:
            if any(background.multicheck(fwd)):          # hgraph.py:74-76
                return 'background'
        elif seq in background:                          # hgraph.py:77-79
            return 'background'
    return None


def repeat_table(parts, background, k, internal_repeats):
    """dict canonical-window -> [ids in processing order, with multiplicity] (hgraph.py:23-94).

    `parts` is the processing-ordered [(id, SEQ)] list.  Dict insertion order is observable
    (popitem order in clique_stream), so keys are inserted exactly as the reference does: part by part,
    window by window.
    """
    table = {}
    for part_id, seq in parts:
        if part_is_excluded(seq, k, internal_repeats, background) is not None:
            continue
        for w in windows(seq, k):
            key = min(w, revcomp(w))
            bucket = table.get(key)
            if bucket is None:
                table[key] = [part_id]
            else:
                bucket.append(part_id)
    return table


def maximal_cliques(clique_sets):
    """Drop cliques that are subsets of another one (hgraph.py:96-115).

    The sequence of set operations is kept identical to the reference because the iteration order of the
    resulting set of indices decides the order in which cliques reach the graph.
    """
    member_of = defaultdict(set)
    for pos, clique in enumerate(clique_sets):
        for node in clique:
            member_of[node].add(pos)
    keep = set()
    for pos, clique in enumerate(clique_sets):
        common = set()
        for node in clique:
            if len(common) == 0:
                common |= member_of[node]
            else:
                common &= member_of[node]
        if len(common) == 1:
            keep.add(pos)
        else:
            keep.update(common - set([pos]))
    return [clique_sets[pos] for pos in keep]


def clique_stream(table):
    """Distinct id-tuples -> sets -> maximal cliques -> deques, in the reference's order (hgraph.py:130-155)."""
    distinct = set()
    while table:
        _, ids = table.popitem()                         # LIFO: last inserted key first
        distinct.add(tuple(ids))
    as_sets = []
    while distinct:
        as_sets.append(set(distinct.pop()))
    as_sets = maximal_cliques(as_sets)
    final = set()
    while as_sets:
        final.add(tuple(as_sets.pop()))
    while final:
        yield deque(final.pop())


def assemble_graph(cliques):
    """Insert cliques into an nx.Graph exactly as hgraph.py:157-200 does (node and adjacency order)."""
    graph = nx.Graph()
    for clique in cliques:
        if len(clique) == 1:
            graph.add_node(clique.popleft())
            continue
        pending = set(clique)
        while clique:
            u = clique.popleft()
            pending.remove(u)
            graph.add_node(u)
            for v in pending - set(graph[u]):
                graph.add_edge(u, v)
    return graph


def homology_graph(seq_list, background, k, internal_repeats, seq_file=None):
    """hgraph.get_homology_graph (hgraph.py:202-249): returns (graph, processing-ordered [(id, SEQ)]).

    Consumes `seq_list` like the reference does, and writes the 'id,SEQ' side file when asked to.
    """
    parts = uniquify(seq_list)
    del seq_list[:]
    if seq_file is not None:
        with open(seq_file, 'w') as out:
            for part_id, seq in parts:
                out.write('{},{}\n'.format(part_id, seq))
    table = repeat_table(parts, background, k, internal_repeats)
    graph = assemble_graph(clique_stream(table))
    return graph, parts


def ordered_stream(seq_list, background, k, internal_repeats):
    """The order-exact interface between the hot path and the host replay:
    distinct id-tuples in the order the reference first adds them to its clique set."""
    parts = uniquify(list(seq_list))
    table = repeat_table(parts, background, k, internal_repeats)
    seen = set()
    stream = []
    while table:
        _, ids = table.popitem()
        tup = tuple(ids)
        if tup not in seen:
            seen.add(tup)
            stream.append(tup)
    return stream
