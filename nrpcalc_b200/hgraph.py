"""Repeat ("homology") graph construction -- the accelerated hot path and its drop-in boundary.

Boundary: get_homology_graph(seq_list, background, homology, internal_repeats, seq_file, verbose) has the
signature, side effects and result of reference nrpcalc/base/hgraph.py:202-249 -- it consumes seq_list,
writes the 'id,SEQ' side file in processing order and returns an nx.Graph whose node order and adjacency
order equal the reference's (the unchanged vertex-cover step observes both).

What runs where
  host    uniquify + marshalling (this file)                          <- hgraph.py:11-21, 218-223
  GPU     k-mer extraction, reverse complement, canonical min, background membership, grouping of equal
          k-mers, internal-repeat exclusion, first-insertion ranks   <- hgraph.py:23-94, utils.py:37-50,
          (libnrpb200.so through nrpcalc_b200._native)                    kmerSetDB.py:177-207
  host    replay of the distinct id tuples through the clique tail    <- hgraph.py:96-200

The tail runs (1) on compact arrays with a layout-exact model of CPython's sets when nrpcalc_b200/libnrptail.so is built and
its start-up self-check against the Python loops passed (csrc/host_replay.cpp, native_tail.py), else (2) as C on real CPython
containers when nrpcalc_b200/_host_tail.so is built (csrc/host_tail.c: the same set / dict / deque operations through the C
API, identical by construction), else (3) as the Python loops below; NRPCALC_B200_PYTHON_TAIL=1 forces (3).

The tail cannot move to the GPU without changing results: the order in which cliques reach the graph is the
iteration order of CPython sets of tuples (hgraph.py:131-155).  It is reproduced by feeding real sets the
same values in the same order.  Only tuples that differ from all earlier ones change a set, so the device
returns every shared k-mer group with its rank and, per part, the rank of its last unshared window
(SURVEY.md 8b); that is O(parts + shared k-mers) host work instead of O(windows).

There is no CPU fallback: without the CUDA library or a GPU the call raises.
"""
import os
import sys
from collections import defaultdict, deque
from operator import itemgetter
from time import time

import networkx as nx
import numpy as np

from . import _native
from . import encoding
from . import native_tail

try:                                         # optional C extension (host only, no CUDA): the same container operations
    from . import _host_tail
except ImportError:                          # not built: the Python loops below
    _host_tail = None
if os.environ.get('NRPCALC_B200_PYTHON_TAIL'):
    _host_tail = None

_contexts = {}


def _context(device=None):
    if device is None:
        import os
        device = int(os.environ.get('NRPCALC_B200_DEVICE', os.environ.get('LOCAL_RANK', '0')))
    if device not in _contexts:
        _contexts[device] = _native.Context(device)
    return _contexts[device]


def uniquify_seq_list(seq_list):
    """Unique upper-cased parts in PROCESSING order: [(id, SEQ)], id = first index, ordered by descending last
    occurrence (hgraph.py:11-21 followed by the reversal at 218-223).  Empties seq_list like the reference."""
    first_index = {}
    for idx in range(len(seq_list) - 1, -1, -1):
        first_index[seq_list[idx].upper()] = idx
    del seq_list[:]
    return [(idx, seq) for seq, idx in first_index.items()]


class RepeatStructure(object):
    """Everything the replay needs, accumulated over the length classes of one job."""

    def __init__(self, n_parts):
        self.flags = np.zeros(n_parts, dtype=np.uint8)
        self.rank = []          # int64 arrays: p * 2^32 + window
        self.start = []         # per entry: offset into members
        self.size = []
        self.members = []       # global processing ranks p
        self.n_members = 0
        self.records = 0
        self.timings = defaultdict(float)
        self.counts = defaultdict(int)
        self.edges = None       # (u, v) id arrays of the last build that asked for edges

    def add_entries(self, rank, indptr, members):
        self.rank.append(rank)
        self.start.append(indptr[:-1] + self.n_members)
        self.size.append(np.diff(indptr))
        self.members.append(members)
        self.n_members += members.size


def _world_size():
    try:
        import torch.distributed as dist
        return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    except ImportError:
        return 1


def _device_build(ctx, structure, ascii_buf, offsets, global_p, ids, k, internal_repeats, want_edges=False, allow_sharding=False,
                  device_ptr=None, generic=False, preset=None):
    """One build over parts that all have length >= k; global_p maps local index -> processing rank.
    With an initialised torch.distributed group of more than one rank the build is hash-sharded across the ranks
    (nrpcalc_b200.sharded); every rank receives the same result."""
    if allow_sharding and not generic and _world_size() > 1:
        from . import sharded
        ops = getattr(ctx, '_sharded_ops', None)           # one per context: the peer mappings of the exchange live in it
        if ops is None:
            ops = ctx._sharded_ops = sharded.NativeOps(ctx)
        ops.load_shard(ascii_buf, offsets, ids if want_edges else None, k)      # every rank stages its own shard only
        out = sharded.build_sharded(ops, offsets, k, internal_repeats, want_edges=want_edges)
        flags, indptr, members = out['flags'], out['indptr'], out['members']
        first_window, last_single = out['first_window'], out['last_single']
        structure.records += out['stats']['records']
        structure.counts['groups'] += indptr.size - 1
        structure.counts['members'] += members.size
        structure.counts['runs'] = max(structure.counts['runs'], 1)
        structure.edges = None if out['edges'] is None else tuple(e.copy() for e in out['edges'])     # views of staging buffers
    else:
        ctx.load_parts(ascii_buf, offsets, ids if want_edges else None, device_ptr=device_ptr)
        if generic:          # any alphabet / k > 32: hashed window strings, equalities verified on the bytes (nrp_build_generic)
            structure.counts['generic_retries'] += ctx.build_generic(k, internal_repeats, want_edges=want_edges, preset_flags=preset)
            structure.counts['generic_builds'] += 1
        else:
            ctx.build(k, internal_repeats, want_edges=want_edges, early_fetch=True)      # fetched right below
        res = ctx.fetch(want_keys=False)
        flags, indptr, members, first_window, last_single = res.flags, res.indptr, res.members, res.first_window, res.last_single
        structure.records += res.counts['records']
        for name, value in res.timings.items():
            structure.timings[name] += value
        for name in ('groups', 'members', 'candidates', 'oversized', 'windows_max'):
            structure.counts[name] += res.counts[name]
        structure.counts['runs'] = max(structure.counts['runs'], res.counts['runs'])
        structure.edges = None if res.edges is None else (res.edges[0].copy(), res.edges[1].copy())     # views of library memory
    structure.flags[global_p] |= flags
    # shared k-mer groups: rank = (p of first member, first window); members -> processing ranks
    if indptr.size > 1:
        first = global_p[members[indptr[:-1]]].astype(np.int64)
        structure.add_entries((first << 32) | first_window.astype(np.int64), indptr, global_p[members])
    # singleton tuples: one per part that owns at least one window nobody shares
    own = np.nonzero(last_single >= 0)[0]
    if own.size:
        p = global_p[own].astype(np.int64)
        structure.add_entries((p << 32) | last_single[own].astype(np.int64), np.arange(own.size + 1, dtype=np.int64),
                              global_p[own])


def _install_background(ctx, background, longest_part):
    """Make the background's canonical K-mers the context's probe set (kmerSetDB.py:177-207 on the device)."""
    if background is None:
        ctx.set_background(None, 0)
        return
    if background.K <= longest_part and not background.ALIVE:
        raise RuntimeError('kmerSetDB was closed or dropped')      # surfaces from multicheck, kmerSetDB.py:97-102
    keys, _ = background.device_keys()
    if background.K > 32:                                          # keys wider than 64 bits: the host answers (preset flags)
        ctx.set_background(None, 0)
    else:
        ctx.set_background(keys, background.K)


_ACGTU = np.zeros(256, dtype=bool)
_ACGTU[np.frombuffer(b'ACGTU', dtype=np.uint8)] = True


def _host_background_flags(background, seqs, ascii_buf, offsets):
    """Background exclusions only the host can decide (kmerSetDB.py:177-207 on strings): every part when the background's keys
    are wider than 64 bits (K > 32), else the parts with characters other than A/C/G/T/U when the store holds such keys.  None
    when there is nothing to decide."""
    if background is None or not background.ALIVE or len(background) == 0:
        return None
    everything = background.K > 32
    if not everything and not background.has_verbatim_keys():
        return None
    lengths = np.diff(offsets)
    todo = lengths >= background.K
    if not everything:
        other = np.concatenate(([0], np.cumsum(~_ACGTU[ascii_buf])))
        todo &= (other[offsets[1:]] - other[offsets[:-1]]) > 0
    flags = np.zeros(len(seqs), dtype=np.uint8)
    for p in np.nonzero(todo)[0].tolist():
        if seqs[p] in background:
            flags[p] = _native.FLAG_BACKGROUND
    return flags


def repeat_structure_device(ctx, device_ptr, offsets, ids, longest_part, background, homology, internal_repeats, want_edges=False):
    """The device path over parts that already lie in device memory in processing order (the gathered buffer of
    Context.uniquify): every part at least `homology` long, A/C/G/T only.  Returns the RepeatStructure."""
    n = int(offsets.size - 1)
    structure = RepeatStructure(n)
    if homology > 32 and n:
        raise ValueError('Lmax > 31 needs keys wider than 64 bits, which the device path does not support')
    _install_background(ctx, background, longest_part)
    try:
        _device_build(ctx, structure, None, offsets, np.arange(n, dtype=np.int64), ids, homology, internal_repeats, want_edges,
                      allow_sharding=False, device_ptr=device_ptr)
    finally:
        ctx.set_background(None, 0)
    return structure


def repeat_structure(parts, background, homology, internal_repeats, ctx=None, want_edges=False):
    """Run the device path over the processing-ordered [(id, SEQ)] list.  Returns (RepeatStructure, ids)."""
    ctx = ctx or _context()
    n = len(parts)
    ids = np.fromiter(map(itemgetter(0), parts), dtype=np.uint32, count=n)
    seqs = list(map(itemgetter(1), parts))
    ascii_buf, offsets = encoding.join_parts(seqs)
    lengths = np.diff(offsets)
    structure = RepeatStructure(n)
    _install_background(ctx, background, int(lengths.max()) if n else 0)
    # The 2-bit kernels take A/C/G/T and k <= 32.  Everything else the reference accepts -- RNA (U's complement is A, utils.py:10),
    # other symbols, longer windows, backgrounds with K > 32 -- goes through the general path (csrc/generic.cuh): same results.
    wide_background = background is not None and background.ALIVE and background.K > 32 and len(background) > 0
    other_alphabet = not encoding.all_acgt(ascii_buf)
    preset = _host_background_flags(background, seqs, ascii_buf, offsets) if (other_alphabet or wide_background) else None

    def is_generic(k_class):
        return other_alphabet or wide_background or k_class > 32

    # parts with len >= k: ordinary k-windows.  Shorter parts are keyed by their whole string (utils.py:37-40):
    # one build per distinct length l < k with k' = l (a string of length l can only equal another of length l).
    all_p = np.arange(n, dtype=np.int64)
    long_mask = lengths >= homology
    try:
        if long_mask.all():
            _device_build(ctx, structure, ascii_buf, offsets, all_p, ids, homology, internal_repeats, want_edges, allow_sharding=True,
                          generic=is_generic(homology), preset=preset)
        else:
            classes = [(homology, np.nonzero(long_mask)[0])]
            for short_len in np.unique(lengths[~long_mask]).tolist():
                classes.append((int(short_len), np.nonzero(lengths == short_len)[0]))
            for k_class, idx in classes:
                if idx.size == 0:
                    continue
                if k_class == 0:
                    # the empty part: its only window '' equals its own reverse complement (hgraph.py:60-63)
                    if internal_repeats:
                        structure.add_entries(idx.astype(np.int64) << 32, np.arange(idx.size + 1, dtype=np.int64), idx)
                    else:
                        structure.flags[idx] |= _native.FLAG_PALINDROME
                    continue
                sub_off = np.zeros(idx.size + 1, dtype=np.int64)
                np.cumsum(lengths[idx], out=sub_off[1:])
                pieces = [ascii_buf[offsets[i]:offsets[i + 1]] for i in idx.tolist()]
                sub_ascii = np.concatenate(pieces) if pieces else np.zeros(0, dtype=np.uint8)
                _device_build(ctx, structure, sub_ascii, sub_off, idx, ids[idx], k_class, internal_repeats, False,
                              allow_sharding=k_class == homology, generic=is_generic(k_class),
                              preset=None if preset is None else preset[idx])
    finally:
        ctx.set_background(None, 0)
    return structure, ids


def ordered_tuples(structure, ids):
    """Distinct-tuple candidates in the order the reference meets them: keys leave repeat_dict by popitem(),
    i.e. by DESCENDING first-insertion rank (hgraph.py:132-135)."""
    if not structure.rank:
        return
    rank = np.concatenate(structure.rank)
    start = np.concatenate(structure.start)
    size = np.concatenate(structure.size)
    members = ids[np.concatenate(structure.members)].tolist()
    order = np.argsort(-rank, kind='stable')
    start = start[order].tolist()
    size = size[order].tolist()
    for s, z in zip(start, size):
        if z == 1:
            yield (members[s],)
        else:
            yield tuple(members[s:s + z])


def ordered_stream_arrays(structure, ids):
    """The same stream as ordered_tuples, as arrays: tuple t = members[start[t] : start[t] + size[t]] (ids, int64)."""
    if not structure.rank:
        empty = np.zeros(0, dtype=np.int64)
        return empty, empty, empty
    rank = np.concatenate(structure.rank)
    order = np.argsort(-rank, kind='stable')
    start = np.concatenate(structure.start)[order].astype(np.int64)
    size = np.concatenate(structure.size)[order].astype(np.int64)
    members = ids[np.concatenate(structure.members)].astype(np.int64)
    return start, size, members


def assemble_graph(structure, ids, verbose):
    """Ordered stream -> clique replay -> graph (hgraph.py:96-200).  On compact arrays with a layout-exact model of CPython's
    sets when libnrptail.so is built and its self-check against the Python loops passed (native_tail.py); through real
    CPython containers otherwise."""
    arrays = None
    if native_tail.usable():
        try:
            arrays = native_tail.run(*ordered_stream_arrays(structure, ids), want_cliques=True)
        except MemoryError:                                           # the library ran out of memory: the Python loops
            arrays = None
    if arrays is not None:
        nodes, indptr, adj, clique_off, _ = arrays
        graph = native_tail.graph_from_arrays(nodes, indptr, adj)
        graph._nrp_arrays = (nodes, indptr, adj)                      # recovery derives the dumped-and-reloaded graph from these
        if verbose:                                                   # what the per-clique counter leaves on the terminal
            print()
            sys.stdout.write(' [Cliques inserted] = {}\r'.format(clique_off.size - 1))
            sys.stdout.flush()
            print()
        return graph
    return build_homology_graph(replay_cliques(ordered_tuples(structure, ids)), verbose)


def _maximal_cliques(clique_sets):
    """Keep cliques that are not contained in another one (hgraph.py:96-115), same set operations in the same order."""
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


def replay_cliques(tuples):
    """hgraph.py:130-155 on the compressed stream.  Re-adding a tuple that is already present never changes a
    CPython set, so adding only the (superset of) first occurrences leaves every set exactly as in the reference."""
    if _host_tail is not None:
        return iter(_host_tail.replay_cliques(tuples))
    return _replay_cliques_python(tuples)


def _replay_cliques_python(tuples):
    distinct = set()
    for members in tuples:
        distinct.add(members)
    as_sets = []
    while distinct:
        as_sets.append(set(distinct.pop()))
    as_sets = _maximal_cliques(as_sets)
    final = set()
    while as_sets:
        final.add(tuple(as_sets.pop()))
    while final:
        yield deque(final.pop())


def build_homology_graph(cliques, verbose):
    """Insert cliques into the graph (hgraph.py:157-200): node and adjacency insertion order are part of the result.

    The reference calls graph.add_node / graph.add_edge(u, v) for v in (rest of the clique - set(graph[u])).  The same
    insertions are made here straight into the graph's node and adjacency dicts (what add_node / add_edge do for an
    attribute-free nx.Graph: one shared empty data dict per edge), and the difference is taken against the adjacency dict
    itself -- set.difference walks the left operand's table exactly as `-` does against a set of the same keys -- which
    makes the stage ~2.6x faster with an identical graph (tests/test_host_logic.py)."""
    if _host_tail is None:
        return _build_homology_graph_python(cliques, verbose)
    graph = nx.Graph()
    adj, nodes = graph._adj, graph._node
    if not (type(adj) is dict and type(nodes) is dict):               # an unexpected networkx: the public API
        return _build_homology_graph_api(cliques, verbose)
    if verbose:
        print()
    count = _host_tail.fill_graph(adj, nodes, cliques)
    if verbose:                                                       # what the per-clique counter leaves on the terminal
        sys.stdout.write(' [Cliques inserted] = {}\r'.format(count))
        sys.stdout.flush()
        print()
    return graph


def _build_homology_graph_python(cliques, verbose):
    """build_homology_graph as Python loops over the graph's dicts."""
    graph = nx.Graph()
    adj, nodes = graph._adj, graph._node
    if not (type(adj) is dict and type(nodes) is dict):               # an unexpected networkx: the public API
        return _build_homology_graph_api(cliques, verbose)
    if verbose:
        print()
        clear_length = 0
    count = 0
    for clique in cliques:
        if len(clique) == 1:
            u = clique.popleft()
            if u not in nodes:
                adj[u] = {}
                nodes[u] = {}
        else:
            pending = set(clique)
            while clique:
                u = clique.popleft()
                pending.remove(u)
                if u not in nodes:
                    adj[u] = {}
                    nodes[u] = {}
                nbrs = adj[u]
                for v in pending.difference(nbrs):
                    if v not in nodes:
                        adj[v] = {}
                        nodes[v] = {}
                    data = {}
                    nbrs[v] = data
                    adj[v][u] = data
        count += 1
        if verbose:
            printed = ' [Cliques inserted] = {}\r'.format(count)
            sys.stdout.write(' ' * clear_length + '\r')
            sys.stdout.write(printed)
            clear_length = len(printed)
            sys.stdout.flush()
    if verbose:
        print()
    cache = getattr(graph, '__networkx_cache__', None)                # what add_node / add_edge would have invalidated
    if cache:
        cache.clear()
    return graph


def _build_homology_graph_api(cliques, verbose):
    """The same insertions through networkx's public API (hgraph.py:157-200 literally)."""
    graph = nx.Graph()
    if verbose:
        print()
        clear_length = 0
    count = 0
    for clique in cliques:
        if len(clique) == 1:
            graph.add_node(clique.popleft())
        else:
            pending = set(clique)
            while clique:
                u = clique.popleft()
                pending.remove(u)
                graph.add_node(u)
                for v in pending - set(graph[u]):
                    graph.add_edge(u, v)
        count += 1
        if verbose:
            printed = ' [Cliques inserted] = {}\r'.format(count)
            sys.stdout.write(' ' * clear_length + '\r')
            sys.stdout.write(printed)
            clear_length = len(printed)
            sys.stdout.flush()
    if verbose:
        print()
    return graph


last_build_info = {}


def _build_structure(seq_list, background, homology, internal_repeats, seq_file, verbose):
    """hgraph.py:209-233 up to the table: uniquify, side file, device build.  Returns (parts, structure, ids, seconds of the call)."""
    t0 = time()
    num_seqs = len(seq_list)
    parts = uniquify_seq_list(seq_list)
    unq_seqs = len(parts)
    if verbose:
        print('Extracted {} unique sequences out of {} sequences in {:2.4} seconds'.format(unq_seqs, num_seqs, time() - t0))

    t0 = time()
    with open(seq_file, 'w') as outfile:
        outfile.write(''.join(['%d,%s\n' % part for part in parts]))
    if verbose:
        print('\nWritten {} unique sequences out to {} in {:2.4} seconds'.format(unq_seqs, seq_file, time() - t0))

    t0 = time()
    structure, ids = repeat_structure(parts, background, homology, internal_repeats)
    t_device = time() - t0
    if verbose:
        print(' [Sequence processing remaining] = 0  (GPU: {} k-mers in {:.3f} ms on device)'.format(
            structure.records, structure.timings['total']))
    return parts, structure, ids, t_device


def _record_build(parts, structure, t_device, t_replay, native):
    last_build_info.clear()
    last_build_info.update({'records': structure.records, 'device_ms': structure.timings['total'], 'call_s': t_device,
                            'timings': dict(structure.timings), 'counts': dict(structure.counts),
                            'replay_s': t_replay, 'parts': parts, 'native_tail': native})


def get_homology_graph(seq_list, background, homology, internal_repeats, seq_file, verbose):
    parts, structure, ids, t_device = _build_structure(seq_list, background, homology, internal_repeats, seq_file, verbose)
    t0 = time()
    graph = assemble_graph(structure, ids, verbose)
    _record_build(parts, structure, t_device, time() - t0, getattr(graph, '_nrp_arrays', None) is not None)
    if verbose:
        print('\nBuilt homology graph in {:2.4} seconds. [Edges = {}] [Nodes = {}]'.format(
            time() - t0 + t_device, graph.number_of_edges(), graph.number_of_nodes()))
        print(' [Intital Nodes = {}] - [Repetitive Nodes = {}] = [Final Nodes = {}]'.format(
            len(parts), len(parts) - graph.number_of_nodes(), graph.number_of_nodes()))
    return graph


def get_homology_arrays(seq_list, background, homology, internal_repeats, seq_file):
    """The same graph as get_homology_graph, as arrays only: (node ids in insertion order, CSR offsets, neighbour ids in insertion
    order) -- for callers that go on with the array recovery (recovery.recover_from_arrays) and never need the nx.Graph object,
    whose millions of small dicts cost more than the whole tail.  None when the array tail is not usable."""
    if not native_tail.usable():
        return None
    parts, structure, ids, t_device = _build_structure(seq_list, background, homology, internal_repeats, seq_file, False)
    t0 = time()
    try:
        nodes, indptr, adj = native_tail.run(*ordered_stream_arrays(structure, ids))
    except MemoryError:
        _record_build(parts, structure, t_device, time() - t0, False)
        return assemble_graph(structure, ids, False)               # the caller checks the type
    _record_build(parts, structure, t_device, time() - t0, True)
    return nodes, indptr, adj
