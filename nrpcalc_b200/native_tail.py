"""ctypes binding of libnrptail.so (csrc/host_replay.cpp + cpyset.hpp): the order-exact clique replay and graph assembly
of reference nrpcalc/base/hgraph.py:96-200 on compact arrays.

The C++ side models CPython's set layout (probing, resizing, dummies, the bulk operations' heuristics, the tuple hash)
instead of creating millions of small Python objects.  That is only valid while the running interpreter follows the same
rules, so the first use runs a self-check: a randomised stream goes through this path AND through the Python loops of
hgraph.py (real sets); any difference disables the native path for the process (hgraph falls back to the Python loops).
Host code only -- no CUDA, no GPU needed."""
import ctypes
import gc
import os
import random
import sys

import networkx as nx
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libnrptail.so')
_lib = None
_state = {'checked': False, 'usable': False, 'why': 'not loaded'}


def _load():
    global _lib
    if _lib is None and os.path.exists(LIB_PATH) and not os.environ.get('NRPCALC_B200_PYTHON_TAIL'):
        lib = ctypes.CDLL(LIB_PATH)
        lib.nrp_tail_run.restype = ctypes.c_void_p
        lib.nrp_tail_run.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
        lib.nrp_tail_reload.restype = ctypes.c_void_p
        lib.nrp_tail_reload.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
        lib.nrp_tail_sizes.restype = None
        lib.nrp_tail_sizes.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        lib.nrp_tail_fetch.restype = None
        lib.nrp_tail_fetch.argtypes = [ctypes.c_void_p] + [ctypes.c_void_p] * 5
        lib.nrp_tail_free.restype = None
        lib.nrp_tail_free.argtypes = [ctypes.c_void_p]
        lib.nrp_tail_tuple_hash.restype = ctypes.c_int64
        lib.nrp_tail_tuple_hash.argtypes = [ctypes.c_void_p, ctypes.c_int64]
        lib.nrp_tail_set_script.restype = ctypes.c_int64
        lib.nrp_tail_set_script.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64]
        if hasattr(lib, 'nrp_tail_cover'):
            lib.nrp_tail_cover.restype = ctypes.c_int
            lib.nrp_tail_cover.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
            lib.nrp_tail_recover.restype = ctypes.c_int64
            lib.nrp_tail_recover.argtypes = [ctypes.c_void_p] * 6 + [ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
            lib.nrp_tail_adjlist.restype = ctypes.c_int64
            lib.nrp_tail_adjlist.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64]
        _lib = lib
    return _lib


def run(start, size, members, want_cliques=False):
    """The tail on arrays: tuple t of the ordered stream = members[start[t] : start[t] + size[t]].
    Returns (nodes, indptr, adjacency) -- node ids in insertion order and each node's neighbours in insertion order --
    plus (clique_off, clique_nodes) when want_cliques."""
    lib = _load()
    start = np.ascontiguousarray(start, dtype=np.int64)
    size = np.ascontiguousarray(size, dtype=np.int64)
    members = np.ascontiguousarray(members, dtype=np.int64)
    handle = lib.nrp_tail_run(start.ctypes.data, size.ctypes.data, start.size, members.ctypes.data)
    return _collect(lib, handle, want_cliques)


def _collect(lib, handle, want_cliques=False):
    if not handle:
        raise MemoryError('libnrptail: out of memory')
    try:
        sizes = np.zeros(4, dtype=np.int64)
        lib.nrp_tail_sizes(handle, sizes.ctypes.data)
        nodes = np.empty(int(sizes[0]), dtype=np.int64)
        indptr = np.empty(int(sizes[0]) + 1, dtype=np.int64)
        adj = np.empty(int(sizes[1]), dtype=np.int64)
        c_off = np.empty(int(sizes[2]) + 1, dtype=np.int64) if want_cliques else None
        c_nodes = np.empty(int(sizes[3]), dtype=np.int64) if want_cliques else None
        lib.nrp_tail_fetch(handle, nodes.ctypes.data, indptr.ctypes.data, adj.ctypes.data,
                           c_off.ctypes.data if want_cliques else None, c_nodes.ctypes.data if want_cliques else None)
    finally:
        lib.nrp_tail_free(handle)
    return (nodes, indptr, adj, c_off, c_nodes) if want_cliques else (nodes, indptr, adj)


def reload_arrays(nodes, indptr, adj):
    """Arrays of the graph nx.read_adjlist(file written by nx.write_adjlist(G)) returns, from the arrays of G
    (reference recovery.py:26-30; the dump / reload changes node and adjacency order, which later rounds observe)."""
    lib = _load()
    nodes = np.ascontiguousarray(nodes, dtype=np.int64)
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    adj = np.ascontiguousarray(adj, dtype=np.int64)
    return _collect(lib, lib.nrp_tail_reload(nodes.ctypes.data, indptr.ctypes.data, adj.ctypes.data, nodes.size))


def adjacency_indices(nodes, adj):
    """Neighbour ids -> node indices (int32) for a graph given as (node ids in iteration order, flat adjacency of ids)."""
    nodes = np.asarray(nodes, dtype=np.int64)
    if nodes.size and int(nodes.max()) <= 16 * nodes.size + (1 << 20):      # ids are first indices in the caller's list: a flat table
        table = np.zeros(int(nodes.max()) + 1, dtype=np.int32)
        table[nodes] = np.arange(nodes.size, dtype=np.int32)
        return table[adj]
    order = np.argsort(nodes, kind='stable')
    return order[np.searchsorted(nodes[order], adj)].astype(np.int32)


COVER_MODES = {'nrp2': 0, 'nrpG': 1}


def cover_arrays(indptr, adj_index, keep, mode):
    """'nrp2' / 'nrpG' vertex cover (reference vercov.py:119-154) of the graph restricted to the nodes with keep != 0, on arrays.
    Consumes the module-global `random` generator exactly as the Python routines' random.shuffle calls would.
    Returns a uint8 mask of the cover."""
    lib = _load()
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    adj_index = np.ascontiguousarray(adj_index, dtype=np.int32)
    keep = np.ascontiguousarray(keep, dtype=np.uint8)
    version, words, gauss = random.getstate()
    state = np.array(words, dtype=np.uint32)
    cover = np.zeros(keep.size, dtype=np.uint8)
    code = lib.nrp_tail_cover(indptr.ctypes.data, adj_index.ctypes.data, keep.size, keep.ctypes.data, COVER_MODES[mode], state.ctypes.data,
                              cover.ctypes.data)
    if code == -1:
        raise MemoryError('libnrptail: out of memory')
    if code != 0:
        raise RuntimeError('libnrptail: cover failed ({})'.format(code))
    random.setstate((version, tuple(int(w) for w in state), gauss))
    return cover


def recover_arrays(nodes, indptr, adj, mode, adj0=None):
    """Independent-set recovery (reference recovery.py:32-96 around 'nrp2' / 'nrpG') of the graph given by its arrays (node ids in
    insertion order, CSR adjacency of ids in insertion order), entirely on arrays.  Consumes the global `random` generator like the
    Python routines.  Returns (set of independent node ids, rounds, reloaded arrays)."""
    lib = _load()
    nodes = np.ascontiguousarray(nodes, dtype=np.int64)
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    if adj0 is None:
        adj0 = adjacency_indices(nodes, adj)
    adj0 = np.ascontiguousarray(adj0, dtype=np.int32)
    nodes1, indptr1, adj1_ids = reload_arrays(nodes, indptr, adj)
    adj1 = adjacency_indices(nodes1, adj1_ids)
    version, words, gauss = random.getstate()
    state = np.array(words, dtype=np.uint32)
    chosen = np.zeros(nodes.size, dtype=np.uint8)
    rounds = lib.nrp_tail_recover(nodes.ctypes.data, indptr.ctypes.data, adj0.ctypes.data, nodes1.ctypes.data, indptr1.ctypes.data,
                                  adj1.ctypes.data, nodes.size, COVER_MODES[mode], state.ctypes.data, chosen.ctypes.data)
    if rounds == -1:
        raise MemoryError('libnrptail: out of memory')
    if rounds < 0:
        raise RuntimeError('libnrptail: recovery failed ({})'.format(rounds))
    random.setstate((version, tuple(int(w) for w in state), gauss))
    return set(nodes1[chosen != 0].tolist()), int(rounds), (nodes1, indptr1, adj1_ids)


def adjlist_body(nodes, indptr, adj_index):
    """The lines nx.write_adjlist writes for this graph (without the comment header), as bytes."""
    lib = _load()
    nodes = np.ascontiguousarray(nodes, dtype=np.int64)
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    adj_index = np.ascontiguousarray(adj_index, dtype=np.int32)
    cap = int(indptr[-1]) * 12 + nodes.size * 14 + 64
    out = np.empty(cap, dtype=np.uint8)
    n = lib.nrp_tail_adjlist(nodes.ctypes.data, indptr.ctypes.data, adj_index.ctypes.data, nodes.size, out.ctypes.data, cap)
    if n < 0:
        raise MemoryError('libnrptail: adjacency list does not fit')
    return out[:n].tobytes()


_cover_state = {'checked': False, 'usable': False, 'why': 'not checked'}


def _cover_self_check():
    """Random graphs through the native routines and through vercov.py (real networkx graphs, the real random module): the
    covers and the generator's state afterwards must be equal."""
    from . import vercov
    rng = random.Random(77)
    for n, m, mode in ((60, 90, 'nrp2'), (400, 700, 'nrpG'), (3000, 4200, 'nrp2'), (3000, 2500, 'nrpG')):
        graph = nx.Graph()
        order = list(range(n))
        rng.shuffle(order)
        for u in order:
            if rng.random() < 0.9:
                graph.add_node(u)
        for _ in range(m):
            u, v = rng.randrange(n), rng.randrange(n)
            if u != v:
                graph.add_edge(u, v)
        if mode == 'nrpG':                                 # a re-derived working graph, as recovery builds for later rounds
            graph = nx.Graph(graph.subgraph([u for i, u in enumerate(graph.nodes()) if i % 7]))
        nodes = np.fromiter(graph.nodes(), dtype=np.int64)
        indptr = np.zeros(nodes.size + 1, dtype=np.int64)
        np.cumsum([len(graph[u]) for u in graph.nodes()], out=indptr[1:])
        adj = np.fromiter((v for u in graph.nodes() for v in graph[u]), dtype=np.int64, count=int(indptr[-1]))
        keep = np.ones(nodes.size, dtype=np.uint8)
        if n == 400:
            keep[::5] = 0                                   # the cover of an induced subgraph
        saved = random.getstate()
        try:
            random.seed(1234 + n)
            mask = cover_arrays(indptr, adjacency_indices(nodes, adj), keep, mode)
            state_native = random.getstate()
            random.seed(1234 + n)
            graph.remove_nodes_from(nodes[keep == 0].tolist())     # (Graph.copy() would re-derive the adjacency order)
            expect = vercov.ROUTINES[mode][0](graph, False)
            state_python = random.getstate()
        finally:
            random.setstate(saved)
        if set(nodes[mask != 0].tolist()) != set(expect):
            return False, 'cover differs ({} nodes, {})'.format(n, mode)
        if state_native != state_python:
            return False, 'random state differs ({} nodes, {})'.format(n, mode)
    # the whole recovery (several rounds; working graphs re-derived from the candidate SET) against the graph-based loop
    from . import recovery
    import tempfile
    scratch = tempfile.mkdtemp(prefix='nrp_tail_check_')
    try:
        for n, m, mode in ((300, 420, 'nrp2'), (2500, 6000, 'nrp2'), (2500, 2600, 'nrpG'), (9000, 5200, 'nrp2')):
            graph = nx.Graph()
            ids = rng.sample(range(4 * n), n)
            for u in ids:
                graph.add_node(u)
            for _ in range(m):
                u, v = rng.choice(ids), rng.choice(ids)
                if u != v:
                    graph.add_edge(u, v)
            nodes = np.fromiter(graph.nodes(), dtype=np.int64)
            indptr = np.zeros(nodes.size + 1, dtype=np.int64)
            np.cumsum([len(graph[u]) for u in graph.nodes()], out=indptr[1:])
            adj = np.fromiter((v for u in graph.nodes() for v in graph[u]), dtype=np.int64, count=int(indptr[-1]))
            saved = random.getstate()
            try:
                random.seed(99 + n)
                got, _, _ = recover_arrays(nodes, indptr, adj, mode)
                state_native = random.getstate()
                random.seed(99 + n)
                expect = recovery._recover_on_graphs(graph, os.path.join(scratch, 'graph.txt'), mode, False)
                state_python = random.getstate()
            finally:
                random.setstate(saved)
            if got != expect:
                return False, 'recovery differs ({} nodes, {})'.format(n, mode)
            if state_native != state_python:
                return False, 'random state after recovery differs ({} nodes, {})'.format(n, mode)
    finally:
        import shutil
        shutil.rmtree(scratch, ignore_errors=True)
    return True, 'ok'


def cover_usable():
    """True when libnrptail.so has the cover routines and they agree with vercov.py on this interpreter.  Checked once per process;
    a passed check is remembered as a second line of the stamp file next to the library (same key as usable(): interpreter
    version, networkx version, a digest of the library) -- the check costs 0.6 s, more than a 100k-part job."""
    if not _cover_state['checked']:
        _cover_state['checked'] = True
        lib = _load()
        if lib is None or not hasattr(lib, 'nrp_tail_cover'):
            _cover_state['usable'], _cover_state['why'] = False, 'libnrptail.so without cover routines'
        else:
            stamp = LIB_PATH + '.ok'
            key = None
            try:
                key = _verdict_key()
                if os.path.exists(stamp) and open(stamp).read().split('\n')[:2] == [key, 'cover ok']:
                    _cover_state['usable'], _cover_state['why'] = True, 'ok (remembered)'
                    return True
            except OSError:
                key = None
            try:
                _cover_state['usable'], _cover_state['why'] = _cover_self_check()
            except Exception as exc:
                _cover_state['usable'], _cover_state['why'] = False, 'self-check raised {!r}'.format(exc)
            if not _cover_state['usable']:
                sys.stderr.write('nrpcalc_b200: native cover routines disabled ({}); the Python routines run\n'.format(_cover_state['why']))
            elif key is not None and usable():             # (the first line of the stamp is the array tail's own verdict)
                try:
                    with open(stamp, 'w') as fh:
                        fh.write(key + '\ncover ok')
                except OSError:
                    pass
    return _cover_state['usable']


def graph_from_arrays(nodes, indptr, adj):
    """An attribute-free nx.Graph with exactly this node order and these adjacency orders (one shared empty data dict per
    edge, as Graph.add_edge creates)."""
    graph = nx.Graph()
    node_dict, adj_dict = graph._node, graph._adj
    if not (type(node_dict) is dict and type(adj_dict) is dict):
        raise RuntimeError('unexpected networkx Graph internals')
    ids = nodes.tolist()
    bounds = indptr.tolist()
    flat = adj.tolist()
    collecting = gc.isenabled()
    gc.disable()               # ~a million acyclic dicts are born here: generational scans would double the time and free nothing
    try:
        for i, u in enumerate(ids):
            node_dict[u] = {}
            adj_dict[u] = dict.fromkeys(flat[bounds[i]:bounds[i + 1]])
        for u, nbrs in adj_dict.items():
            for v, data in nbrs.items():
                if data is None:
                    data = {}
                    nbrs[v] = data
                    adj_dict[v][u] = data
    finally:
        if collecting:
            gc.enable()
    cache = getattr(graph, '__networkx_cache__', None)
    if cache:
        cache.clear()
    return graph


def stream_arrays(tuples):
    """(start, size, members) of an iterable of id tuples (tests and the self-check; the product passes its arrays directly)."""
    tuples = list(tuples)
    size = np.fromiter((len(t) for t in tuples), dtype=np.int64, count=len(tuples))
    start = np.zeros(len(tuples), dtype=np.int64)
    if len(tuples) > 1:
        np.cumsum(size[:-1], out=start[1:])
    members = np.fromiter((x for t in tuples for x in t), dtype=np.int64, count=int(size.sum()))
    return start, size, members


def _self_check():
    """A randomised stream with heavy sharing through both implementations (the last case takes the set of distinct tuples
    past 50000 entries, where CPython's growth rule changes)."""
    from . import hgraph
    rng = random.Random(20261020)
    for n_nodes, n_tuples, top in ((40, 300, 4), (3000, 9000, 6), (30000, 64000, 3)):
        tuples = [(rng.randrange(n_nodes),) for _ in range(n_tuples // 2)]
        for _ in range(n_tuples // 2):
            k = rng.randrange(2, top + 1)
            base = rng.randrange(n_nodes)
            tuples.append(tuple(min(n_nodes - 1, base + rng.randrange(0, 12)) if rng.random() < 0.7 else rng.randrange(n_nodes) for _ in range(k)))
        rng.shuffle(tuples)
        nodes, indptr, adj = run(*stream_arrays(tuples))
        expect = hgraph._build_homology_graph_python(hgraph._replay_cliques_python(iter(tuples)), False)
        if nodes.tolist() != list(expect.nodes()):
            return False, 'node order differs'
        bounds = indptr.tolist()
        flat = adj.tolist()
        for i, u in enumerate(expect.nodes()):
            if flat[bounds[i]:bounds[i + 1]] != list(expect[u]):
                return False, 'adjacency order differs'
    return True, 'ok'


def _verdict_key():
    """What a remembered verdict is valid for: this interpreter, this networkx, this library (by content: a copy of the tree to
    another machine with the same image keeps the verdict; size and time stamp would not)."""
    import hashlib
    with open(LIB_PATH, 'rb') as fh:
        digest = hashlib.sha1(fh.read()).hexdigest()
    return '{}|{}|{}'.format(sys.version, nx.__version__, digest)


def usable():
    """True when the library is built and its set model matches this interpreter.  Checked once per process; a passed check
    is remembered next to the library (keyed by interpreter version, networkx version and a digest of the library)."""
    if not _state['checked']:
        _state['checked'] = True
        if _load() is None:
            _state['usable'], _state['why'] = False, 'libnrptail.so not built (or NRPCALC_B200_PYTHON_TAIL set)'
            return False
        stamp = LIB_PATH + '.ok'
        try:
            key = _verdict_key()
            if os.path.exists(stamp) and open(stamp).read().split('\n')[0] == key:
                _state['usable'], _state['why'] = True, 'ok (remembered)'
                return True
        except OSError:
            key = None
        try:
            _state['usable'], _state['why'] = _self_check()
        except Exception as exc:                           # never let the accelerator break the product
            _state['usable'], _state['why'] = False, 'self-check raised {!r}'.format(exc)
        if not _state['usable']:                           # said once: the results stay identical, the tail is just ~3.6x slower
            sys.stderr.write('nrpcalc_b200: the array tail (libnrptail.so) does not match this interpreter ({}); '
                             'the Python loops run instead\n'.format(_state['why']))
        if _state['usable'] and key is not None:
            try:
                with open(stamp, 'w') as fh:
                    fh.write(key)
            except OSError:
                pass
    return _state['usable']


def status():
    usable()
    return dict(_state)
