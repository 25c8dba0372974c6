"""Independent-set recovery around a vertex-cover routine -- host step, behaviour identical to the
reference (nrpcalc/base/recovery.py:32-96).

The cover routines destroy the graph they are given, so the full graph is dumped once as an adjacency list
and reloaded after every round (recovery.py:26-30,55,78).  The reload changes adjacency order (file order
instead of insertion order), which later rounds observe; the dump/reload is therefore part of result parity
and is kept.  The RELOADED graph, however, is never modified -- later rounds only read it and build their
working graph with nx.Graph(reloaded.subgraph(...)) (recovery.py:71-72,86) -- so every reload after the first
would parse the same file into an identical object: it is parsed once and reused (a third of the recovery time
at 100k parts).  When the graph came from the array tail (hgraph.assemble_graph keeps its node / adjacency arrays)
the reloaded graph is not parsed at all: what write_adjlist + read_adjlist do to the order (each edge listed once,
under its earlier endpoint; nodes and edges added line by line) is replayed on the arrays
(native_tail.reload_arrays, checked against networkx in tests/test_host_logic.py).  The file is still written.
"""
import os
import sys
from time import time

import networkx as nx

from . import native_tail
from . import vercov


def _write_adjlist_from_arrays(arrays, graph_file, adj_index=None):
    """What nx.write_adjlist(graph, graph_file) writes (header of three comment lines, then one line per node), from the arrays."""
    import time as _time
    nodes, indptr, adj = arrays
    header = '#' + ' '.join(sys.argv) + '\n' + '# GMT {}\n'.format(_time.asctime(_time.gmtime())) + '# \n'
    body = native_tail.adjlist_body(nodes, indptr, native_tail.adjacency_indices(nodes, adj) if adj_index is None else adj_index)
    with open(graph_file, 'wb') as out:
        out.write(header.encode('utf-8'))
        out.write(body)


def recover_from_arrays(arrays, graph_file, vercov_func):
    """The whole recovery on the arrays of the assembled graph (libnrptail.so: cover routines, candidate sets and working graphs of
    every round modelled exactly -- host_cover.cpp); the adjacency-list side file is still written.  'nrp2' / 'nrpG' only."""
    nodes = arrays[0]
    if nodes.size == 0:
        return set()
    adj0 = native_tail.adjacency_indices(arrays[0], arrays[2])
    _write_adjlist_from_arrays(arrays, graph_file, adj0)
    mode = vercov_func if vercov_func in native_tail.COVER_MODES else 'nrp2'       # unknown names fall through to 'nrp2' (recovery.py:18-24)
    chosen, _, _ = native_tail.recover_arrays(arrays[0], arrays[1], arrays[2], mode, adj0=adj0)
    return chosen


def native_recovery_applies(arrays, vercov_func, verbose):
    return (arrays is not None and not verbose and vercov_func != '2apx' and not os.environ.get('NRPCALC_B200_PYTHON_RECOVERY')
            and native_tail.usable() and native_tail.cover_usable())


def get_recovered_non_homologs(homology_graph, graph_file, vercov_func=None, verbose=True):
    arrays = getattr(homology_graph, '_nrp_arrays', None)
    if (arrays is not None and arrays[0].size == homology_graph.number_of_nodes() and arrays[2].size == 2 * homology_graph.number_of_edges()
            and native_recovery_applies(arrays, vercov_func, verbose)):
        return recover_from_arrays(arrays, graph_file, vercov_func)
    return _recover_on_graphs(homology_graph, graph_file, vercov_func, verbose)


def _recover_on_graphs(homology_graph, graph_file, vercov_func=None, verbose=True):
    independent = set()
    candidates = set(homology_graph.nodes())
    arrays = getattr(homology_graph, '_nrp_arrays', None)             # taken now: the cover routine empties the graph
    if arrays is not None and not (arrays[0].size == homology_graph.number_of_nodes() and
                                   arrays[2].size == 2 * homology_graph.number_of_edges() and native_tail.usable()):
        arrays = None
    if verbose:
        print('\n [+] Initial independent set = {}, computing vertex cover on remaining {} nodes.'.format(
            len(independent), 0, len(candidates)))
    if not homology_graph.number_of_nodes():
        if verbose:
            print(' [X] Graph is empty, further independent set expansion not possible, terminating.')
        return independent

    # unknown names fall through to 'nrp2' exactly like recovery.py:18-24
    routine, routine_name = vercov.ROUTINES.get(vercov_func, vercov.ROUTINES['nrp2'])
    if verbose:
        print(' [+] Vertex Cover Function: {}'.format(routine_name))
        sys.stdout.write(' [+] Dumping graph into: {}'.format(graph_file))
    t0 = time()
    nx.write_adjlist(homology_graph, graph_file)
    if verbose:
        sys.stdout.write(' in {} seconds\n'.format(time() - t0))

    iteration = -1
    reloaded = None
    while True:
        iteration += 1
        if verbose:
            print('\n----------------------')
            print('Now running iteration: {}'.format(iteration))
            print('----------------------')
        t0 = time()
        if iteration > 0:
            homology_graph = nx.Graph(homology_graph.subgraph(candidates))
        cover = routine(homology_graph, verbose)
        if verbose:
            print('\n [+] Computed vertex cover of size: {} (in {:.4} seconds)'.format(len(cover), time() - t0))
            print(' [+] Loading graph from: {}'.format(graph_file))
        if reloaded is None:
            if arrays is not None:
                try:
                    reloaded = native_tail.graph_from_arrays(*native_tail.reload_arrays(*arrays))
                except MemoryError:
                    reloaded = None
            if reloaded is None:
                reloaded = nx.read_adjlist(graph_file, nodetype=int)
        homology_graph = reloaded

        fresh = candidates - cover
        independent |= fresh
        candidates = cover
        before = len(candidates)
        neighbours = homology_graph._adj if type(getattr(homology_graph, '_adj', None)) is dict else homology_graph
        for node in fresh:                       # the adjacency dict itself: same keys as graph[node], no view object per node
            candidates.difference_update(neighbours[node])
        after = len(candidates)
        if verbose:
            print(' [+] Current independent set size:  {}'.format(len(independent)))
            print(' [+] Potential nodes for expansion: {} (projected independent set size: {})'.format(
                len(candidates), len(independent) + len(candidates)))
        if len(candidates) == 0 or before == after:
            if verbose:
                print(' [X] Cannot expand independent set, terminating.')
            break
    return independent
