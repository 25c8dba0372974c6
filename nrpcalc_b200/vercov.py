"""Vertex-cover heuristics of Finder mode -- host step, behaviour identical to the reference
(nrpcalc/base/vercov.py:9-154).

The hot path (graph construction) is what this package accelerates; the cover routines stay on the host.
They are order sensitive: which node a tie is broken towards depends on node-iteration order, adjacency
order and on the module-global `random` state, and the final toolbox inherits all of it.  Every container
operation below therefore happens in the same sequence as in the reference, so that the same graph and the
same random.seed() give the same cover.

  cover_2apx      <- nx_vercov            vercov.py:9-10     ('2apx')
  cover_nrp2      <- nrp_vercov_approx    vercov.py:119-138  ('nrp2', default)
  cover_greedy    <- nrp_vercov_greedy    vercov.py:140-154  ('nrpG')
"""
from collections import deque
from random import shuffle

import networkx.algorithms.approximation as nxa


def cover_2apx(graph, verbose):
    return nxa.vertex_cover.min_weighted_vertex_cover(graph)


def _top_degree(degree_pairs):
    """Nodes of maximum degree among (node, degree) pairs, in iteration order, then shuffled (vercov.py:12-20).
    The running maximum starts at 0, so isolated nodes tie with an empty list."""
    best, nodes = 0, []
    for node, degree in degree_pairs:
        if degree == best:
            nodes.append(node)
        elif degree > best:
            best, nodes = degree, [node]
    shuffle(nodes)
    return nodes


def _top_degree_nodes(graph, among=None):
    """graph.degree() / graph.degree(nbunch) yield (node, len(adj[node])) in node order resp. nbunch order for a graph without
    self-loops (the homology graph has none: hgraph.py:176-184 never pairs a node with itself); read off the adjacency dicts."""
    adj = getattr(graph, '_adj', None)
    if type(adj) is not dict:
        return _top_degree(graph.degree() if among is None else graph.degree(among))
    if among is None:
        return _top_degree((node, len(nbrs)) for node, nbrs in adj.items())
    return _top_degree((node, len(adj[node])) for node in among if node in adj)


def _low_degree_neighbours(graph, node):
    """Neighbours of `node` that become pendants or isolated once it is gone (vercov.py:28-34)."""
    found = [nb for nb in graph[node] if len(graph[nb]) <= 2]
    found.sort(key=lambda nb: len(graph[nb]) - 1)
    return deque(found)


def _drop_pendants(graph, queue, cover, verbose):
    """Pendant elimination (vercov.py:47-68): the only neighbour of a degree-1 node joins the cover."""
    while queue:
        node = queue.popleft()
        if node in graph:
            degree = len(graph[node])
            if degree == 1:
                if verbose:
                    print('  [x] Pendant node {} eliminated'.format(node))
                anchor = next(iter(graph[node]))
                graph.remove_edge(anchor, node)
                cover.add(anchor)
                exposed = [nb for nb in graph[anchor] if len(graph[nb]) <= 2]
                if exposed:
                    exposed.sort(key=lambda nb: len(graph[nb]) - 1)
                    queue.extendleft(exposed)
                graph.remove_nodes_from([node, anchor])
            elif degree == 0:
                # the reference removes isolated nodes only while printing (vercov.py:62-65); keep that
                if verbose:
                    print('  [x] Isolated node {} eliminated'.format(node))
                    graph.remove_node(node)
        elif verbose:
            print('  [+] Node {} covered'.format(node))


def _remove_with_pendants(graph, node, cover, verbose):
    if node in graph:
        queue = _low_degree_neighbours(graph, node)
        graph.remove_node(node)
        if queue:
            _drop_pendants(graph, queue, cover, verbose)


def _start_cover(graph, verbose):
    """Initial pendant sweep (vercov.py:70-89)."""
    cover = set()
    if verbose:
        print('\n Pendant checking is in progress...')
    queue = deque(sorted([node for node in graph if len(graph[node]) <= 1], key=lambda node: len(graph[node])))
    if not queue:
        if verbose:
            print('  [x] No pendants found')
    else:
        if verbose:
            print('  [+] {} Pendants found\n\n Pendant elimination initiated...'.format(len(queue)))
        _drop_pendants(graph, queue, cover, verbose)
        if verbose:
            print()
    return cover


def cover_nrp2(graph, verbose):
    """'nrp2': take a maximum-degree node and its maximum-degree neighbour (vercov.py:119-138).  Destroys graph."""
    cover = _start_cover(graph, verbose)
    while graph.number_of_edges():
        if verbose:
            print(' [Edges remaining = {}] [Nodes remaining = {}]'.format(graph.number_of_edges(), graph.number_of_nodes()))
        for node in _top_degree_nodes(graph):
            if node in graph:
                partner = _top_degree_nodes(graph, among=graph.neighbors(node))
                if partner:
                    partner = partner.pop()
                    cover.update([node, partner])
                    for member in (node, partner):
                        _remove_with_pendants(graph, member, cover, verbose)
                else:
                    graph.remove_node(node)
    return cover


def cover_greedy(graph, verbose):
    """'nrpG': take maximum-degree nodes only (vercov.py:140-154).  Destroys graph."""
    cover = _start_cover(graph, verbose)
    while graph.number_of_edges():
        if verbose:
            print(' [Edges remaining = {}] [Nodes remaining = {}]'.format(graph.number_of_edges(), graph.number_of_nodes()))
        for node in _top_degree_nodes(graph):
            if node in graph:
                cover.add(node)
                _remove_with_pendants(graph, node, cover, verbose)
    return cover


ROUTINES = {
    'nrpG': (cover_greedy, 'NRP Greedy'),
    '2apx': (cover_2apx, 'Standard 2-approximation'),
    'nrp2': (cover_nrp2, 'NRP 2-approximation'),
}
