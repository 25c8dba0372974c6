"""Finder-mode job: validate the request, build the repeat graph on the GPU, recover an independent set, report.

Stand-alone counterpart of the reference's orchestrator (nrpcalc/base/finder.py:68-196) for installations that do not
have the reference package; when the reference IS installed, `nrpcalc_b200.patch()` rebinds only the graph-build seam
(hgraph.get_homology_graph) and the reference's own finder / recovery / vercov run unchanged.

Behaviour kept (what a caller can observe, stdout aside): the caller's list is not consumed; a request that fails
validation raises RuntimeError('Invalid Constraints, or Background'); a closed background is reported but only raises
later, from the membership test (reference finder.py:107-116 sets an unused name); the job keeps seq_list.txt and
repeat_graph.txt in ./<uuid>/ while it runs; the result maps index -> upper-cased part in processing order; the optional
FASTA file has '>index N' headers and 80-column lines (finder.py:175-189).
"""
import textwrap
import time
import uuid

from . import hgraph
from . import kmerSetDB
from . import projector
from . import recovery

VERTEX_COVERS = ('2apx', 'nrpG', 'nrp2')

# timings of the last job (seconds): graph_s (marshalling + device + replay), recovery_s, output_s
last_run_info = {}


def _problems(seq_list, homology, internal_repeats, background, vercov, output_file):
    """Yield (what is wrong, what to do) for every rule the request breaks, in the reference's order of checks
    (finder.py:15-66, 107-125)."""
    bad = next((seq for seq in seq_list if not isinstance(seq, str)), None)
    if bad is not None:
        yield 'every part must be a str, found {}'.format(type(bad).__name__), 'pass the parts as strings'
    if not isinstance(homology, int):
        yield 'Lmax must be an int, not {}'.format(type(homology).__name__), 'pass an integer Lmax'
    elif homology - 1 < 5:
        yield 'Lmax must be at least 5, not {}'.format(homology - 1), 'raise Lmax'
    if internal_repeats not in (True, False):
        yield 'internal_repeats must be a bool, not {}'.format(type(internal_repeats).__name__), 'pass True or False'
    if background is not None and not isinstance(background, kmerSetDB.kmerSetDB):
        yield 'background is not a kmerSetDB', 'create it with nrpcalc_b200.background(...)'
    if not isinstance(vercov, str) or vercov not in VERTEX_COVERS:
        yield 'vercov must be one of {}, not {!r}'.format(', '.join(VERTEX_COVERS), vercov), 'pick one of the three routines'
    if output_file is not None and not isinstance(output_file, str):
        yield 'output_file must be a str or None, not {}'.format(type(output_file).__name__), 'pass a path or None'


def _validate(seq_list, homology, internal_repeats, background, vercov, output_file, verbose):
    if verbose:
        print('\n[Checking Request]\n parts: {}   Lmax: {}   internal repeats: {}   background: {}   vertex cover: {}   output: {}'.format(
            len(seq_list), homology - 1 if isinstance(homology, int) else homology, internal_repeats, background, vercov, output_file))
    problem = next(_problems(seq_list, homology, internal_repeats, background, vercov, output_file), None)
    if problem is not None:
        print('\n [ERROR]    {}\n [SOLUTION] {}\n'.format(*problem))
        raise RuntimeError('Invalid Constraints, or Background')
    if isinstance(background, kmerSetDB.kmerSetDB) and not background.ALIVE:
        print('\n [WARNING]  the background is closed or dropped; the membership test will raise')
    if verbose:
        print(' request accepted\n')


def _write_fasta(path, parts_dict):
    with open(path, 'w') as out:
        for index, seq in parts_dict.items():
            out.write('>index {}\n{}\n'.format(index, '\n'.join(textwrap.wrap(seq, 80))))


def nrp_finder(seq_list, homology, allow_internal_repeat=False, background=None, vercov_func=None, output_file=None,
               verbose=True, check_constraints=True):
    if verbose:
        print('\n[Non-Repetitive Parts Calculator - Finder Mode, B200]')
    seq_list = list(seq_list)
    if check_constraints:
        _validate(seq_list, homology, allow_internal_repeat, background, vercov_func, output_file, verbose)

    scratch = projector.ScratchDir(str(uuid.uuid4()))
    with scratch:
        seq_file, graph_file = scratch.file('seq_list.txt'), scratch.file('repeat_graph.txt')
        t0 = time.perf_counter()
        graph = None
        if recovery.native_recovery_applies((), vercov_func, verbose):
            # quiet run with 'nrp2' / 'nrpG': graph and recovery stay on arrays (same result; no nx.Graph object is built)
            graph = hgraph.get_homology_arrays(seq_list, background, homology, allow_internal_repeat, seq_file)
        if isinstance(graph, tuple):
            t1 = time.perf_counter()
            chosen = recovery.recover_from_arrays(graph, graph_file, vercov_func)
        else:
            if graph is None:
                graph = hgraph.get_homology_graph(seq_list, background, homology, allow_internal_repeat, seq_file, verbose)
            t1 = time.perf_counter()
            chosen = recovery.get_recovered_non_homologs(graph, graph_file, vercov_func, verbose)
        t2 = time.perf_counter()
        # the unique parts in processing order: what seq_list.txt holds (the reference re-reads the file, finder.py:175-181)
        ordered = hgraph.last_build_info.get('parts')
        if ordered is None:
            with open(seq_file) as lines:
                ordered = [(int(index), seq) for index, seq in (line.rstrip('\n').split(',') for line in lines)]
        parts_dict = {index: seq for index, seq in ordered if index in chosen}
        if output_file:
            _write_fasta(output_file, parts_dict)
    last_run_info.clear()
    last_run_info.update({'graph_s': t1 - t0, 'recovery_s': t2 - t1, 'output_s': time.perf_counter() - t2})
    if verbose:
        print('\nNon-Repetitive Toolbox Size: {}'.format(len(parts_dict)))
    return parts_dict
