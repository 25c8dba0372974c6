"""B200-native Finder mode of the Non-Repetitive Parts Calculator.

Drop-in for the two entry points of ayaanhossain/nrpcalc that Finder mode needs:

    import nrpcalc_b200 as nrpcalc
    bkg = nrpcalc.background(path='./my_genome/', Lmax=15)
    bkg.multiadd(chromosomes)
    toolbox = nrpcalc.finder(seq_list, Lmax=15, internal_repeats=False, background=bkg, vercov='nrp2')

Same signatures, same return values (reference nrpcalc/nrpcalc.py:47-224 and 226-391); for the same input and
the same random.seed() the returned dictionary is identical to the reference's.  Graph construction runs on
the GPU through libnrpb200.so; Maker mode is out of scope.
"""
import sys

from . import finder as _finder
from . import kmerSetDB as _kmerSetDB
from .packed import finder_packed          # additive entry: packed part buffer in, uniquify on the GPU (packed.py)

__version__ = '0.1.0'


def background(path, Lmax, verbose=True, device=None):
    """kmerSetDB background object storing the canonical (Lmax+1)-mers of the sequences added to it
    (reference nrpcalc/nrpcalc.py:47-224).  device=True (additive keyword, or NRPCALC_B200_BG_DEVICE=1) builds the key
    set of add / multiadd on the GPU instead of in the storage layer's host code."""
    return _kmerSetDB.kmerSetDB(path=path, homology=Lmax + 1, verbose=verbose, device=device)


def finder(seq_list, Lmax, internal_repeats=False, background=None, vercov='nrp2', output_file=None, verbose=True):
    """Non-repetitive subset of seq_list: no two returned parts share a repeat longer than Lmax
    (reference nrpcalc/nrpcalc.py:226-391).  Returns {index in seq_list: part}."""
    return _finder.nrp_finder(seq_list=seq_list, homology=Lmax + 1, background=background,
                              allow_internal_repeat=internal_repeats, vercov_func=vercov, output_file=output_file,
                              verbose=verbose, check_constraints=True)


_patched = {}


def patch(reference=None):
    """Drop the GPU graph build into an installed reference: rebinds the ONE seam of the hot path,
    `nrpcalc.base.hgraph.get_homology_graph` (reference hgraph.py:202-249, called from finder.py:159-165), to this
    package's implementation, so that the reference's own `nrpcalc.finder`, recovery and vertex-cover code run unchanged on
    a graph built by libnrpb200.so.  `reference` is the imported reference package (default: the module named `nrpcalc`).
    A reference kmerSetDB passed as background is read through its public iterator (its K-mers are re-encoded once per
    job).  Returns the reference module; `unpatch()` restores the original function."""
    ref = reference if reference is not None else sys.modules.get('nrpcalc')
    if ref is None:
        import nrpcalc as ref                                   # raises ImportError when the reference is not installed
    from . import hgraph as _hgraph
    target = sys.modules.get(ref.__name__ + '.base.hgraph')
    if target is None:
        import importlib
        target = importlib.import_module(ref.__name__ + '.base.hgraph')
    if target not in _patched:
        _patched[target] = target.get_homology_graph

        def get_homology_graph(seq_list, background, homology, internal_repeats, seq_file, verbose):
            return _hgraph.get_homology_graph(seq_list, _kmerSetDB.adopt(background), homology, internal_repeats, seq_file, verbose)

        get_homology_graph.__doc__ = 'nrpcalc_b200 graph build behind the reference seam (hgraph.py:202-249)'
        target.get_homology_graph = get_homology_graph
    return ref


def unpatch():
    """Undo patch()."""
    for target, original in list(_patched.items()):
        target.get_homology_graph = original
        del _patched[target]
