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
from . import finder as _finder
from . import kmerSetDB as _kmerSetDB

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
