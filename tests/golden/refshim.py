"""Import the UNMODIFIED reference (ayaanhossain/nrpcalc) in the dev container.

Test/fixture infrastructure only -- never imported by the product package.  The reference tree
(`/root/reference`) exists only in the development container, not on the GPU box, so everything that
needs it lives in `make_golden.py` (fixture generation) and in tests marked `needs_reference`.

The reference imports three third-party modules that are not installed here (Bio, RNA, ShareDB;
reference nrpcalc/base/utils.py:4,6 and nrpcalc/base/kmerSetDB.py:8).  None of them is reachable from
Finder mode except ShareDB's exact-key get/put/delete, which a dict reproduces.
"""
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _find_reference():
    """The reference tree: $NRPCALC_REFERENCE, the offline install under baseline/_ref (pip install --target, travels to the GPU
    box with the repository snapshot), then the development container's /root/reference."""
    for root in (os.environ.get('NRPCALC_REFERENCE'), os.path.join(_REPO, 'baseline', '_ref'), '/root/reference'):
        if root and os.path.isfile(os.path.join(root, 'nrpcalc', 'base', 'hgraph.py')):
            return root
    return os.environ.get('NRPCALC_REFERENCE', '/root/reference')


REFERENCE_ROOT = _find_reference()


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'nrpcalc', 'base', 'hgraph.py'))


class _DictShareDB(dict):
    """Stand-in for ShareDB(path=, map_size=): missing keys raise KeyError (kmerSetDB.py:54-65)."""

    def __init__(self, path=None, map_size=None):
        super().__init__()
        self.path = path
        if path is not None:
            os.makedirs(path, exist_ok=True)   # kmerSetDB.drop() rmtree's the parent directory

    def sync(self):
        pass

    def close(self):
        pass

    def remove(self, key):
        self.pop(key, None)


def load_reference():
    """Return the reference's top-level `nrpcalc` module (v1.7.6), importing it on first use."""
    if 'nrpcalc' in sys.modules and getattr(sys.modules['nrpcalc'], '_is_reference_shim', False):
        return sys.modules['nrpcalc']
    if not reference_available():
        raise RuntimeError('reference tree not found at %s' % REFERENCE_ROOT)
    bio = types.ModuleType('Bio')
    bio.SeqIO = types.ModuleType('Bio.SeqIO')
    sequtils = types.ModuleType('Bio.SeqUtils')
    sequtils.MeltingTemp = types.ModuleType('Bio.SeqUtils.MeltingTemp')
    bio.SeqUtils = sequtils
    sharedb = types.ModuleType('ShareDB')
    sharedb.ShareDB = _DictShareDB
    sys.modules.update({
        'Bio': bio, 'Bio.SeqIO': bio.SeqIO, 'Bio.SeqUtils': sequtils,
        'Bio.SeqUtils.MeltingTemp': sequtils.MeltingTemp,
        'RNA': types.ModuleType('RNA'), 'ShareDB': sharedb})
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import nrpcalc
    finally:
        sys.path.remove(REFERENCE_ROOT)
    nrpcalc._is_reference_shim = True
    return nrpcalc
