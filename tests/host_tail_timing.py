"""Wall clock of the HOST side of nrpcalc_b200.finder on this machine's CPU, with the numpy oracle standing in for the device
call (tests/fake_device.py) -- the numbers of DESIGN.md section 6 "API level".  Test-side script (it imports the oracle), not
collected by pytest.

    python tests/host_tail_timing.py [n_parts] [Lmax]          # default 300000 14; NRPCALC_B200_PYTHON_TAIL=1 for the Python loops
"""
import os
import pickle
import random
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import cases                                     # noqa: E402
import fake_device                               # noqa: E402
import networkx as nx                            # noqa: E402
import nrpcalc_b200                              # noqa: E402
from nrpcalc_b200 import hgraph, native_tail, recovery, vercov   # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
    lmax = int(sys.argv[2]) if len(sys.argv) > 2 else 14
    parts = cases.random_parts(n, 100, 1)
    cache = os.path.join(tempfile.gettempdir(), 'nrp_struct_{}_{}.pkl'.format(n, lmax))
    spent = {}

    def timed(module, name, label=None):
        inner = getattr(module, name)

        def outer(*args, **kwargs):
            t0 = time.perf_counter()
            try:
                return inner(*args, **kwargs)
            finally:
                spent[label or name] = spent.get(label or name, 0.0) + time.perf_counter() - t0
        setattr(module, name, outer)

    def device_stand_in(uniq, background, homology, internal_repeats, ctx=None, want_edges=False):
        if os.path.exists(cache):
            return pickle.load(open(cache, 'rb'))
        result = fake_device.oracle_repeat_structure(uniq, background, homology, internal_repeats)
        pickle.dump(result, open(cache, 'wb'))
        return result

    hgraph.repeat_structure = device_stand_in
    for module, name in ((hgraph, 'repeat_structure'), (hgraph, 'uniquify_seq_list'), (hgraph, 'assemble_graph'),
                         (recovery, 'get_recovered_non_homologs'), (nx, 'write_adjlist'), (native_tail, 'graph_from_arrays'),
                         (native_tail, 'run'), (native_tail, 'reload_arrays'), (native_tail, 'recover_arrays'),
                         (recovery, 'recover_from_arrays'), (recovery, '_write_adjlist_from_arrays'), (native_tail, 'adjacency_indices')):
        timed(module, name)
    routine, label = vercov.ROUTINES['nrp2']

    def cover(*args, **kwargs):
        t0 = time.perf_counter()
        try:
            return routine(*args, **kwargs)
        finally:
            spent['cover_nrp2'] = spent.get('cover_nrp2', 0.0) + time.perf_counter() - t0
    vercov.ROUTINES['nrp2'] = (cover, label)
    print('native tail:', native_tail.status())
    os.chdir(tempfile.mkdtemp(prefix='nrp_tail_timing_'))
    random.seed(3)
    t0 = time.perf_counter()
    found = nrpcalc_b200.finder(parts, lmax, verbose=False)
    total = time.perf_counter() - t0
    stand_in = spent.pop('repeat_structure')
    print('finder(): {} parts, Lmax={} -> toolbox {}; host {:.2f} s (+ {:.2f} s for the device stand-in)'.format(
        n, lmax, len(found), total - stand_in, stand_in))
    for name, seconds in sorted(spent.items(), key=lambda kv: -kv[1]):
        print('  {:32s} {:6.2f} s'.format(name, seconds))


if __name__ == '__main__':
    main()
