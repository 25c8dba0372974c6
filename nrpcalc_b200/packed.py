"""Finder mode over a PRE-PACKED part list -- the additive zero-copy entry.

    toolbox = nrpcalc_b200.finder_packed(buffer, Lmax, offsets=offsets)          # or part_len=100 for parts of one length

`buffer` holds the parts back to back as ASCII bytes (bytes, bytearray, memoryview or a uint8 numpy array -- e.g. a file read or
mapped in one piece), `offsets` their n + 1 byte offsets.  The result is what `finder(list_of_strings, ...)` returns for the same
parts under the same random.seed(): {index of the part's first occurrence: upper-cased part}, in processing order.

What is different from `finder`: no Python string is created for a part unless it is returned.  The steps the reference runs on
Python objects before the graph build -- upper-casing, dropping exact duplicates with id = first occurrence, processing in
descending order of last occurrence (reference nrpcalc/base/hgraph.py:11-21, 218-223) -- run on the GPU (nrp_uniquify:
hash, sort, byte-exact verification, gather), the gathered buffer feeds the build without leaving the device, and the side file
seq_list.txt (whose only reader is the reference's own output step, finder.py:175-181) is not written.  Inputs the device entry
does not take -- characters outside A/C/G/T, parts shorter than Lmax + 1, Lmax > 31, a job sharded over several ranks, or a hash
collision between different parts -- go through `finder` on strings built from the buffer: same result, host speed.
"""
import time
import uuid

import numpy as np

from . import finder as _finder
from . import hgraph
from . import native_tail
from . import projector
from . import recovery

# how the last job ran: {'path': 'device' | 'strings', 'unique': n, 'uniquify_s', 'graph_s', 'recovery_s', 'output_s'}
last_run_info = {}


def _as_bytes_array(buffer):
    if isinstance(buffer, np.ndarray):
        if buffer.dtype != np.uint8:
            raise TypeError('the part buffer must be a uint8 array, not {}'.format(buffer.dtype))
        return np.ascontiguousarray(buffer).reshape(-1)
    if isinstance(buffer, str):
        raise TypeError('the part buffer must hold bytes; encode the string or pass a list to finder()')
    return np.frombuffer(buffer, dtype=np.uint8)


def _offsets_of(buf, offsets, part_len):
    if (offsets is None) == (part_len is None):
        raise ValueError('pass exactly one of offsets= and part_len=')
    if offsets is None:
        part_len = int(part_len)
        if part_len < 1 or buf.size % part_len:
            raise ValueError('the buffer is not a whole number of parts of {} bytes'.format(part_len))
        return None, buf.size // part_len
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    if offsets.ndim != 1 or offsets.size < 1 or offsets[0] != 0 or (np.diff(offsets) < 0).any() or offsets[-1] > buf.size:
        raise ValueError('offsets must start at 0, never decrease and end inside the buffer')
    return offsets, offsets.size - 1


def _strings(buf, offsets, part_len, n):
    if offsets is None:
        blob = buf.tobytes().decode('latin-1')
        return [blob[i * part_len:(i + 1) * part_len] for i in range(n)]
    blob = buf[:int(offsets[-1])].tobytes().decode('latin-1')
    bounds = offsets.tolist()
    return [blob[bounds[i]:bounds[i + 1]] for i in range(n)]


def finder_packed(buffer, Lmax, offsets=None, part_len=None, internal_repeats=False, background=None, vercov='nrp2', output_file=None,
                  verbose=False):
    """Non-repetitive subset of the packed parts; see the module docstring.  Arguments after `part_len` as in finder()
    (reference nrpcalc/nrpcalc.py:226-391)."""
    buf = _as_bytes_array(buffer)
    offsets, n = _offsets_of(buf, offsets, part_len)
    homology = Lmax + 1 if isinstance(Lmax, int) else Lmax
    _finder._validate((), homology, internal_repeats, background, vercov, output_file, verbose)
    last_run_info.clear()

    def on_strings(reason):
        last_run_info.update({'path': 'strings', 'reason': reason})
        return _finder.nrp_finder(_strings(buf, offsets, part_len, n), homology, allow_internal_repeat=internal_repeats, background=background,
                                  vercov_func=vercov, output_file=output_file, verbose=verbose, check_constraints=False)

    if homology > 32 or hgraph._world_size() > 1 or not native_tail.usable():
        return on_strings('Lmax > 31, a sharded job or no array tail')
    ctx = hgraph._context()
    t0 = time.perf_counter()
    ids, source, new_off, dev, status, longest = ctx.uniquify(buf, offsets, part_len or 0)
    t1 = time.perf_counter()
    if status & 2:
        return on_strings('hash collision between different parts')
    if status & 1:
        return on_strings('characters outside A/C/G/T')
    if ids.size == 0:
        return on_strings('no parts')
    if int(np.diff(new_off).min()) < homology:
        return on_strings('parts shorter than Lmax + 1')

    with projector.ScratchDir(str(uuid.uuid4())) as root:
        graph_file = '{}/repeat_graph.txt'.format(root)
        structure = hgraph.repeat_structure_device(ctx, dev, new_off, ids, longest, background, homology, internal_repeats)
        arrays = native_tail.run(*hgraph.ordered_stream_arrays(structure, ids))
        t2 = time.perf_counter()
        if recovery.native_recovery_applies(arrays, vercov, verbose):
            chosen = recovery.recover_from_arrays(arrays, graph_file, vercov)
        else:
            graph = native_tail.graph_from_arrays(*arrays)
            graph._nrp_arrays = arrays
            chosen = recovery.get_recovered_non_homologs(graph, graph_file, vercov, verbose)
        t3 = time.perf_counter()
        keep = np.nonzero(np.isin(ids, np.fromiter(chosen, dtype=np.int64, count=len(chosen))))[0] if chosen else np.zeros(0, dtype=np.int64)
        parts_dict = {}
        starts = source[keep].astype(np.int64) * part_len if offsets is None else offsets[source[keep]]
        sizes = np.diff(new_off)[keep]
        for q, a, z in zip(keep.tolist(), starts.tolist(), sizes.tolist()):
            parts_dict[int(ids[q])] = buf[a:a + z].tobytes().decode('ascii').upper()
        if output_file:
            _finder._write_fasta(output_file, parts_dict)
    last_run_info.update({'path': 'device', 'unique': int(ids.size), 'parts': int(n), 'uniquify_s': t1 - t0, 'graph_s': t2 - t1,
                          'recovery_s': t3 - t2, 'output_s': time.perf_counter() - t3, 'device_ms': structure.timings['total']})
    if verbose:
        print('\nNon-Repetitive Toolbox Size: {}'.format(len(parts_dict)))
    return parts_dict
