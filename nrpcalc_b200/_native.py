"""ctypes binding of libnrpb200.so (C ABI declared in include/nrp_b200.h).

There is deliberately no fallback: if the CUDA library is missing or no GPU is visible the calls raise.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# NRPCALC_B200_LIB selects another build of the same library (A/B kernel variants, tools/build_variants.sh)
LIB_PATH = os.environ.get('NRPCALC_B200_LIB') or os.path.join(_HERE, 'libnrpb200.so')

N_COUNTS = 20
N_TIMINGS = 16
COUNT_NAMES = ('records', 'excluded', 'groups', 'members', 'edges', 'pairs', 'candidates', 'runs', 'buckets',
               'oversized', 'windows_max', 'key_bytes', 'launches', 'direct', 'fallbacks', 'packed', 'uniform_kernel', 'filter_warp')
TIMING_NAMES = ('total', 'h2d', 'flags', 'histogram', 'split1', 'split2', 'filter', 'sort', 'group', 'edges', 'order')

FLAG_REPEAT, FLAG_PALINDROME, FLAG_BACKGROUND = 1, 2, 4

E_ARG, E_CUDA, E_OOM, E_UNSUPPORTED, E_STATE, E_EDGES = -1, -2, -3, -4, -5, -6

_lib = None


class HostResult(ctypes.Structure):
    """nrp_host_result (include/nrp_b200.h): pointers into the library's page-locked result buffers."""
    _fields_ = [('flags', ctypes.c_void_p), ('indptr', ctypes.c_void_p), ('members', ctypes.c_void_p), ('first_window', ctypes.c_void_p),
                ('keys', ctypes.c_void_p), ('last_single', ctypes.c_void_p), ('edge_u', ctypes.c_void_p), ('edge_v', ctypes.c_void_p),
                ('n_parts', ctypes.c_int64), ('n_groups', ctypes.c_int64), ('n_members', ctypes.c_int64), ('n_edges', ctypes.c_int64)]


def _view(address, count, dtype):
    """numpy view (no copy) of `count` items of library-owned host memory"""
    if count == 0 or not address:
        return np.zeros(0, dtype=dtype)
    nbytes = int(count) * np.dtype(dtype).itemsize
    return np.frombuffer((ctypes.c_char * nbytes).from_address(address), dtype=dtype, count=int(count))


class NativeError(RuntimeError):
    def __init__(self, code, message):
        super().__init__('libnrpb200 error {}: {}'.format(code, message))
        self.code = code


def _declare(lib):
    c_ctx = ctypes.c_void_p
    i64, i32 = ctypes.c_int64, ctypes.c_int
    ptr = ctypes.c_void_p
    lib.nrp_version.restype = i32
    lib.nrp_last_error.restype = ctypes.c_char_p
    lib.nrp_launches_total.restype = i64
    lib.nrp_launches_total.argtypes = []
    sigs = {
        'nrp_ctx_create': (i32, ctypes.POINTER(c_ctx)),
        'nrp_ctx_destroy': (c_ctx,),
        'nrp_ctx_device': (c_ctx,),
        'nrp_ctx_set_option': (c_ctx, ctypes.c_char_p, i64),
        'nrp_bg_set': (c_ctx, ptr, i64, i32),
        'nrp_bg_clear': (c_ctx,),
        'nrp_bg_build': (c_ctx, ptr, ptr, i64, i32, i32, ctypes.POINTER(i64)),
        'nrp_bg_build_fetch': (c_ctx, ptr),
        'nrp_load_parts': (c_ctx, ptr, ptr, ptr, i64, i32),
        'nrp_load_parts_streamed': (c_ctx, ptr, ptr, ptr, i64, i32),
        'nrp_load_parts_uniform': (c_ctx, ptr, i64, i32, ptr, i32),
        'nrp_build': (c_ctx, i32, i32, i32),
        'nrp_build_generic': (c_ctx, i32, i32, i32, ctypes.c_uint64, ptr, ctypes.POINTER(i32)),
        'nrp_result_counts': (c_ctx, ptr),
        'nrp_result_timings': (c_ctx, ptr),
        'nrp_result_flags': (c_ctx, ptr),
        'nrp_result_groups': (c_ctx, ptr, ptr, ptr, ptr),
        'nrp_result_last_single': (c_ctx, ptr),
        'nrp_result_edges': (c_ctx, ptr, ptr),
        'nrp_ctx_set_stream': (c_ctx, ptr),
        'nrp_shard_extract': (c_ctx, i64, i64, i32, i32, i32, ptr),
        'nrp_shard_send_buffers': (c_ctx, ctypes.POINTER(ptr), ctypes.POINTER(ptr), ctypes.POINTER(i32)),
        'nrp_shard_group': (c_ctx, ptr, ptr, i64, i32, i32),
        'nrp_flags_export': (c_ctx, ptr),
        'nrp_flags_import': (c_ctx, ptr),
        'nrp_shard_finish': (c_ctx, i32, i32, i32),
        'nrp_shard_last_single': (c_ctx, ptr, i64, i32, i32),
        'nrp_edges_merge': (c_ctx, ptr, ptr, i64),
        'nrp_peer_export': (c_ctx, i64, i32, ptr, ptr),
        'nrp_peer_open': (c_ctx, i32, i32, ptr),
        'nrp_shard_count': (c_ctx, i64, i64, i32, i32, i32, ptr),
        'nrp_shard_scatter_peers': (c_ctx, i32, ptr),
        'nrp_shard_group_received': (c_ctx, i64, i32, i32),
        'nrp_result_device': (c_ctx, i32, ctypes.POINTER(ptr), ctypes.POINTER(i64)),
        'nrp_result_host': (c_ctx, i32, ctypes.POINTER(HostResult)),
        'nrp_peer_export_direct': (c_ctx, i64, i32, i32, ptr),
        'nrp_peer_open_direct': (c_ctx, i32, i32, ptr),
        'nrp_shard_scatter_direct': (c_ctx, i64, i64, i32, i32, i32, ctypes.POINTER(i32)),
        'nrp_shard_group_direct': (c_ctx, i32, i32),
        'nrp_shard_moved': (c_ctx, ptr),
        'nrp_peer_close': (c_ctx,),
        'nrp_flags_repeat_list': (c_ctx, ptr, i32),
        'nrp_flags_repeat_apply': (c_ctx, ptr, i32, i32),
        'nrp_load_parts_shard': (c_ctx, ptr, ptr, i32, ptr, i64, i64, i64, i32),
        'nrp_parts_capacity': (c_ctx, ctypes.POINTER(i64)),
        'nrp_peer_export_parts': (c_ctx, ptr),
        'nrp_peer_open_parts': (c_ctx, i32, i32, ptr, ptr),
        'nrp_uniquify': (c_ctx, ptr, ptr, i32, i64, ptr),
        'nrp_uniquify_fetch': (c_ctx, ptr, ptr, ptr, ctypes.POINTER(ptr)),
    }
    for name, argtypes in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = list(argtypes)
        fn.restype = i32
    return lib


def load_library():
    """Load libnrpb200.so (built in-tree by `make -C nrpcalc_b200/csrc` or __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(E_STATE, 'CUDA library not built: {} is missing (run __graft_entry__.build())'.format(LIB_PATH))
        _lib = _declare(ctypes.CDLL(LIB_PATH))
    return _lib


def _check(lib, code):
    if code != 0:
        raise NativeError(code, lib.nrp_last_error().decode('utf-8', 'replace'))


def _p(array):
    return None if array is None else ctypes.c_void_p(array.ctypes.data)


class BuildResult(object):
    """Host copies of one nrp_build() result."""

    def __init__(self, counts, timings, flags, indptr, members, first_window, keys, last_single, edges):
        self.counts = counts
        self.timings = timings
        self.flags = flags
        self.indptr = indptr
        self.members = members
        self.first_window = first_window
        self.keys = keys
        self.last_single = last_single
        self.edges = edges


class Context(object):
    """One libnrpb200 context = one CUDA device, one stream, reusable device buffers."""

    def __init__(self, device=0):
        self._lib = load_library()
        handle = ctypes.c_void_p()
        _check(self._lib, self._lib.nrp_ctx_create(int(device), ctypes.byref(handle)))
        self._h = handle
        self.device = int(device)
        self.n_parts = 0
        self._keep = None

    def close(self):
        if self._h is not None:
            self._lib.nrp_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name, value):
        _check(self._lib, self._lib.nrp_ctx_set_option(self._h, name.encode(), int(value)))

    def set_background(self, canon_kmers, K):
        if canon_kmers is None or len(canon_kmers) == 0:
            _check(self._lib, self._lib.nrp_bg_clear(self._h))
            return
        keys = np.ascontiguousarray(canon_kmers, dtype=np.uint64)
        _check(self._lib, self._lib.nrp_bg_set(self._h, _p(keys), keys.size, int(K)))

    def bg_build(self, ascii_buf, offsets, K, install=False):
        """Distinct canonical K-mers (ascending uint64) of every K-window of the sequences, built on the device."""
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        ascii_buf = np.ascontiguousarray(ascii_buf, dtype=np.uint8)
        n = ctypes.c_int64(0)
        _check(self._lib, self._lib.nrp_bg_build(self._h, _p(ascii_buf), _p(offsets), offsets.size - 1, int(K), 1 if install else 0, ctypes.byref(n)))
        keys = np.zeros(n.value, dtype=np.uint64)
        _check(self._lib, self._lib.nrp_bg_build_fetch(self._h, _p(keys)))
        return keys

    def uniquify(self, ascii_buf, offsets=None, part_len=0):
        """Device-side uniquify of a packed part list (nrp_uniquify): returns (ids, source, offsets, device address of the gathered
        upper-cased unique parts in processing order, status bits, longest part).  status & 1: some character is outside A/C/G/T;
        status & 2: a hash collision between different parts -- the result is unusable, take the host path."""
        ascii_buf = np.ascontiguousarray(ascii_buf, dtype=np.uint8)
        if offsets is not None:
            offsets = np.ascontiguousarray(offsets, dtype=np.int64)
            n_parts = offsets.size - 1
            assert ascii_buf.size >= int(offsets[-1])
        else:
            n_parts = ascii_buf.size // int(part_len) if part_len else 0
        out = np.zeros(4, dtype=np.int64)
        _check(self._lib, self._lib.nrp_uniquify(self._h, _p(ascii_buf) if ascii_buf.size else None, _p(offsets), int(part_len or 0), int(n_parts), _p(out)))
        n_unique, status, max_len = int(out[0]), int(out[2]), int(out[3])
        ids = np.zeros(n_unique, dtype=np.uint32)
        source = np.zeros(n_unique, dtype=np.uint32)
        new_off = np.zeros(n_unique + 1, dtype=np.int64)
        dev = ctypes.c_void_p()
        _check(self._lib, self._lib.nrp_uniquify_fetch(self._h, _p(ids), _p(source), _p(new_off), ctypes.byref(dev)))
        return ids, source, new_off, dev.value or 0, status, max_len

    def load_parts(self, ascii_buf, offsets, part_id=None, device_ptr=None, chunks=1):
        """ascii_buf: uint8 array (host) or None with device_ptr = address of the bytes in device memory.
        chunks > 1 (page-locked host buffer): streamed staging -- the copy is enqueued in pieces and the next build()
        extracts each piece as it lands; the context keeps a reference to ascii_buf until then."""
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n_parts = offsets.size - 1
        if part_id is not None:
            part_id = np.ascontiguousarray(part_id, dtype=np.uint32)
            assert part_id.size == n_parts
        if device_ptr is not None:
            src, on_device = ctypes.c_void_p(int(device_ptr)), 1
        else:
            ascii_buf = np.ascontiguousarray(ascii_buf, dtype=np.uint8)
            assert ascii_buf.size >= int(offsets[-1])
            src, on_device = _p(ascii_buf), 0
        if chunks > 1 and not on_device:
            self._keep = ascii_buf                    # borrowed until the next build has returned
            _check(self._lib, self._lib.nrp_load_parts_streamed(self._h, src, _p(offsets), _p(part_id), n_parts, int(chunks)))
        else:
            _check(self._lib, self._lib.nrp_load_parts(self._h, src, _p(offsets), _p(part_id), n_parts, on_device))
        self.n_parts = n_parts

    def load_parts_uniform(self, ascii_buf, n_parts, part_len, part_id=None, chunks=8):
        """Parts of one length in a page-locked uint8 host array: no offsets, copies only enqueued (nrp_load_parts_uniform);
        the context keeps references to the arrays until the next build has returned."""
        ascii_buf = np.ascontiguousarray(ascii_buf, dtype=np.uint8)
        assert ascii_buf.size >= int(n_parts) * int(part_len)
        if part_id is not None:
            part_id = np.ascontiguousarray(part_id, dtype=np.uint32)
            assert part_id.size == n_parts
        self._keep = (ascii_buf, part_id)
        _check(self._lib, self._lib.nrp_load_parts_uniform(self._h, _p(ascii_buf), int(n_parts), int(part_len), _p(part_id), int(chunks)))
        self.n_parts = int(n_parts)

    def load_parts_shard(self, ascii_buf, offsets, part_len, part_id, n_parts, stage_lo, stage_hi, chunks=4):
        """Shard-only staging (nrp_load_parts_shard): only the bytes of the parts [stage_lo, stage_hi) travel to the device;
        offsets=None for parts of one length part_len.  The context keeps references to the host arrays until the next build."""
        ascii_buf = np.ascontiguousarray(ascii_buf, dtype=np.uint8)
        if offsets is not None:
            offsets = np.ascontiguousarray(offsets, dtype=np.int64)
            assert offsets.size == n_parts + 1
        if part_id is not None:
            part_id = np.ascontiguousarray(part_id, dtype=np.uint32)
            assert part_id.size == n_parts
        self._keep = (ascii_buf, offsets, part_id)
        _check(self._lib, self._lib.nrp_load_parts_shard(self._h, _p(ascii_buf), _p(offsets), int(part_len or 0), _p(part_id), int(n_parts),
                                                         int(stage_lo), int(stage_hi), int(chunks)))
        self.n_parts = int(n_parts)

    def parts_capacity(self):
        out = ctypes.c_int64(0)
        _check(self._lib, self._lib.nrp_parts_capacity(self._h, ctypes.byref(out)))
        return out.value

    def peer_export_parts(self):
        handle = np.zeros(64, dtype=np.uint8)
        _check(self._lib, self._lib.nrp_peer_export_parts(self._h, _p(handle)))
        return handle

    def peer_open_parts(self, handles, my_rank, part_bounds):
        """handles: uint8 array of n_peers x 64 bytes; part_bounds: n_peers + 1 first part indices"""
        blob = np.ascontiguousarray(handles, dtype=np.uint8).reshape(-1)
        bounds = np.ascontiguousarray(part_bounds, dtype=np.int64)
        _check(self._lib, self._lib.nrp_peer_open_parts(self._h, bounds.size - 1, int(my_rank), _p(blob), _p(bounds)))

    def build(self, k, internal_repeats, want_edges=True, early_fetch=False):
        """nrp_build.  early_fetch=True (callers that go on to fetch() the result): whatever is final after the grouping stage
        starts its way to the page-locked result buffer beside the rest of the build (library option early_fetch)."""
        if bool(early_fetch) != getattr(self, '_early_fetch', False):
            self.set_option('early_fetch', 1 if early_fetch else 0)
            self._early_fetch = bool(early_fetch)
        _check(self._lib, self._lib.nrp_build(self._h, int(k), 1 if internal_repeats else 0, 1 if want_edges else 0))

    def build_generic(self, k, internal_repeats, want_edges=True, preset_flags=None, seed=0x9E3779B97F4A7C15, attempts=6, hash_bits=0):
        """nrp_build_generic: any byte alphabet, any k; windows keyed by a seeded hash, equalities verified on the bytes.  A
        reported collision between different strings repeats the build with the next seed.  Returns the number of retries.
        hash_bits > 0 keeps only that many bits of the hash (tests of the collision path)."""
        if preset_flags is not None:
            preset_flags = np.ascontiguousarray(preset_flags, dtype=np.uint8)
            assert preset_flags.size == self.n_parts
        status = ctypes.c_int(0)
        for attempt in range(attempts):
            _check(self._lib, self._lib.nrp_build_generic(self._h, int(k), 1 if internal_repeats else 0, 1 if want_edges else 0,
                                                          ctypes.c_uint64((seed & ((1 << 58) - 1)) | (int(hash_bits) << 58)), _p(preset_flags),
                                                          ctypes.byref(status)))
            if status.value == 0:
                return attempt
            seed = (seed * 0xD1342543DE82EF95 + 0x2545F4914F6CDD1D) & 0xFFFFFFFFFFFFFFFF
        raise NativeError(E_STATE, 'hash collisions under {} different seeds (status {})'.format(attempts, status.value))

    # ---- sharded path (device pointers are plain ints, e.g. torch.Tensor.data_ptr()) --------------------
    def set_stream(self, cuda_stream):
        _check(self._lib, self._lib.nrp_ctx_set_stream(self._h, ctypes.c_void_p(int(cuda_stream) if cuda_stream else None)))

    def shard_extract(self, part_lo, part_hi, k, internal_repeats, n_owners):
        counts = np.zeros(n_owners, dtype=np.int64)
        _check(self._lib, self._lib.nrp_shard_extract(self._h, int(part_lo), int(part_hi), int(k), 1 if internal_repeats else 0,
                                                      int(n_owners), _p(counts)))
        return counts

    def shard_send_buffers(self):
        keys, vals, kb = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int()
        _check(self._lib, self._lib.nrp_shard_send_buffers(self._h, ctypes.byref(keys), ctypes.byref(vals), ctypes.byref(kb)))
        return keys.value or 0, vals.value or 0, kb.value

    def shard_group(self, keys_ptr, vals_ptr, n_records, k, internal_repeats):
        _check(self._lib, self._lib.nrp_shard_group(self._h, ctypes.c_void_p(int(keys_ptr) or None), ctypes.c_void_p(int(vals_ptr) or None),
                                                    int(n_records), int(k), 1 if internal_repeats else 0))

    def peer_export(self, capacity_records, k):
        hk = np.zeros(64, dtype=np.uint8)
        hv = np.zeros(64, dtype=np.uint8)
        _check(self._lib, self._lib.nrp_peer_export(self._h, int(capacity_records), int(k), _p(hk), _p(hv)))
        return hk.tobytes() + hv.tobytes()

    def peer_open(self, handles, my_rank):
        blob = np.frombuffer(b''.join(handles), dtype=np.uint8).copy()
        _check(self._lib, self._lib.nrp_peer_open(self._h, len(handles), int(my_rank), _p(blob)))

    def peer_export_direct(self, windows_total, n_owners, k):
        handles = np.zeros(192, dtype=np.uint8)
        _check(self._lib, self._lib.nrp_peer_export_direct(self._h, int(windows_total), int(n_owners), int(k), _p(handles)))
        return handles.tobytes()

    def peer_open_direct(self, handles, my_rank):
        blob = np.frombuffer(b''.join(handles), dtype=np.uint8).copy()
        _check(self._lib, self._lib.nrp_peer_open_direct(self._h, len(handles), int(my_rank), _p(blob)))

    def peer_close(self):
        """Close every peer mapping (collective by contract: all ranks, then a barrier, before anyone re-exports)."""
        _check(self._lib, self._lib.nrp_peer_close(self._h))

    def shard_scatter_direct(self, part_lo, part_hi, k, internal_repeats, n_owners):
        overflow = ctypes.c_int(0)
        _check(self._lib, self._lib.nrp_shard_scatter_direct(self._h, int(part_lo), int(part_hi), int(k), 1 if internal_repeats else 0,
                                                             int(n_owners), ctypes.byref(overflow)))
        return overflow.value != 0

    def shard_group_direct(self, k, internal_repeats):
        _check(self._lib, self._lib.nrp_shard_group_direct(self._h, int(k), 1 if internal_repeats else 0))

    def shard_moved(self):
        out = np.zeros(2, dtype=np.int64)
        _check(self._lib, self._lib.nrp_shard_moved(self._h, _p(out)))
        return int(out[0]), int(out[1])

    def shard_count(self, part_lo, part_hi, k, internal_repeats, n_owners):
        counts = np.zeros(n_owners, dtype=np.int64)
        _check(self._lib, self._lib.nrp_shard_count(self._h, int(part_lo), int(part_hi), int(k), 1 if internal_repeats else 0,
                                                    int(n_owners), _p(counts)))
        return counts

    def shard_scatter_peers(self, k, write_offsets):
        offs = np.ascontiguousarray(write_offsets, dtype=np.int64)
        _check(self._lib, self._lib.nrp_shard_scatter_peers(self._h, int(k), _p(offs)))

    def shard_group_received(self, n_records, k, internal_repeats):
        _check(self._lib, self._lib.nrp_shard_group_received(self._h, int(n_records), int(k), 1 if internal_repeats else 0))

    def flags_export(self, dst_ptr):
        _check(self._lib, self._lib.nrp_flags_export(self._h, ctypes.c_void_p(int(dst_ptr))))

    def flags_import(self, src_ptr):
        _check(self._lib, self._lib.nrp_flags_import(self._h, ctypes.c_void_p(int(src_ptr))))

    def shard_finish(self, k, internal_repeats, want_edges):
        """True when the stage finished; False when a sparse flag list was truncated (combine the flags densely and call again)."""
        code = self._lib.nrp_shard_finish(self._h, int(k), 1 if internal_repeats else 0, 1 if want_edges else 0)
        if code == 1:                               # NRP_RETRY_DENSE_FLAGS
            return False
        _check(self._lib, code)
        return True

    def flags_repeat_list(self, list_ptr, cap):
        _check(self._lib, self._lib.nrp_flags_repeat_list(self._h, ctypes.c_void_p(int(list_ptr)), int(cap)))

    def flags_repeat_apply(self, lists_ptr, world, cap):
        _check(self._lib, self._lib.nrp_flags_repeat_apply(self._h, ctypes.c_void_p(int(lists_ptr)), int(world), int(cap)))

    def shard_last_single(self, shared_keys, k, device_ptr=None, n=0):
        """shared keys as a host uint64 array, or (device_ptr, n) for keys already in device memory"""
        if device_ptr is not None:
            _check(self._lib, self._lib.nrp_shard_last_single(self._h, ctypes.c_void_p(int(device_ptr) or None), int(n), int(k), 1))
            return
        keys = np.ascontiguousarray(shared_keys, dtype=np.uint64)
        _check(self._lib, self._lib.nrp_shard_last_single(self._h, _p(keys) if keys.size else None, keys.size, int(k), 0))

    def edges_merge(self, u_ptr, v_ptr, n):
        _check(self._lib, self._lib.nrp_edges_merge(self._h, ctypes.c_void_p(int(u_ptr) or None), ctypes.c_void_p(int(v_ptr) or None), int(n)))

    def result_device(self, what):
        ptr, count = ctypes.c_void_p(), ctypes.c_int64()
        _check(self._lib, self._lib.nrp_result_device(self._h, int(what), ctypes.byref(ptr), ctypes.byref(count)))
        return ptr.value or 0, count.value

    def counts(self):
        out = np.zeros(N_COUNTS, dtype=np.int64)
        _check(self._lib, self._lib.nrp_result_counts(self._h, _p(out)))
        return dict(zip(COUNT_NAMES, out.tolist()))

    def timings(self):
        out = np.zeros(N_TIMINGS, dtype=np.float64)
        _check(self._lib, self._lib.nrp_result_timings(self._h, _p(out)))
        return dict(zip(TIMING_NAMES, out.tolist()))

    def fetch(self, want_keys=True, copy=False):
        """The whole result of the last build on the host (one call, one synchronisation: nrp_result_host).
        The arrays are VIEWS of the library's page-locked result buffers, valid until the next load_parts / build on this
        context; copy=True returns private copies."""
        counts = self.counts()
        hr = HostResult()
        _check(self._lib, self._lib.nrp_result_host(self._h, 1 if want_keys else 0, ctypes.byref(hr)))
        n, ng, nm, ne = hr.n_parts, hr.n_groups, hr.n_members, hr.n_edges
        fix = (lambda a: a.copy()) if copy else (lambda a: a)
        flags = fix(_view(hr.flags, n, np.uint8))
        indptr = fix(_view(hr.indptr, ng + 1, np.int64))
        members = fix(_view(hr.members, nm, np.uint32))
        first_window = fix(_view(hr.first_window, ng, np.uint32))
        keys = fix(_view(hr.keys, ng, np.uint64)) if want_keys else None
        last_single = fix(_view(hr.last_single, n, np.int32))
        edges = None
        if counts['edges'] >= 0:
            edges = (fix(_view(hr.edge_u, ne, np.uint32)), fix(_view(hr.edge_v, ne, np.uint32)))
        return BuildResult(counts, self.timings(), flags, indptr, members, first_window, keys, last_single, edges)
