"""Background k-mer set -- drop-in for the reference's kmerSetDB (nrpcalc/base/kmerSetDB.py:12-330).

Same public surface (K, ALIVE, LEN, add, multiadd, __contains__, multicheck, __iter__, __len__, remove, multiremove,
clear, close, drop, repr) and the same semantics: sequences are U->T normalised, every canonical K-window becomes a
key, and a query is True iff ANY of its canonical K-windows is a key (kmerSetDB.py:177-192).

What changes is the storage and who answers the hot path's membership test:
  * the reference keeps one LMDB record per k-mer through the third-party ShareDB; here the keys over A/C/G/T are one
    sorted uint64 array in the 2-bit encoding (persisted as keys.npy under <path>/kmerSetDB.ShareDB/), which is exactly
    the form the GPU probe table is built from (`device_keys()` -> nrp_bg_set in include/nrp_b200.h);
  * windows are classified ONE BY ONE: a window made of upper-case A/C/G/T only goes to the 2-bit store whatever else the
    sequence contains (an 'N' or a soft-masked stretch spoils only the K windows that cover it); windows with other
    characters, and whole strings shorter than K (utils.py:37-40), are kept verbatim in a side set so that
    add / remove / len / iter behave like the reference.  The side set can never match a window of an A/C/G/T part.
  * inserts are buffered: add() is O(windows of the sequence); the sorted array is merged at the next read, and the store
    is written at multiadd / multiremove / close (and at interpreter exit), not on every call.
The on-disk layout is NOT LMDB compatible with ShareDB stores (out of scope).

Building the key set (add / multiadd, kmerSetDB.py:129-175 -- one LMDB put per k-mer in the reference): the storage
layer's own vectorised host code by default; with device=True (or NRPCALC_B200_BG_DEVICE=1) the canonical K-mers of the
A/C/G/T runs are extracted, sorted and deduplicated on the GPU (nrp_bg_build in include/nrp_b200.h) and only merged into
the persistent array here.  A requested device build raises when the CUDA library or a GPU is missing.
"""
import atexit
import json
import os
import shutil
import sys
import weakref
from functools import wraps

import numpy as np

from . import encoding

_RESERVED = ('K', 'LEN')                       # metadata keys of the reference's store, never treated as sequences
_COMPLEMENT = str.maketrans('ATGCU', 'TACGA')  # utils.py:10 (other characters pass through)
_STRICT = np.zeros(256, dtype=bool)            # upper-case A/C/G/T: what the 2-bit store holds
_STRICT[np.frombuffer(b'ACGT', dtype=np.uint8)] = True
_open_stores = weakref.WeakSet()


@atexit.register
def _flush_open_stores():
    for store in list(_open_stores):
        try:
            if store.ALIVE:
                store._persist()
        except Exception:                      # interpreter shutdown: best effort
            pass


def _canonical_string(window):
    return min(window, window.translate(_COMPLEMENT)[::-1])


class _Windows(object):
    """The canonical K-windows of a batch of (U->T normalised) sequences, split by store."""

    def __init__(self):
        self.runs = []          # maximal upper-case A/C/G/T stretches of length >= K: all their windows are 2-bit keys
        self.n_plain = 0        # number of 2-bit windows (with multiplicity)
        self.verbatim = []      # canonical strings of every other window, one entry per window

    def add_sequence(self, seq, K):
        if K >= len(seq):       # the whole string is the only window (utils.py:37-40)
            if len(seq) == K and K <= 32 and _is_plain(seq):
                self.runs.append(seq)
                self.n_plain += 1
            else:
                self.verbatim.append(_canonical_string(seq))
            return
        raw = np.frombuffer(seq.encode('utf-8', 'surrogatepass'), dtype=np.uint8)
        if raw.size != len(seq) or K > 32:                     # non-ASCII text or keys wider than 64 bits: all verbatim
            self.verbatim.extend(_canonical_string(seq[i:i + K]) for i in range(len(seq) - K + 1))
            return
        ok = _STRICT[raw]
        if ok.all():
            self.runs.append(seq)
            self.n_plain += len(seq) - K + 1
            return
        # windows that avoid every other character lie inside the plain stretches; the rest is spelled out
        bad = np.concatenate(([0], np.cumsum(~ok)))
        spoiled = np.nonzero(bad[K:] - bad[:-K] > 0)[0]
        self.verbatim.extend(_canonical_string(seq[i:i + K]) for i in spoiled.tolist())
        edges = np.flatnonzero(np.diff(np.concatenate(([0], ok.view(np.int8), [0]))))
        for lo, hi in zip(edges[0::2].tolist(), edges[1::2].tolist()):
            if hi - lo >= K:
                self.runs.append(seq[lo:hi])
                self.n_plain += hi - lo - K + 1


def _is_plain(seq):
    raw = np.frombuffer(seq.encode('utf-8', 'surrogatepass'), dtype=np.uint8)
    return raw.size == len(seq) and bool(_STRICT[raw].all())


class kmerSetDB(object):

    def __init__(self, path, homology, verbose=False, device=None):
        problem = None
        if not isinstance(path, str):
            problem = 'path must be a str, not {}'.format(type(path).__name__)
        elif not isinstance(homology, int):
            problem = 'Lmax must be an int, not {}'.format(type(homology).__name__)
        elif homology - 1 < 5:
            problem = 'Lmax must be at least 5, not {}'.format(homology - 1)
        if problem:
            print('\n[Non-Repetitive Parts Calculator - Background]\n\n [ERROR]    {}\n'.format(problem))
            raise ValueError('Invalid Path or Lmax')

        self.PATH = path.removesuffix('/') + '/'
        if not self.PATH.endswith('kmerSetDB.ShareDB'):
            self.PATH += 'kmerSetDB.ShareDB'
        self.K = int(homology)
        self._len_bias = 0
        self.VERB = bool(verbose)
        if device is None:
            device = os.environ.get('NRPCALC_B200_BG_DEVICE', '0') == '1'
        self._device = bool(device)                    # build the key set on the GPU (nrp_bg_build)
        self._reset_content()
        self._load_or_create()
        self.ALIVE = True
        _open_stores.add(self)

    # ---- content ----------------------------------------------------------------------------------------------
    def _reset_content(self):
        self._keys = np.zeros(0, dtype=np.uint64)      # sorted distinct canonical K-mers over ACGT
        self._pending = []                             # uint64 arrays waiting to be merged into _keys
        self._extra = set()                            # keys outside the 2-bit form, verbatim
        self._dirty = False
        self._version = 0                              # bumped on every content change (device table refresh)

    def has_verbatim_keys(self):
        """True when the store holds keys outside the 2-bit form (windows with other characters, whole strings shorter than K,
        every key when K > 32): only the host can answer membership for them."""
        return bool(self._extra)

    def _merged(self):
        """The sorted key array with every buffered insert merged in."""
        if self._pending:
            self._keys = np.union1d(self._keys, np.concatenate(self._pending))
            self._pending = []
        return self._keys

    @property
    def LEN(self):
        return int(self._merged().size) + len(self._extra) + self._len_bias

    @LEN.setter
    def LEN(self, value):
        self._len_bias = int(value) - int(self._merged().size) - len(self._extra)

    _len_bias = 0      # what multiremove(clear=True) subtracts beyond the keys it really removed (kmerSetDB.py:267-269)

    def _plain_keys(self, windows):
        """uint64 canonical K-mers of the plain stretches (distinct from the device, per window from the host code)."""
        if not windows.runs:
            return np.zeros(0, dtype=np.uint64)
        ascii_buf, offsets = encoding.join_parts(windows.runs)
        if self._device:
            from . import hgraph                     # the process-wide device context
            return hgraph._context().bg_build(ascii_buf, offsets, self.K)
        return encoding.canonical_kmers(ascii_buf, offsets, self.K)

    def _windows_of(self, seqs, normalise=True):
        windows = _Windows()
        for seq in seqs:
            windows.add_sequence(seq.replace('U', 'T') if normalise else seq, self.K)
        return windows

    def _insert(self, seqs):
        windows = self._windows_of(seqs)
        keys = self._plain_keys(windows)
        if keys.size:
            self._pending.append(keys)
        self._extra.update(windows.verbatim)
        self._touch()

    def _delete(self, seqs, clear=False):
        windows = self._windows_of(seqs)
        keys = self._plain_keys(windows)
        before = self.LEN
        if keys.size:
            self._keys = np.setdiff1d(self._merged(), keys)
        self._extra.difference_update(windows.verbatim)
        if clear:                                      # one decrement per streamed window, present or not
            self.LEN = before - windows.n_plain - len(windows.verbatim)
        self._touch()

    def _touch(self):
        self._version += 1
        self._dirty = True

    def _has(self, windows):
        keys = self._plain_keys(windows) if windows.runs else None
        if keys is not None and keys.size:
            store = self._merged()
            if store.size:
                pos = np.minimum(np.searchsorted(store, keys), store.size - 1)
                if bool((store[pos] == keys).any()):
                    return True
        return any(key in self._extra for key in windows.verbatim)

    def device_keys(self):
        """(sorted uint64 canonical K-mers, content version) for the GPU probe table."""
        return self._merged(), self._version

    # ---- persistence ------------------------------------------------------------------------------------------
    def _load_or_create(self):
        os.makedirs(self.PATH, exist_ok=True)
        meta_file = os.path.join(self.PATH, 'meta.json')
        if not os.path.exists(meta_file):
            self._persist(force=True)
            return
        with open(meta_file) as fh:
            meta = json.load(fh)
        self.K = int(meta['K'])                        # a stored K wins over the argument (kmerSetDB.py:60-65)
        self._extra = set(meta.get('extra', []))
        keys_file = os.path.join(self.PATH, 'keys.npy')
        if os.path.exists(keys_file):
            self._keys = np.load(keys_file)
        self.LEN = int(meta['LEN'])

    def _persist(self, force=False):
        if not (self._dirty or force):
            return
        os.makedirs(self.PATH, exist_ok=True)
        np.save(os.path.join(self.PATH, 'keys.npy'), self._merged())
        with open(os.path.join(self.PATH, 'meta.json'), 'w') as fh:
            json.dump({'K': self.K, 'LEN': self.LEN, 'extra': sorted(self._extra)}, fh)
        self._dirty = False

    # ---- helpers ----------------------------------------------------------------------------------------------
    def __repr__(self):
        return 'kmerSetDB stored at {} with {} {}-mers'.format(self.PATH.removesuffix('kmerSetDB.ShareDB'), self.LEN, self.K)

    __str__ = __repr__

    def alivemethod(method):
        @wraps(method)
        def guarded(self, *args, **kwargs):
            if not self.ALIVE:
                raise RuntimeError('kmerSetDB was closed or dropped')
            return method(self, *args, **kwargs)
        return guarded

    def _sequences(self, seq_list, action):
        """The sequences of a multi-call that are not metadata keys, with the reference's progress line when verbose."""
        batch = []
        for seq in seq_list:
            if not batch and self.VERB:
                print('\n[Background Processing]')
            if seq in _RESERVED:
                continue
            if self.VERB:
                line = '  {} Seq {}: {}...'.format(action, len(batch), seq[:10])
                sys.stdout.write(' ' * len(line) + '\r' + line)
                sys.stdout.flush()
            batch.append(seq)
        return batch

    # ---- public API (kmerSetDB.py:129-330) --------------------------------------------------------------------
    @alivemethod
    def add(self, seq):
        if seq not in _RESERVED:
            self._insert([seq])

    @alivemethod
    def multiadd(self, seq_list):
        self._insert(self._sequences(seq_list, 'Adding'))
        self._persist()
        if self.VERB:
            print()

    @alivemethod
    def __contains__(self, seq, rna=True):
        return self._has(self._windows_of([seq], normalise=rna))      # kmerSetDB.py:185-186

    @alivemethod
    def multicheck(self, seq_list):
        for seq in seq_list:
            yield self._has(self._windows_of([seq]))

    @alivemethod
    def __iter__(self):
        for key in encoding.decode_kmers(self._merged(), self.K):
            yield key
        for key in sorted(self._extra):
            yield key

    @alivemethod
    def __len__(self):
        return self.LEN

    @alivemethod
    def remove(self, seq):
        if seq not in _RESERVED:
            self._delete([seq])

    @alivemethod
    def multiremove(self, seq_list, clear=False):
        self._delete(self._sequences(seq_list, 'Removing'), clear=clear)
        self._persist()
        if self.VERB:
            print()

    def clear(self, Lmax=None):
        self.drop()
        self._reset_content()
        self._len_bias = 0
        if Lmax is not None:
            if not isinstance(Lmax, int):
                print('\n [ERROR]    Lmax must be an int, not {}\n'.format(type(Lmax).__name__))
                raise ValueError
            self.K = int(Lmax) + 1
            if Lmax < 5:
                print('\n [ERROR]    Lmax must be at least 5, not {}\n'.format(Lmax))
                raise ValueError
        self._persist(force=True)
        self.ALIVE = True
        _open_stores.add(self)

    def close(self):
        if not self.ALIVE:
            return False
        self._persist()
        self.ALIVE = False
        return True

    def drop(self):
        if not self.ALIVE:
            return False
        shutil.rmtree(self.PATH.removesuffix('kmerSetDB.ShareDB'), ignore_errors=True)
        self._dirty = False
        self.ALIVE = False
        return True


class _Adopted(object):
    """Read-only view of a FOREIGN background object (the reference's kmerSetDB behind nrpcalc_b200.patch()): K, ALIVE and
    the 2-bit key array the device probes, re-encoded from the object's public iterator (kmerSetDB.py:209-217)."""

    def __init__(self, foreign):
        self._foreign = foreign
        self._cache = None

    K = property(lambda self: int(self._foreign.K))
    ALIVE = property(lambda self: bool(self._foreign.ALIVE))

    def __len__(self):
        return len(self._foreign)

    def device_keys(self):
        stamp = len(self._foreign)
        if self._cache is None or self._cache[0] != stamp:
            K = self.K
            plain = [key for key in self._foreign if len(key) == K and K <= 32 and _is_plain(key)]
            keys = np.zeros(0, dtype=np.uint64)
            if plain:
                ascii_buf, offsets = encoding.join_parts(plain)
                keys = np.unique(encoding.canonical_kmers(ascii_buf, offsets, K))
            self._cache = (stamp, keys)
        return self._cache[1], self._cache[0]


def adopt(background):
    """None and this package's kmerSetDB pass through; any other object with the reference's background interface
    (K, ALIVE, __iter__, __len__) is wrapped so that the device path can read its keys."""
    if background is None or isinstance(background, (kmerSetDB, _Adopted)):
        return background
    if all(hasattr(background, name) for name in ('K', 'ALIVE', '__iter__', '__len__')):
        return _Adopted(background)
    return background
