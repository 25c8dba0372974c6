"""Background k-mer set -- drop-in for the reference's kmerSetDB (nrpcalc/base/kmerSetDB.py:12-330).

Same public surface (K, ALIVE, add, multiadd, __contains__, multicheck, __iter__, __len__, remove,
multiremove, clear, close, drop, repr) and the same semantics: sequences are U->T normalised, every
canonical K-window becomes a key, and a query is True iff ANY of its canonical K-windows is a key
(kmerSetDB.py:177-192).

What changes is the storage and who answers the hot path's membership test:
  * the reference keeps one LMDB record per k-mer through the third-party ShareDB; here the keys over
    A/C/G/T are one sorted uint64 array in the 2-bit encoding (persisted as keys.npy under
    <path>/kmerSetDB.ShareDB/), which is exactly the form the GPU probe table is built from
    (`device_keys()` -> nrp_bg_set in include/nrp_b200.h);
  * keys that do not fit the 2-bit form (whole strings shorter than K, utils.py:37-40, or k-mers with other
    letters) are kept verbatim in a small side set so that add/remove/len/iter behave like the reference;
    they can never match a window of an A/C/G/T part.
The on-disk layout is NOT LMDB compatible with ShareDB stores (out of scope).

Building the key set (add / multiadd, kmerSetDB.py:129-175 -- one LMDB put per k-mer in the reference): the storage
layer's own vectorised host code by default; with device=True (or NRPCALC_B200_BG_DEVICE=1) the canonical K-mers are
extracted, sorted and deduplicated on the GPU (nrp_bg_build in include/nrp_b200.h) and only merged into the persistent
array here.  A requested device build raises when the CUDA library or a GPU is missing -- there is no silent fallback.
"""
import json
import os
import shutil
import sys
from functools import wraps

import numpy as np

from . import encoding


def _canonical_strings(seq, K):
    """canonical windows as strings, reference semantics (utils.py:37-50)"""
    table = str.maketrans('ATGCU', 'TACGA')
    if K >= len(seq):
        wins = [seq]
    else:
        wins = [seq[i:i + K] for i in range(len(seq) - K + 1)]
    return [min(w, w.translate(table)[::-1]) for w in wins]


class kmerSetDB(object):

    def __init__(self, path, homology, verbose=False, device=None):
        try:
            if not isinstance(path, str):
                print('\n[Non-Repetitive Parts Calculator - Background]')
                print('\n [ERROR]    Path must be a string, not {}'.format(path))
                print(' [SOLUTION] Try correcting Path\n')
                raise ValueError
            if not isinstance(homology, int):
                print('\n[Non-Repetitive Parts Calculator - Background]')
                print('\n [ERROR]    Lmax must be an integer, not {}'.format(type(homology)))
                print(' [SOLUTION] Try correcting Lmax\n')
                raise ValueError
            if homology - 1 < 5:
                print('\n[Non-Repetitive Parts Calculator - Background]')
                print('\n [ERROR]    Lmax must be greater than 4, not \'{}\''.format(homology - 1))
                print(' [SOLUTION] Try correcting Lmax\n')
                raise ValueError
        except ValueError:
            raise ValueError('Invalid Path or Lmax')

        self.PATH = path.removesuffix('/') + '/'
        if not self.PATH.endswith('kmerSetDB.ShareDB'):
            self.PATH += 'kmerSetDB.ShareDB'
        self._keys = np.zeros(0, dtype=np.uint64)     # sorted unique canonical K-mers over ACGT
        self._extra = set()                            # keys outside the 2-bit form, verbatim
        self.LEN = 0
        self.K = int(homology)
        self._version = 0                              # bumped on every content change (device table refresh)
        if device is None:
            device = os.environ.get('NRPCALC_B200_BG_DEVICE', '0') == '1'
        self._device = bool(device)                    # build the key set on the GPU (nrp_bg_build)
        self._open_store()
        self.VERB = bool(verbose)
        self.ALIVE = True

    # ---- persistence ----------------------------------------------------------------------------------
    def _open_store(self):
        os.makedirs(self.PATH, exist_ok=True)
        meta_file = os.path.join(self.PATH, 'meta.json')
        if os.path.exists(meta_file):
            with open(meta_file) as fh:
                meta = json.load(fh)
            self.K = int(meta['K'])                    # a stored K wins over the argument (kmerSetDB.py:60-65)
            self.LEN = int(meta['LEN'])
            self._extra = set(meta.get('extra', []))
            keys_file = os.path.join(self.PATH, 'keys.npy')
            if os.path.exists(keys_file):
                self._keys = np.load(keys_file)
        else:
            self._sync()

    def _sync(self):
        np.save(os.path.join(self.PATH, 'keys.npy'), self._keys)
        with open(os.path.join(self.PATH, 'meta.json'), 'w') as fh:
            json.dump({'K': self.K, 'LEN': self.LEN, 'extra': sorted(self._extra)}, fh)

    # ---- helpers -----------------------------------------------------------------------------------------
    def __repr__(self):
        return 'kmerSetDB stored at {} with {} {}-mers'.format(
            self.PATH.removesuffix('kmerSetDB.ShareDB'), self.LEN, self.K)

    def __str__(self):
        return self.__repr__()

    def alivemethod(method):
        @wraps(method)
        def wrapper(self, *args, **kwargs):
            if self.ALIVE:
                return method(self, *args, **kwargs)
            raise RuntimeError('kmerSetDB was closed or dropped')
        return wrapper

    def _verb_action(self, action, index, seq):
        if self.VERB:
            printed = '  {} Seq {}: {}...'.format(action, index, seq[:min(len(seq), 10)])
            sys.stdout.write(' ' * len(printed) + '\r')
            sys.stdout.write(printed)
            sys.stdout.flush()

    def _split(self, seqs):
        """U->T normalise; return (uint64 canonical K-mers of the ACGT sequences with len > K or == K,
        list of verbatim keys for everything else)."""
        plain, odd = [], []
        for seq in seqs:
            seq = seq.replace('U', 'T')
            if len(seq) >= self.K and self.K <= 32 and encoding.is_acgt(seq) and seq.isupper():
                plain.append(seq)
            else:
                odd.extend(_canonical_strings(seq, self.K))
        if plain:
            ascii_buf, offsets = encoding.join_parts(plain)
            if self._device:
                from . import hgraph                 # the process-wide device context
                ints = hgraph._context().bg_build(ascii_buf, offsets, self.K)      # distinct, ascending
            else:
                ints = encoding.canonical_kmers(ascii_buf, offsets, self.K)
        else:
            ints = np.zeros(0, dtype=np.uint64)
        return ints, odd

    def _has_int(self, values):
        if self._keys.size == 0:
            return np.zeros(values.size, dtype=bool)
        pos = np.searchsorted(self._keys, values)
        pos[pos == self._keys.size] = self._keys.size - 1
        return self._keys[pos] == values

    def _insert(self, seqs):
        ints, odd = self._split(seqs)
        if ints.size:
            merged = np.union1d(self._keys, ints)
            self.LEN += merged.size - self._keys.size
            self._keys = merged
        for key in odd:
            if key not in self._extra:
                self._extra.add(key)
                self.LEN += 1
        self._version += 1

    def _delete(self, seqs, clear=False):
        ints, odd = self._split(seqs)
        if ints.size:
            if clear:
                self.LEN -= ints.size                  # the reference decrements per streamed k-mer (kmerSetDB.py:267-269)
            kept = np.setdiff1d(self._keys, ints)
            if not clear:
                self.LEN -= self._keys.size - kept.size
            self._keys = kept
        for key in odd:
            if clear:
                self._extra.discard(key)
                self.LEN -= 1
            elif key in self._extra:
                self._extra.remove(key)
                self.LEN -= 1
        self._version += 1

    def device_keys(self):
        """(sorted uint64 canonical K-mers, content version) for the GPU probe table."""
        return self._keys, self._version

    # ---- public API (kmerSetDB.py:129-330) ------------------------------------------------------------
    @alivemethod
    def add(self, seq):
        if seq not in ['K', 'LEN']:
            self._insert([seq])
        self._sync()

    @alivemethod
    def multiadd(self, seq_list):
        started = False
        index = 0
        batch = []
        for seq in seq_list:
            if not started and self.VERB:
                print('\n[Background Processing]')
                started = True
            if seq not in ['K', 'LEN']:
                self._verb_action('Adding', index, seq)
                index += 1
                batch.append(seq)
        self._insert(batch)
        self._sync()
        if self.VERB:
            print()

    @alivemethod
    def __contains__(self, seq, rna=True):
        if rna:
            seq = seq.replace('U', 'T')
        ints, odd = self._split([seq])
        if ints.size and self._has_int(ints).any():
            return True
        return any(key in self._extra for key in odd)

    @alivemethod
    def multicheck(self, seq_list):
        for seq in seq_list:
            yield self.__contains__(seq.replace('U', 'T'), rna=False)

    @alivemethod
    def __iter__(self):
        for key in encoding.decode_kmers(self._keys, self.K):
            yield key
        for key in sorted(self._extra):
            yield key

    @alivemethod
    def __len__(self):
        return self.LEN

    @alivemethod
    def remove(self, seq):
        if seq not in ['K', 'LEN']:
            self._delete([seq])
            self._sync()

    @alivemethod
    def multiremove(self, seq_list, clear=False):
        started = False
        index = 0
        batch = []
        for seq in seq_list:
            if not started and self.VERB:
                print('\n[Background Processing]')
                started = True
            if seq not in ['K', 'LEN']:
                self._verb_action('Removing', index, seq)
                index += 1
                batch.append(seq)
        self._delete(batch, clear=clear)
        self._sync()
        if self.VERB:
            print()

    def clear(self, Lmax=None):
        self.drop()
        self._keys = np.zeros(0, dtype=np.uint64)
        self._extra = set()
        self.LEN = 0
        if Lmax is not None:
            if not isinstance(Lmax, int):
                print('\n [ERROR]    Lmax must be an integer, not {}'.format(type(Lmax)))
                print(' [SOLUTION] Try correcting Lmax\n')
                raise ValueError
            self.K = int(Lmax) + 1
            if Lmax < 5:
                print('\n [ERROR]    Lmax must be greater than 4, not {}'.format(Lmax))
                print(' [SOLUTION] Try correcting Lmax\n')
                raise ValueError
        os.makedirs(self.PATH, exist_ok=True)
        self._version += 1
        self._sync()
        self.ALIVE = True

    def close(self):
        if self.ALIVE:
            self._sync()
            self.ALIVE = False
            return True
        return False

    def drop(self):
        if self.ALIVE:
            shutil.rmtree(self.PATH.removesuffix('kmerSetDB.ShareDB'))
            self.ALIVE = False
            return True
        return False
