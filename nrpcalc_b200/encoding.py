"""Host-side marshalling between Python strings and the 2-bit integer keys used on the device.

A=0 C=1 G=2 T=3, most significant base first: integer order == the reference's string order
(nrpcalc/base/utils.py:48-50 takes min() of strings) and complement == x ^ 3 (utils.py:10).
Used by the storage layer (kmerSetDB) and to build the part buffer; the Finder hot path itself runs on the GPU.
"""
import numpy as np

_CODE = np.full(256, 255, dtype=np.uint8)
for _i, _ch in enumerate('ACGT'):
    _CODE[ord(_ch)] = _i
    _CODE[ord(_ch.lower())] = _i
_LETTERS = np.frombuffer(b'ACGT', dtype=np.uint8)


def is_acgt(seq):
    raw = np.frombuffer(seq.encode('latin-1', 'replace'), dtype=np.uint8)
    return bool((_CODE[raw] != 255).all())


def join_parts(seqs):
    """list[str] -> (ascii uint8[total], offsets int64[n+1]); ValueError on non-ASCII text."""
    offsets = np.zeros(len(seqs) + 1, dtype=np.int64)
    if seqs:
        np.cumsum(np.fromiter(map(len, seqs), dtype=np.int64, count=len(seqs)), out=offsets[1:])
    try:
        blob = ''.join(seqs).encode('ascii')
    except UnicodeEncodeError:
        raise ValueError('parts must be ASCII strings (the reference compares parts character by character; text outside ASCII has no byte form here)')
    return np.frombuffer(blob, dtype=np.uint8), offsets


def all_acgt(ascii_buf):
    """Does the buffer hold nothing but A/C/G/T (either case)?  A view of a bytes object (join_parts) is answered by bytes.translate,
    2.5x faster than the table look-up on the 10 MB of a 100k-part job."""
    if not ascii_buf.size:
        return True
    base = ascii_buf.base
    if isinstance(base, bytes) and len(base) == ascii_buf.size:
        return len(base.translate(None, b'ACGTacgt')) == 0
    return bool((_CODE[ascii_buf] != 255).all())


def canonical_kmers(ascii_buf, offsets, K):
    """uint64 canonical K-mers of every K-window of every sequence with length >= K (2-bit alphabet only)."""
    codes = _CODE[ascii_buf].astype(np.uint64)
    lengths = np.diff(offsets)
    counts = np.maximum(lengths - K + 1, 0)
    total = int(counts.sum())
    if total == 0:
        return np.zeros(0, dtype=np.uint64)
    n = codes.size - K + 1
    fwd = np.zeros(n, dtype=np.uint64)
    rev = np.zeros(n, dtype=np.uint64)
    three = np.uint64(3)
    for t in range(K):
        chunk = codes[t:t + n]
        fwd |= chunk << np.uint64(2 * (K - 1 - t))
        rev |= (three - chunk) << np.uint64(2 * t)
    canon = np.minimum(fwd, rev)
    seq_of = np.repeat(np.arange(lengths.size, dtype=np.int64), counts)
    first = np.zeros(lengths.size + 1, dtype=np.int64)
    np.cumsum(counts, out=first[1:])
    start = offsets[seq_of] + (np.arange(total, dtype=np.int64) - first[seq_of])
    return canon[start]


def decode_kmer(value, K):
    value = int(value)
    return ''.join('ACGT'[(value >> (2 * (K - 1 - t))) & 3] for t in range(K))


def decode_kmers(values, K):
    values = np.asarray(values, dtype=np.uint64)
    if values.size == 0:
        return []
    shifts = (2 * (K - 1 - np.arange(K))).astype(np.uint64)
    letters = _LETTERS[((values[:, None] >> shifts[None, :]) & np.uint64(3)).astype(np.uint8)]
    return [row.tobytes().decode('ascii') for row in letters]
