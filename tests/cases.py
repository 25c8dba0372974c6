"""Seeded input generators shared by the golden-fixture maker and the parity tests.

All randomness comes from numpy's default_rng(seed) or random.Random(seed) so that every case is
reproducible on the GPU box (SURVEY.md 8d names the seeds of the BASELINE configurations).
"""
import random

import numpy as np

_COMP = str.maketrans('ACGT', 'TGCA')
_BASES = np.frombuffer(b'ACGT', dtype=np.uint8)


def rc(seq):
    return seq.translate(_COMP)[::-1]


def random_dna_buffer(n_parts, length, seed):
    """(ascii uint8[n*L], offsets int64[n+1]) of i.i.d. uniform ACGT parts -- the BASELINE shapes."""
    rng = np.random.default_rng(seed)
    ascii_buf = rng.integers(0, 4, size=n_parts * length, dtype=np.uint8)
    for a in range(0, ascii_buf.size, 1 << 26):    # codes -> letters in place, piece by piece
        ascii_buf[a:a + (1 << 26)] = _BASES[ascii_buf[a:a + (1 << 26)]]
    offsets = np.arange(n_parts + 1, dtype=np.int64) * length
    return ascii_buf, offsets


def random_varlen_buffer(n_parts, lo, hi, seed):
    """Variable-length parts, lengths uniform in [lo, hi] (BASELINE config 5 shape)."""
    rng = np.random.default_rng(seed)
    lengths = rng.integers(lo, hi + 1, size=n_parts, dtype=np.int64)
    offsets = np.zeros(n_parts + 1, dtype=np.int64)
    np.cumsum(lengths, out=offsets[1:])
    buf = rng.integers(0, 4, size=int(offsets[-1]), dtype=np.uint8)
    for a in range(0, buf.size, 1 << 26):          # codes -> letters in place, piece by piece (config 5 is 8.75 GB)
        buf[a:a + (1 << 26)] = _BASES[buf[a:a + (1 << 26)]]
    return buf, offsets


def buffer_to_list(ascii_buf, offsets):
    blob = ascii_buf.tobytes().decode('ascii')
    return [blob[offsets[i]:offsets[i + 1]] for i in range(len(offsets) - 1)]


def random_parts(n_parts, length, seed):
    return buffer_to_list(*random_dna_buffer(n_parts, length, seed))


def planted_parts(n_parts, seed, lo=40, hi=120, frag_lo=18, frag_hi=26, p_shared=0.35,
                  p_internal=0.08, p_dup=0.05, p_lower=0.1, p_short=0.03, k_hint=16):
    """Random parts with planted direct / reverse-complement repeats across parts and inside parts,
    exact duplicates, lower-case copies and a few parts shorter than k (SURVEY.md 8c edge cases)."""
    rnd = random.Random(seed)

    def dna(n):
        return ''.join(rnd.choice('ACGT') for _ in range(n))

    fragments = [dna(rnd.randint(frag_lo, frag_hi)) for _ in range(max(4, n_parts // 6))]
    parts = []
    for _ in range(n_parts):
        roll = rnd.random()
        if parts and roll < p_dup:
            seq = rnd.choice(parts)
            if rnd.random() < 0.5:
                seq = seq.lower()
            parts.append(seq)
            continue
        if roll < p_dup + p_short:
            parts.append(dna(rnd.randint(1, k_hint + 1)))
            continue
        seq = dna(rnd.randint(lo, hi))
        if rnd.random() < p_shared:
            frag = rnd.choice(fragments)
            if rnd.random() < 0.5:
                frag = rc(frag)
            at = rnd.randint(0, len(seq))
            seq = seq[:at] + frag + seq[at:]
        if rnd.random() < p_internal:
            a = rnd.randint(0, max(0, len(seq) - frag_hi))
            piece = seq[a:a + rnd.randint(frag_lo, frag_hi)]
            if rnd.random() < 0.5:
                piece = rc(piece)
            seq = seq + dna(rnd.randint(0, 5)) + piece
        if rnd.random() < p_lower:
            seq = seq.lower()
        parts.append(seq)
    return parts


def palindrome_parts(seed, n_parts=40, k=6):
    """Even-k inputs with planted palindromic windows (a window equal to its own reverse complement)."""
    rnd = random.Random(seed)
    pals = []
    for _ in range(6):
        half = ''.join(rnd.choice('ACGT') for _ in range(k // 2))
        pals.append(half + rc(half))
    parts = []
    for i in range(n_parts):
        seq = ''.join(rnd.choice('ACGT') for _ in range(rnd.randint(k, 4 * k)))
        if i % 3 == 0:
            at = rnd.randint(0, len(seq))
            seq = seq[:at] + rnd.choice(pals) + seq[at:]
        parts.append(seq)
    return parts


def random_genome(n_bp, seed):
    rng = np.random.default_rng(seed)
    return _BASES[rng.integers(0, 4, size=n_bp, dtype=np.uint8)].tobytes().decode('ascii')


def background_seqs_for(parts, seed, n_extra=5, take=0.2, span=40):
    """Background sequences: random DNA plus substrings of some of the parts (so that parts get excluded)."""
    rnd = random.Random(seed)
    seqs = [''.join(rnd.choice('ACGT') for _ in range(rnd.randint(50, 200))) for _ in range(n_extra)]
    for seq in parts:
        if len(seq) >= 30 and rnd.random() < take:
            a = rnd.randint(0, len(seq) - 25)
            piece = seq[a:a + span].upper()
            if rnd.random() < 0.5:
                piece = rc(piece)
            seqs.append(piece)
    return seqs
