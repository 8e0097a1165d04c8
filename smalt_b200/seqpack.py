"""Host-side helpers for the reference's packed sequence store (the `.sma` layout):
3 bits per base, 10 bases per little-endian 32-bit word, base i in bits 3*(9 - i%10)
of word i/10 (reference: src/sequence.c:1360-1424 compressSeq / :1499-1550 uncompressSeq).
Sequences of a set are concatenated, each followed by one terminator code 7."""
import numpy as np

TERM = 7


def pack3(codes):
    """codes: uint8 array of 3-bit codes -> uint32 words."""
    codes = np.asarray(codes, np.uint8)
    n = len(codes)
    nw = (n + 9) // 10
    out = np.empty(nw, np.uint32)
    shifts = (3 * (9 - np.arange(10))).astype(np.uint32)
    CH = 1 << 23                       # words per slice: human-scale sets without tens of GB of temporaries
    for w0 in range(0, nw, CH):
        w1 = min(nw, w0 + CH)
        part = codes[10 * w0:min(n, 10 * w1)]
        padded = np.zeros((w1 - w0) * 10, np.uint32)
        padded[:len(part)] = part & 7
        out[w0:w1] = (padded.reshape(w1 - w0, 10) << shifts).sum(axis=1, dtype=np.uint32)
    return out


def unpack3(words, n):
    words = np.asarray(words, np.uint32)
    shifts = (3 * (9 - np.arange(10))).astype(np.uint32)
    out = ((words[:, None] >> shifts[None, :]) & 7).astype(np.uint8).reshape(-1)
    return out[:n]


def concat_set(seqs):
    """-> (codes incl. terminators, seq_offs[nseq+1]) as the reference's SeqSet lays them out
    (offsets count the terminator of every previous sequence, sequence.h SEQSET_TERMCHAR)."""
    offs = np.zeros(len(seqs) + 1, np.uint64)
    parts = []
    pos = 0
    for i, s in enumerate(seqs):
        offs[i] = pos
        parts.append(np.asarray(s, np.uint8))
        parts.append(np.array([TERM], np.uint8))
        pos += len(s) + 1
    offs[len(seqs)] = pos
    return np.concatenate(parts), offs
