"""Readers for the reference's binary index files (wire formats restated from
/root/reference/src: filio.c:55-77 common 12-word header; hashidx.c:1214-1366 `.smi`
hash index, version 3, 8-word header; sequence.c:2448-2680 `.sma` sequence set,
8-word header).  Host-side plumbing: the arrays are handed to smb_index_upload /
smb_refseq_upload unchanged."""
import numpy as np

SIGNATURE = 0x73212173
ENDIANTEST = 0x6E378A19
FILIOTYP_SEQSET = None  # not checked: the type-specific headers are validated instead


def _common_header(buf, path):
    h = np.frombuffer(buf, "<u4", 12)
    if h[0] != SIGNATURE:
        raise ValueError("%s: not a SMALT binary file" % path)
    if h[1] != ENDIANTEST:
        raise ValueError("%s: written with a different endianness" % path)
    return dict(filsiz=int(h[2]), typ=int(h[3]) & 0xFF, version=int(h[4]), headsiz=int(h[5]))


def read_smi(prefix):
    """`<prefix>.smi` -> dict with the arrays exactly as hashTableRead leaves them in memory:
    note that it reads 2*nwords+1 words of the (wordidx, posidx) block although
    2*(nwords+1) were written, so posidx[nwords] stays 0 (hashidx.c:1334 vs :1245-1246)."""
    path = prefix + ".smi"
    buf = open(path, "rb").read()
    com = _common_header(buf, path)
    if com["version"] != 3:
        raise ValueError("%s: unsupported hash index version %d" % (path, com["version"]))
    hd = np.frombuffer(buf, "<u4", 8, 48)
    wordlen, nskip, npos, maxpos, typ, nbits_key, nbits_lo, nwords = (int(x) for x in hd)
    nkeys = (1 << (2 * wordlen)) if typ == 0 else (1 << nbits_key)
    off = 48 + 32
    idx = np.frombuffer(buf, "<u4", nkeys + 1, off).copy()
    off += 4 * (nkeys + 1)
    pos = np.frombuffer(buf, "<u4", npos, off).copy()
    off += 4 * npos
    wordidx = posidx = None
    if typ != 0:
        blk = np.zeros(2 * (nwords + 1), np.uint32)
        blk[:2 * nwords + 1] = np.frombuffer(buf, "<u4", 2 * nwords + 1, off)
        wordidx = blk[:nwords + 1]
        posidx = blk[nwords + 1:]
    return dict(typ=typ, wordlen=wordlen, nskip=nskip, nbits_key=nbits_key, nbits_lo=nbits_lo,
                npos=npos, nwords=nwords, maxpos=maxpos, nkeys=nkeys,
                idx=idx, pos=pos, wordidx=wordidx, posidx=posidx)


def read_sma(prefix):
    """`<prefix>.sma` -> dict(names, seq_offs[nseq+1] (base offsets in the concatenated set,
    one terminator per sequence), words (3-bit packed, 10 bases/word), nbases)."""
    path = prefix + ".sma"
    buf = open(path, "rb").read()
    com = _common_header(buf, path)
    if com["version"] != 4 and com["version"] != 3:
        raise ValueError("%s: unsupported sequence set version %d" % (path, com["version"]))
    hd = np.frombuffer(buf, "<u4", 8, 48)
    if com["version"] == 3:
        nseq = int(hd[0]); namsiz = (int(hd[2]) << 32) + int(hd[1]); seqsiz = (int(hd[4]) << 32) + int(hd[3])
    else:
        nseq = (int(hd[1]) << 32) + int(hd[0])
        namsiz = (int(hd[3]) << 32) + int(hd[2])
        seqsiz = (int(hd[5]) << 32) + int(hd[4])
    off = 48 + 32
    names = [s.decode() for s in buf[off:off + namsiz].split(b"\0")[:nseq]]
    off += namsiz
    seqlen = np.frombuffer(buf[off:off + 4 * nseq], "<u4").astype(np.uint64)
    off += 4 * nseq
    nwords = seqsiz // 10 + 1
    words = np.frombuffer(buf[off:off + 4 * nwords], "<u4").copy()
    seq_offs = np.zeros(nseq + 1, np.uint64)
    seq_offs[1:] = np.cumsum(seqlen)
    return dict(names=names, seq_offs=seq_offs, words=words, nbases=seqsiz + 1, seqsiz=seqsiz)
