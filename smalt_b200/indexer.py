"""Host-side builder/writer of the reference's index files, so that benchmarks and tests
are self-contained (no reference binary needed at run time).  Produces byte-identical
`.smi` / `.sma` files to `smalt index` (checked by tests/test_indexer.py against the compiled
reference).  Restated from /root/reference/src: smalt.c:268-332 (selectHashTyp),
hashidx.c:465-531 (k-mer grid over the concatenated sequences), :829-998 (table layout),
:1214-1255 (.smi), sequence.c:1360-1424 + :2448-2519 (.sma), filio.c:55-77 (common header).
This is plumbing around the hot path, vectorised with numpy."""
import numpy as np

from .seqpack import pack3

SIGNATURE = 0x73212173
ENDIANTEST = 0x6E378A19
FILIOTYP_SEQSET = 1
FILIOTYP_HASHTAB = 2


def hash32mix(a):
    """hashidx.c:163-172 on uint32 arrays"""
    a = a.astype(np.uint32)
    with np.errstate(over="ignore"):
        a = (a + np.uint32(0x7ed55d16)) + (a << np.uint32(12))
        a = (a ^ np.uint32(0xc761c23c)) ^ (a >> np.uint32(19))
        a = (a + np.uint32(0x165667b1)) + (a << np.uint32(5))
        a = (a + np.uint32(0xd3a2646c)) ^ (a << np.uint32(9))
        a = (a + np.uint32(0xfd7046c5)) + (a << np.uint32(3))
        a = (a ^ np.uint32(0xb55a4f09)) ^ (a >> np.uint32(16))
    return a


def select_hash_type(k, nskip, totlen):
    """smalt.c:268-332 -> (typ, nbits_key, nbits_perf)"""
    nbk = 2 * k
    nskip = max(nskip, 1)
    ntup = totlen // nskip
    if (1 << nbk) <= 2 * ntup:
        return 0, nbk, 0
    last_b = 1 if (ntup & 1) else 0
    t = ntup
    for i in range(32):
        t >>= 1
        if t & 1:
            last_b = i
    nbits_key = last_b + 1 if (last_b & 1) else last_b
    nbits_perf = nbk - 32 if nbk > 32 else 0
    if nbits_key + nbits_perf > 26:
        nbits_key = 26 - nbits_perf
    if nbits_key < nbits_perf + 1:
        nbits_key = nbits_perf + 1
    if nbits_key > 26:
        nbits_key = 26
    return 1, nbits_key, nbits_perf


def _kmers_of_set(seqs, k, nskip):
    """All hashed words of the set in scan order: (words uint64, serial numbers uint32), and
    the final k-mer counter.  Follows doWordsInSeq (hashidx.c:465-531): the k-mer grid runs
    over the concatenated sequences; words with a non-standard base are skipped (but counted)."""
    words, serials = [], []
    tuplectr = 0
    offs = 0
    for s in seqs:
        s = np.asarray(s, np.uint8)
        L = len(s)
        if L < k:
            raise ValueError("sequence shorter than k (ERRCODE_SHORTSEQ)")
        n_k = (L - k - offs) // nskip + 1 if L - k - offs >= 0 else 0
        if n_k > 0:
            starts = offs + nskip * np.arange(n_k, dtype=np.int64)
            nonstd = np.concatenate([[0], np.cumsum((s & 4) != 0)])
            ok = (nonstd[starts + k] - nonstd[starts]) == 0
            w = np.zeros(n_k, np.uint64)
            c2 = (s & 3).astype(np.uint64)
            for b in range(k):
                w = (w << np.uint64(2)) | c2[starts + b]
            words.append(w[ok])
            serials.append((tuplectr + np.arange(n_k, dtype=np.int64))[ok])
            last_end = offs + (n_k - 1) * nskip + k - 1
            ktup_i = nskip - (L - 1 - last_end)
        else:
            ktup_i = k + offs - L
        tuplectr += n_k
        d = k - ktup_i
        offs = int(np.fmod(d, nskip))  # C remainder (sign of the dividend)
        if offs:
            offs = nskip - offs
        tuplectr += int(np.trunc((k - ktup_i + offs) / nskip))
    if words:
        return np.concatenate(words), np.concatenate(serials).astype(np.uint32), tuplectr
    return np.zeros(0, np.uint64), np.zeros(0, np.uint32), tuplectr


def build_index(seqs, k=13, nskip=6):
    """-> dict with the in-memory table after hashTableSetUp (posidx[nwords] == npos)."""
    totlen = int(sum(len(s) for s in seqs))
    typ, nbits_key, nbits_lo = select_hash_type(k, nskip, totlen)
    words, serials, tuplectr = _kmers_of_set(seqs, k, nskip)
    npos = len(words)
    maxpos = tuplectr - 1 if tuplectr > 0 else 0
    if typ == 0:
        nkeys = 1 << (2 * k)
        key = words.astype(np.int64)
        order = np.argsort(key, kind="stable")
        idx = np.zeros(nkeys + 1, np.uint32)
        idx[1:] = np.cumsum(np.bincount(key, minlength=nkeys)).astype(np.uint32)
        return dict(typ=0, wordlen=k, nskip=nskip, nbits_key=2 * k, nbits_lo=0, npos=npos, nwords=0,
                    maxpos=maxpos, nkeys=nkeys, idx=idx, pos=serials[order], wordidx=None, posidx=None)
    nkeys = 1 << nbits_key
    wordmask_lo = np.uint64((1 << nbits_lo) - 1)
    word_hi = (words >> np.uint64(nbits_lo)).astype(np.uint32)
    keymod = np.uint32(1 << (nbits_key - nbits_lo))
    key = ((hash32mix(word_hi) % keymod).astype(np.uint64) << np.uint64(nbits_lo)) + (words & wordmask_lo)
    comp = (key << np.uint64(32)) | word_hi.astype(np.uint64)  # sort by (key, word_hi), stable in scan order
    order = np.argsort(comp, kind="stable")
    comp_s = comp[order]
    if npos:
        new_word = np.concatenate([[True], comp_s[1:] != comp_s[:-1]])
    else:
        new_word = np.zeros(0, bool)
    word_start = np.flatnonzero(new_word)
    nwords = len(word_start)
    wordidx = np.zeros(nwords + 1, np.uint32)
    wordidx[:nwords] = (comp_s[word_start] & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    posidx = np.zeros(nwords + 1, np.uint32)
    posidx[:nwords] = word_start.astype(np.uint32)
    posidx[nwords] = npos
    wkey = (comp_s[word_start] >> np.uint64(32)).astype(np.int64)
    idx = np.zeros(nkeys + 1, np.uint32)
    idx[1:] = np.cumsum(np.bincount(wkey, minlength=nkeys)).astype(np.uint32)
    return dict(typ=1, wordlen=k, nskip=nskip, nbits_key=nbits_key, nbits_lo=nbits_lo, npos=npos,
                nwords=nwords, maxpos=maxpos, nkeys=nkeys, idx=idx, pos=serials[order],
                wordidx=wordidx, posidx=posidx)


def kmer_grid(seqs, k, nskip):
    """The k-mer grid bookkeeping of doWordsInSeq (hashidx.c:465-531) per sequence, without touching the
    bases: -> (grid array for smb_index_build, final k-mer counter)."""
    from .capi import INDEX_SEQ_DTYPE
    grid = np.zeros(len(seqs), INDEX_SEQ_DTYPE)
    tuplectr = 0
    offs = 0
    start = 0
    for n, s in enumerate(seqs):
        L = len(s)
        if L < k:
            raise ValueError("sequence shorter than k (ERRCODE_SHORTSEQ)")
        n_k = (L - k - offs) // nskip + 1 if L - k - offs >= 0 else 0
        grid[n] = (start, offs, n_k, tuplectr, 0)
        if n_k > 0:
            last_end = offs + (n_k - 1) * nskip + k - 1
            ktup_i = nskip - (L - 1 - last_end)
        else:
            ktup_i = k + offs - L
        tuplectr += n_k
        d = k - ktup_i
        offs = int(np.fmod(d, nskip))
        if offs:
            offs = nskip - offs
        tuplectr += int(np.trunc((k - ktup_i + offs) / nskip))
        start += L
    return grid, tuplectr


def build_index_gpu(ctx, seqs, k=13, nskip=6, upload=True, words=None):
    """build_index on the GPU (csrc/index_build.cu through smb_index_build): same dict, same bytes.
    words: the 3-bit packed set (pack3 of all sequences + one terminator) if the caller has it already."""
    from .seqpack import pack3
    totlen = int(sum(len(s) for s in seqs))
    typ, nbits_key, nbits_lo = select_hash_type(k, nskip, totlen)
    grid, tuplectr = kmer_grid(seqs, k, nskip)
    if upload:
        if words is None:
            words = pack3(np.concatenate([np.asarray(s, np.uint8) & 7 for s in seqs] + [np.array([7], np.uint8)]))
        offs = np.concatenate([[0], np.cumsum([len(s) for s in seqs])]).astype(np.uint64)
        ctx.refseq_upload(words, totlen, offs)
    r = ctx.index_build(k, nskip, typ, nbits_key if typ else 2 * k, nbits_lo, grid)
    maxpos = tuplectr - 1 if tuplectr > 0 else 0
    return dict(typ=typ, wordlen=k, nskip=nskip, nbits_key=nbits_key if typ else 2 * k, nbits_lo=nbits_lo if typ else 0,
                npos=r["npos"], nwords=r["nwords"], maxpos=maxpos, nkeys=r["nkeys"], idx=r["idx"], pos=r["pos"],
                wordidx=r["wordidx"], posidx=r["posidx"], kernel_ms=r["kernel_ms"])


def as_loaded(ix):
    """The table as hashTableRead leaves it: posidx[nwords] is not read back (hashidx.c:1334)."""
    out = dict(ix)
    if ix["typ"] != 0:
        p = ix["posidx"].copy()
        p[ix["nwords"]] = 0
        out["posidx"] = p
    return out


def _common_header(filsiz, typ, version, headsiz):
    h = np.zeros(12, "<u4")
    h[:6] = (SIGNATURE, ENDIANTEST, (filsiz + 12) & 0xFFFFFFFF, typ, version, headsiz)  # + IOFIL_HEADSIZ
    return h.tobytes()


def write_smi(prefix, ix):
    hd = np.array([ix["wordlen"], ix["nskip"], ix["npos"], ix["maxpos"], ix["typ"], ix["nbits_key"],
                   ix["nbits_lo"], ix["nwords"]], "<u4")
    totsiz = ix["npos"] + ix["nkeys"] + 1 + (0 if ix["typ"] == 0 else 2 * (ix["nwords"] + 1))
    with open(prefix + ".smi", "wb") as f:
        f.write(_common_header(totsiz, FILIOTYP_HASHTAB, 3, 8))
        f.write(hd.tobytes())
        f.write(ix["idx"].astype("<u4").tobytes())
        f.write(ix["pos"].astype("<u4").tobytes())
        if ix["typ"] != 0:
            f.write(ix["wordidx"].astype("<u4").tobytes())
            f.write(ix["posidx"].astype("<u4").tobytes())


def write_sma(prefix, names, seqs, flags=2, words=None):
    """sequence.c:2448-2519; the sequences are stored back to back (no terminators between
    them with the driver's default flags), one terminator code at the very end."""
    nseq = len(seqs)
    nam = b"".join(n.encode() + b"\0" for n in names)
    seqsiz = int(sum(len(s) for s in seqs))
    if words is None:
        words = pack3(np.concatenate([np.asarray(s, np.uint8) & 7 for s in seqs] + [np.array([7], np.uint8)]))
    assert len(words) == seqsiz // 10 + 1
    hd = np.array([nseq & 0xFFFFFFFF, nseq >> 32, len(nam) & 0xFFFFFFFF, len(nam) >> 32,
                   seqsiz & 0xFFFFFFFF, seqsiz >> 32, flags, 0], "<u4")
    seqnamsiz = (len(nam) - 1) // 4 + 1
    totsiz = 8 + len(words) + nseq + seqnamsiz
    with open(prefix + ".sma", "wb") as f:
        f.write(_common_header(totsiz, FILIOTYP_SEQSET, 4, 8))
        f.write(hd.tobytes())
        f.write(nam)
        f.write(np.array([len(s) for s in seqs], "<u4").tobytes())
        f.write(words.astype("<u4").tobytes())
    return words
