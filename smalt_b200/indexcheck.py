"""Size-independent checks of a hash index (dict of smalt_b200.indexer) against the sequences it was built
from - what the parity tests use at sizes where the host builder (itself pinned against `smalt index`) is too
slow: every array is structurally what hashTableSetUp leaves (hashidx.c:829-998), and sampled grid positions
are found under their own k-mer word by the lookup of hashTableGetKtupleHits (hashidx.c:1146-1191), restated
here in numpy."""
import numpy as np

from .indexer import kmer_grid


def hash32mix(a):   # hashidx.c:163-172 on uint32 arrays
    a = a.astype(np.uint64)
    m = np.uint64(0xFFFFFFFF)
    a = ((a + np.uint64(0x7ed55d16)) + (a << np.uint64(12))) & m
    a = ((a ^ np.uint64(0xc761c23c)) ^ (a >> np.uint64(19))) & m
    a = ((a + np.uint64(0x165667b1)) + (a << np.uint64(5))) & m
    a = ((a + np.uint64(0xd3a2646c)) ^ (a << np.uint64(9))) & m
    a = ((a + np.uint64(0xfd7046c5)) + (a << np.uint64(3))) & m
    a = ((a ^ np.uint64(0xb55a4f09)) ^ (a >> np.uint64(16))) & m
    return a


def check_structure(ix):
    """monotone offset arrays, totals, order of the words inside a key and of the positions inside a word"""
    idx = ix["idx"].astype(np.int64)
    pos = ix["pos"]
    npos, nwords = int(ix["npos"]), int(ix["nwords"])
    assert len(pos) == npos and idx[0] == 0 and np.all(np.diff(idx) >= 0), "idx not monotone"
    if ix["typ"] == 0:
        assert idx[-1] == npos, "idx does not end at npos"
        starts = idx[:-1][np.diff(idx) > 0]
    else:
        widx, pidx = ix["wordidx"].astype(np.int64), ix["posidx"].astype(np.int64)
        assert idx[-1] == nwords and len(widx) == nwords + 1 and len(pidx) == nwords + 1
        assert pidx[0] == 0 and pidx[nwords] == npos and np.all(np.diff(pidx) > 0), "posidx: empty or unordered word"
        # words strictly ascending inside a key
        inner = np.ones(nwords, bool)
        inner[idx[:-1][idx[:-1] < nwords]] = False          # first word of a key
        d = np.diff(widx[:nwords])
        assert np.all(d[inner[1:]] > 0), "wordidx not ascending inside a key"
        starts = pidx[:nwords]
    # positions ascending inside a word / key (scan order)
    inner = np.ones(npos, bool)
    inner[starts[starts < npos]] = False
    d = np.diff(pos.astype(np.int64))
    assert np.all(d[inner[1:]] > 0), "positions not ascending inside a word"


def lookup(ix, words):
    """-> (nhits, first position index) per 2k-bit word"""
    words = np.asarray(words, np.uint64)
    idx = ix["idx"]                     # (gathers only: the arrays of a 3 Gb genome are not converted as a whole)
    if ix["typ"] == 0:
        lo = idx[words.astype(np.int64)].astype(np.int64)
        return idx[words.astype(np.int64) + 1].astype(np.int64) - lo, lo
    nbl = np.uint64(ix["nbits_lo"])
    hi = (words >> nbl).astype(np.uint64)
    keymod = np.uint64(1 << (ix["nbits_key"] - ix["nbits_lo"]))
    key = ((hash32mix(hi) % keymod) << nbl) + (words & ((np.uint64(1) << nbl) - np.uint64(1)))
    key = key.astype(np.int64)
    a, b = idx[key].astype(np.int64), idx[key + 1].astype(np.int64)
    widx, pidx = ix["wordidx"], ix["posidx"]
    nh = np.zeros(len(words), np.int64)
    first = np.zeros(len(words), np.int64)
    for i in range(len(words)):   # binary search in the key's words (a few entries)
        j = a[i] + np.searchsorted(widx[a[i]:b[i]], np.uint32(hi[i]))
        if j < b[i] and widx[j] == np.uint32(hi[i]):
            first[i] = int(pidx[j])
            nh[i] = int(pidx[j + 1]) - int(pidx[j])
    return nh, first


def check_samples(ix, seqs, k, nskip, nsample=20000, seed=1):
    """sampled grid positions over standard bases are listed exactly once under their own word"""
    grid, _ = kmer_grid(seqs, k, nskip)
    rng = np.random.default_rng(seed)
    n_k = grid["n_k"].astype(np.int64)
    cum = np.concatenate([[0], np.cumsum(n_k)])
    g = rng.integers(0, cum[-1], nsample)
    si = np.searchsorted(cum, g, side="right") - 1
    local = g - cum[si]
    serial = grid["tup_base"].astype(np.int64)[si] + local
    base = grid["offs"].astype(np.int64)[si] + local * nskip
    words = np.zeros(nsample, np.uint64)
    ok = np.ones(nsample, bool)
    for n, s in enumerate(seqs):
        sel = np.nonzero(si == n)[0]
        if not len(sel):
            continue
        s = np.asarray(s, np.uint8)
        w = np.zeros(len(sel), np.uint64)
        good = np.ones(len(sel), bool)
        for b in range(k):
            c = s[base[sel] + b]
            good &= c < 4
            w = (w << np.uint64(2)) | (c & 3).astype(np.uint64)
        words[sel], ok[sel] = w, good
    nh, first = lookup(ix, words[ok])
    pos = ix["pos"]
    found = 0
    for j, sr in enumerate(serial[ok]):
        sl = pos[first[j]:first[j] + nh[j]]
        i = np.searchsorted(sl, sr)
        assert i < len(sl) and sl[i] == sr, "grid position %d not under its word" % sr
        found += 1
    # positions over non-standard bases are not in the table at all
    assert int(ix["npos"]) <= int(cum[-1])
    return found
