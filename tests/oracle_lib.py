"""ctypes bindings for the CHECKERS used by the tests (never by the product):

* ``Oracle``  - oracle/liboracle.so, the plain-C restatement (oracle/smalt_oracle.c)
* ``RefLib``  - oracle/_ref/libsmalt_ref.so, the UNMODIFIED reference compiled from
  /root/reference/src plus the shims of oracle/ref_harness.c (only present when
  oracle/_ref has been built; it travels to the GPU box with the snapshot).

Sequences are numpy uint8 arrays of 3-bit alphabet codes A0 C1 G2 T3 X4 N5.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
i32p = C.POINTER(C.c_int)

DEFAULT_PEN = (1, -2, -4, -3)  # match, mismatch, gapopen, gapext (score.c:41-47)


def _p(a, t):
    return a.ctypes.data_as(t)


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])


def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "libsmalt_ref.so"))


def ref_binary(name):
    p = os.path.join(REF_DIR, name)
    return p if os.path.exists(p) else None


class SoScoring(C.Structure):
    _fields_ = [("match", C.c_int), ("mismatch", C.c_int), ("gap_init", C.c_int),
                ("gap_ext", C.c_int), ("S", (C.c_int8 * 8) * 8)]


class SoIndex(C.Structure):
    _fields_ = [("typ", C.c_int), ("wordlen", C.c_int), ("nskip", C.c_int),
                ("nbits_key", C.c_int), ("nbits_lo", C.c_int),
                ("nkeys", C.c_uint32), ("npos", C.c_uint32), ("nwords", C.c_uint32),
                ("maxpos", C.c_uint32), ("keymod", C.c_uint32),
                ("wordmask", C.c_uint64), ("wordmask_lo", C.c_uint64),
                ("wordmask_hi", C.c_uint64),
                ("idx", u32p), ("pos", u32p), ("wordidx", u32p), ("posidx", u32p)]


class SoHitInfo(C.Structure):
    _fields_ = [("qlen", C.c_uint32), ("n_seeds", C.c_uint32), ("seed_rank", C.c_uint32),
                ("status", C.c_uint8),
                ("posidx", u32p), ("nhits", u32p), ("cix", u32p), ("qoffs", u32p),
                ("sortkey", u32p), ("sidx", u32p), ("qmask", u8p), ("qbuf", u8p),
                ("frame_cnt", u32p), ("frame_ix", u32p), ("n_alloc", C.c_uint32)]


class SoCand(C.Structure):
    _fields_ = [("qs", C.c_uint32), ("qe", C.c_uint32), ("rs", C.c_uint64), ("re", C.c_uint64),
                ("band_l", C.c_int32), ("band_r", C.c_int32), ("dqo", C.c_uint32), ("dro", C.c_int32),
                ("sqidx", C.c_int32), ("cover", C.c_uint32), ("flags", C.c_uint8)]


class SoHitList(C.Structure):
    _fields_ = [("nhits", C.c_int), ("nhits_max", C.c_int), ("nhits_alloc", C.c_int),
                ("status", C.c_uint8), ("sqdat", u64p), ("qmask", u8p), ("qlen", C.c_uint32)]


def _align_out(maxres, maxdiff):
    return (np.zeros(5 * maxres, np.int32), np.zeros(maxdiff, np.uint8),
            np.zeros(maxres, np.int32))


def _unpack_results(nres, out5, diffbuf, difflen):
    res, off = [], 0
    for i in range(nres):
        n = int(difflen[i])
        res.append((tuple(int(x) for x in out5[5 * i:5 * i + 5]), bytes(diffbuf[off:off + n])))
        off += n
    return res


class Oracle:
    """The C restatement (the checker)."""

    def __init__(self, pen=DEFAULT_PEN):
        build_oracle()
        self.lib = C.CDLL(os.path.join(ORACLE_DIR, "liboracle.so"))
        self.sc = SoScoring()
        self.lib.so_scoring_init(C.byref(self.sc), *pen)
        self.lib.so_hitinfo_create.restype = C.POINTER(SoHitInfo)
        self.lib.so_hitlist_create.restype = C.POINTER(SoHitList)
        self.lib.so_lookup.restype = C.c_uint32
        self.lib.so_cover_deficit.restype = C.c_uint32
        self.lib.so_number_of_hits.restype = C.c_uint32
        self.lib.so_hit_numbers.restype = C.c_uint32
        self.lib.so_collect_hits_segment.argtypes = [
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int]
        self.lib.so_lookup.argtypes = [C.c_void_p, C.c_uint64, u32p]

    def sw_striped(self, read, ref):
        score = C.c_int(0)
        err = self.lib.so_sw_striped(C.byref(self.sc), _p(read, u8p), len(read),
                                     _p(ref, u8p), len(ref), C.byref(score))
        return err, score.value

    def band_fast(self, read, ref, l_edge, r_edge, pl, pr, ul, ur):
        score = C.c_int(0)
        cells = C.c_longlong(0)
        err = self.lib.so_band_fast(C.byref(self.sc), _p(read, u8p), len(read),
                                    _p(ref, u8p), len(ref), l_edge, r_edge, pl, pr, ul, ur,
                                    C.byref(score), C.byref(cells))
        return err, score.value, cells.value

    def band_align(self, read, ref, l_edge, r_edge, pl, pr, ul, ur, minscore, minscorlen,
                   maxres=256):
        maxdiff = maxres * (len(read) + len(ref) + 8)
        out5, diffbuf, difflen = _align_out(maxres, maxdiff)
        nres = C.c_int(0)
        cells = C.c_longlong(0)
        err = self.lib.so_band_align(C.byref(self.sc), _p(read, u8p), len(read),
                                     _p(ref, u8p), len(ref), l_edge, r_edge, pl, pr, ul, ur,
                                     minscore, minscorlen, maxres, C.byref(nres),
                                     _p(out5, i32p), maxdiff, _p(diffbuf, u8p),
                                     _p(difflen, i32p), C.byref(cells))
        return err, _unpack_results(nres.value, out5, diffbuf, difflen), cells.value

    # ---- K1 ----
    def make_index(self, smi):
        """smi: dict from smalt_b200.smifile.read_smi (arrays kept alive by the caller)."""
        ix = SoIndex()
        z = np.zeros(1, np.uint32)
        widx = smi["wordidx"] if smi["wordidx"] is not None else z
        pidx = smi["posidx"] if smi["posidx"] is not None else z
        self.lib.so_index_setup(C.byref(ix), smi["typ"], smi["wordlen"], smi["nskip"],
                                smi["nbits_key"], smi["nbits_lo"],
                                C.c_uint32(smi["npos"]), C.c_uint32(smi["nwords"]),
                                _p(smi["idx"], u32p), _p(smi["pos"], u32p),
                                _p(widx, u32p), _p(pidx, u32p))
        ix._keep = (smi, widx, pidx)
        return ix

    def lookup(self, ix, words):
        nh = np.zeros(len(words), np.uint32)
        px = np.zeros(len(words), np.uint32)
        for i, w in enumerate(words):
            p = C.c_uint32(0)
            nh[i] = self.lib.so_lookup(C.byref(ix), C.c_uint64(int(w)), C.byref(p))
            px[i] = p.value
        return nh, px

    def hitinfo(self, ix, read, qual, is_reverse, is_short, maxhit_per_tuple=10000,
                maxhit_total=16384, basq=0, h=None):
        if h is None:
            h = self.lib.so_hitinfo_create(C.c_uint32(max(len(read), 512)), ix.nskip)
        qp = _p(qual, u8p) if qual is not None else None
        err = self.lib.so_collect_hitinfo(h, C.byref(ix), _p(read, u8p), qp,
                                          C.c_uint32(len(read)), int(is_reverse), int(is_short),
                                          C.c_uint32(maxhit_per_tuple), C.c_uint32(maxhit_total),
                                          int(basq))
        if err:
            return err, None, h
        hc = h.contents
        n = hc.n_seeds
        rank = C.c_uint32(0)
        d = dict(
            n_seeds=n, seed_rank=hc.seed_rank, status=hc.status,
            posidx=np.ctypeslib.as_array(hc.posidx, (max(n, 1),))[:n].copy(),
            nhits=np.ctypeslib.as_array(hc.nhits, (max(n, 1),))[:n].copy(),
            qoffs=np.ctypeslib.as_array(hc.qoffs, (max(n, 1),))[:n].copy(),
            sortkey=np.ctypeslib.as_array(hc.sortkey, (max(n, 1),))[:n].copy(),
            sidx=np.ctypeslib.as_array(hc.sidx, (max(n, 1),))[:n].copy(),
            qmask=np.ctypeslib.as_array(hc.qmask, (len(read),)).copy(),
            cover_deficit=self.lib.so_cover_deficit(h, ix.wordlen, ix.nskip),
            nhit_all=self.lib.so_number_of_hits(h, C.c_uint32(maxhit_per_tuple)),
        )
        d["nhit_tot"] = self.lib.so_hit_numbers(h, C.byref(rank))
        d["nhit_rank"] = rank.value
        return 0, d, h

    def hitlist_segment(self, ix, h, lo, hi, nhit_max=10000, use_short=1, hl=None):
        if hl is None:
            hl = self.lib.so_hitlist_create(16384)
        err = self.lib.so_collect_hits_segment(hl, h, C.byref(ix), int(lo), int(hi),
                                               C.c_uint32(nhit_max), int(use_short))
        n = hl.contents.nhits
        dat = np.ctypeslib.as_array(hl.contents.sqdat, (max(n, 1),))[:n].copy()
        return err, dat, hl

    def hitlist_cutoff(self, ix, h, nhit_max=10000, hl=None):
        if hl is None:
            hl = self.lib.so_hitlist_create(16384)
        err = self.lib.so_collect_hits_cutoff(hl, h, C.byref(ix), C.c_uint32(nhit_max))
        n = hl.contents.nhits
        dat = np.ctypeslib.as_array(hl.contents.sqdat, (max(n, 1),))[:n].copy()
        qm = np.ctypeslib.as_array(hl.contents.qmask, (hl.contents.qlen,)).copy()
        return err, dat, qm, hl

    # ---- candidate selection (segment.c) + score replay (rmap.c) ----
    def candidates(self, ix, read, seq_offs, termchar=1, min_cover=0, min_swatscor_below_max=-1, best=False,
                   target_depth=200, max_depth=8000, sensitive=False, nhit_max=10000, maxhit_total=16384,
                   qual=None, basq=0):
        """mapSingleRead between the seed tables and the scoring (rmap.c:1258-1337) on the oracle's own K1:
        -> err, stats dict, list of candidate dicts (segAliCandsCalcSegmentOffsets with edgelen 0)"""
        lib = self.lib
        lib.so_cands_create.restype = C.c_void_p
        lib.so_cands_count.restype = C.c_uint32
        qlen = len(read)
        k, nskip = ix.wordlen, ix.nskip
        mm = self.sc.match - self.sc.mismatch
        min_ktup = (min_cover - k) // nskip if min_cover >= k + nskip else 1
        min_cover = (min_ktup - 1) * nskip + k
        if min_swatscor_below_max < 0:
            below = qlen - 1
        else:
            below = (min_swatscor_below_max // mm) * nskip
            if below < k or best:
                below = k + 2 * (nskip - 1)
        cands = C.c_void_p(lib.so_cands_create())
        lib.so_cands_blank(cands)
        soffs = np.ascontiguousarray(seq_offs, np.uint64)
        nseq = len(soffs) - 1
        hs, cd, nh, nh_tot = [], [], 0, 0
        err = 0
        try:
            for st in (0, 1):
                e, d, h = self.hitinfo(ix, read, qual, st, 1, nhit_max, maxhit_total, basq)
                hs.append(h)
                if e:
                    return e, None, None
                cd.append(d["cover_deficit"])
                nh += d["nhit_rank"]
                nh_tot += d["nhit_tot"]
            for st in (0, 1):
                for s in range(nseq):
                    e, hits, hl = self.hitlist_segment(ix, hs[st], int(soffs[s]), int(soffs[s + 1]), nhit_max, 1)
                    lib.so_hitlist_delete(hl)
                    if e and e != 32:
                        return e, None, None
                    hits = np.ascontiguousarray(hits, np.uint64)
                    err = lib.so_cands_add_list(cands, _p(hits, u64p), len(hits), st, C.c_uint32(qlen), k, nskip, None,
                                                C.c_uint32(min_ktup), C.c_uint32(min_cover), s)
                    if err:
                        return err, None, None
            err = lib.so_cands_stats(cands, C.c_uint32(below), C.c_uint32(cd[0]), C.c_uint32(cd[1]), target_depth,
                                     max_depth, int(sensitive))
            if err:
                return err, None, None
            v = [C.c_uint32(0) for _ in range(4)]
            n = lib.so_cands_count(cands, *[C.byref(x) for x in v])
            stats = dict(n_sort=n, max_cover=v[0].value, max2nd_cover=v[1].value, n_mincover=v[2].value,
                         n_all=v[3].value, cover_deficit=(cd[0], cd[1]), nhit=nh, nhit_tot=nh_tot)
            out = []
            for c in range(n):
                sc = SoCand()
                err = lib.so_cands_offsets(cands, C.c_uint32(c), 0, C.c_uint32(qlen), _p(soffs, u64p), nseq, termchar,
                                           C.byref(sc))
                if err:
                    return err, stats, out
                out.append(dict(qs=sc.qs, qe=sc.qe, rs=sc.rs, re=sc.re, band_l=sc.band_l, band_r=sc.band_r,
                                dqo=sc.dqo, dro=sc.dro, sqidx=sc.sqidx, flags=sc.flags, cover=sc.cover))
            return 0, stats, out
        finally:
            for h in hs:
                lib.so_hitinfo_delete(h)
            lib.so_cands_delete(cands)

    def cigar(self, diffstr, clip_start=0, clip_end=0, softclip=True, xmismatch=False):
        """-> (text bytes | None, nm | error) of so_cigar (oracle/smalt_oracle_cigar.c); error: -1 / -59"""
        d = bytes(diffstr)
        if not d.endswith(b"\0"):
            d += b"\0"
        cap = 6 * len(d) + 64
        out = C.create_string_buffer(cap)
        nm = C.c_int(0)
        self.lib.so_cigar.restype = C.c_int
        n = self.lib.so_cigar(d, int(clip_start), int(clip_end), (2 if softclip else 0) | (4 if xmismatch else 0),
                              out, cap, C.byref(nm))
        if n < 0:
            return None, n
        return out.raw[:n], nm.value

    def score_replay(self, cover, rev, score, band_l, band_r, cover_deficit, qlen, ktup, nskip, min_swatscor,
                     min_swatscor_below_max, best):
        n = len(cover)
        cover = np.ascontiguousarray(cover, np.uint32)
        rev = np.ascontiguousarray(rev, np.uint8)
        score = np.ascontiguousarray(score, np.int32)
        bl = np.ascontiguousarray(band_l, np.int32)
        br = np.ascontiguousarray(band_r, np.int32)
        cdf = np.ascontiguousarray(cover_deficit, np.uint32)
        o = [C.c_int(0) for _ in range(6)]
        align = np.zeros(max(n, 1), np.uint8)
        obl = np.zeros(max(n, 1), np.int32)
        obr = np.zeros(max(n, 1), np.int32)
        err = self.lib.so_score_replay(n, _p(cover, u32p), _p(rev, u8p), _p(score, i32p), _p(bl, i32p), _p(br, i32p),
                                       _p(cdf, u32p), C.c_uint32(qlen), ktup, nskip, self.sc.match, self.sc.mismatch,
                                       -self.sc.gap_init, -self.sc.gap_ext, min_swatscor, min_swatscor_below_max,
                                       int(best), *[C.byref(x) for x in o], _p(align, u8p), _p(obl, i32p), _p(obr, i32p))
        return err, dict(nscored=o[0].value, max1=o[1].value, max2=o[2].value, min_swatscor=o[3].value,
                         scorlen_min=o[4].value, bandwidth_min=o[5].value, align=align[:n], band_l=obl[:n], band_r=obr[:n])

class RefLib:
    """The real reference (oracle/_ref/libsmalt_ref.so)."""

    def __init__(self, pen=DEFAULT_PEN):
        self.lib = C.CDLL(os.path.join(REF_DIR, "libsmalt_ref.so"))
        err = self.lib.refh_init(*pen)
        assert err == 0, err

    def sw_striped(self, read, ref):
        score = C.c_int(0)
        err = self.lib.refh_sw_striped(_p(read, u8p), len(read), _p(ref, u8p), len(ref),
                                       C.byref(score))
        return err, score.value

    def band_fast(self, read, ref, l_edge, r_edge, pl, pr, ul, ur):
        score = C.c_int(0)
        err = self.lib.refh_band_fast(_p(read, u8p), len(read), _p(ref, u8p), len(ref),
                                      l_edge, r_edge, pl, pr, ul, ur, C.byref(score))
        return err, score.value

    def band_align(self, read, ref, l_edge, r_edge, pl, pr, ul, ur, minscore, minscorlen,
                   maxres=256):
        maxdiff = maxres * (len(read) + len(ref) + 8)
        out5, diffbuf, difflen = _align_out(maxres, maxdiff)
        nres = C.c_int(0)
        err = self.lib.refh_band_align(_p(read, u8p), len(read), _p(ref, u8p), len(ref),
                                       l_edge, r_edge, pl, pr, ul, ur, minscore, minscorlen,
                                       maxres, C.byref(nres), _p(out5, i32p), maxdiff,
                                       _p(diffbuf, u8p), _p(difflen, i32p))
        return err, _unpack_results(min(nres.value, maxres), out5, diffbuf, difflen)

    # ---- K1 ----
    def index_load(self, prefix):
        err = self.lib.refh_index_load(prefix.encode())
        assert err == 0, err

    def lookup(self, words):
        words = np.ascontiguousarray(words, np.uint64)
        nh = np.zeros(len(words), np.uint32)
        px = np.zeros(len(words), np.uint32)
        err = self.lib.refh_lookup(_p(words, u64p), len(words), _p(nh, u32p), _p(px, u32p))
        assert err == 0
        return nh, px

    def fetch(self, seqidx, offs, length):
        out = np.zeros(length + 8, np.uint8)
        n = C.c_uint(0)
        err = self.lib.refh_fetch(C.c_longlong(seqidx), C.c_uint(offs), C.c_uint(length),
                                  _p(out, u8p), C.byref(n))
        return err, out[:n.value].copy()

    def hitinfo(self, read, qual, is_reverse, is_short, maxhit_per_tuple=10000,
                maxhit_total=16384, basq=0):
        maxn = len(read) + 8
        a = {k: np.zeros(maxn, np.uint32) for k in ("posidx", "nhits", "qoffs", "sortkey", "sidx")}
        qmask = np.zeros(maxn, np.uint8)
        n_seeds, rank = C.c_uint32(0), C.c_uint32(0)
        status = C.c_uint8(0)
        cd, nr, nt, na = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
        qs = None if qual is None else bytes(qual)
        err = self.lib.refh_hitinfo(_p(read, u8p), len(read), qs, int(is_reverse), int(is_short),
                                    C.c_uint(maxhit_per_tuple), C.c_uint(maxhit_total), int(basq),
                                    maxn, C.byref(n_seeds), C.byref(rank),
                                    _p(a["posidx"], u32p), _p(a["nhits"], u32p),
                                    _p(a["qoffs"], u32p), _p(a["sortkey"], u32p),
                                    _p(a["sidx"], u32p), _p(qmask, u8p), C.byref(status),
                                    C.byref(cd), C.byref(nr), C.byref(nt), C.byref(na))
        if err:
            return err, None
        n = n_seeds.value
        d = {k: v[:n].copy() for k, v in a.items()}
        d.update(n_seeds=n, seed_rank=rank.value, status=status.value,
                 qmask=qmask[:len(read)].copy(), cover_deficit=cd.value,
                 nhit_rank=nr.value, nhit_tot=nt.value, nhit_all=na.value)
        return 0, d

    def hitlist(self, is_reverse, seqidx, nhit_max=10000, use_short=1, qlen=0, maxhits=1 << 20):
        dat = np.zeros(maxhits, np.uint64)
        qm = np.zeros(max(qlen, 1), np.uint8)
        n = C.c_int(0)
        err = self.lib.refh_hitlist(int(is_reverse), C.c_longlong(seqidx), C.c_uint(nhit_max),
                                    int(use_short), maxhits, C.byref(n), _p(dat, u64p),
                                    qlen, qm.ctypes.data_as(C.c_char_p))
        return err, dat[:n.value].copy(), qm

    def cigar(self, diffstr, clip_start=0, clip_end=0, softclip=True, xmismatch=False):
        """the reference's diffStrPrintfStr (diffstr.c:1084-1121, DIFFSTRFORM_CIGEXT = 3 / _XMISMATCH = 4) and
        diffStrGetLevenshteinDistance (:1496) -> (errcode, text bytes, nm)"""
        d = bytes(diffstr)
        if not d.endswith(b"\0"):
            d += b"\0"
        out = C.create_string_buffer(6 * len(d) + 64)
        nchar = C.c_int(0)
        self.lib.diffStrPrintfStr.restype = C.c_int
        self.lib.diffStrGetLevenshteinDistance.restype = C.c_int
        e = self.lib.diffStrPrintfStr(out, C.byref(nchar), d, C.c_char(4 if xmismatch else 3), int(clip_start),
                                      int(clip_end), C.c_char(1 if softclip else 0))
        return e, out.raw[:nchar.value], self.lib.diffStrGetLevenshteinDistance(d)

    def candidates(self, read, qual=None, min_cover=0, min_swatscor_below_max=-1, mismatchdiff=3, best=False,
                   target_depth=200, max_depth=8000, sensitive=False, nhit_max=10000, maxhit_total=16384, basq=0,
                   maxcand=8192):
        """the reference's own segment.c on the reference's own seed tables and hit lists (refh_candidates)"""
        for st in (0, 1):
            e, _ = self.hitinfo(read, qual, st, 1, nhit_max, maxhit_total, basq)
            if e:
                return e, None, None
        stats = np.zeros(8, np.uint32)
        out = np.zeros(11 * maxcand, np.int64)
        err = self.lib.refh_candidates(C.c_uint(nhit_max), C.c_uint(min_cover), int(min_swatscor_below_max),
                                       int(mismatchdiff), int(best), int(target_depth), int(max_depth), int(sensitive),
                                       C.c_uint(len(read)), maxcand, _p(stats, u32p),
                                       out.ctypes.data_as(C.POINTER(C.c_longlong)))
        if err:
            return err, None, None
        st = dict(n_sort=int(stats[0]), n_mincover=int(stats[1]), max_cover=int(stats[2]), max2nd_cover=int(stats[3]),
                  cover_deficit=(int(stats[4]), int(stats[5])), nhit=int(stats[6]), nhit_tot=int(stats[7]))
        keys = ("qs", "qe", "rs", "re", "band_l", "band_r", "dqo", "dro", "sqidx", "flags", "cover")
        cands = [dict(zip(keys, (int(x) for x in out[11 * c:11 * c + 11]))) for c in range(min(st["n_sort"], maxcand))]
        return 0, st, cands
