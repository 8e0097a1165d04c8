"""ctypes binding of include/smalt_b200.h (the reference-facing C ABI)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

SMB_TASK_READ_REVCOMP = 1
SMB_TASK_REF_PACKED = 2
SMB_ERR_CAPACITY = 103


class SmbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("smalt_b200 error %d: %s" % (code, msg))
        self.code = code


class SwTask(C.Structure):
    _fields_ = [("read_off", C.c_uint64), ("ref_off", C.c_uint64), ("read_len", C.c_uint32),
                ("ref_len", C.c_uint32), ("flags", C.c_uint32), ("reserved", C.c_uint32)]


class BandTask(C.Structure):
    _fields_ = [("read_off", C.c_uint64), ("ref_off", C.c_uint64), ("read_len", C.c_uint32),
                ("ref_len", C.c_uint32), ("flags", C.c_uint32),
                ("l_edge", C.c_int32), ("r_edge", C.c_int32), ("p_left", C.c_int32),
                ("p_right", C.c_int32), ("u_left", C.c_int32), ("u_right", C.c_int32),
                ("minscore", C.c_int32), ("minscorlen", C.c_int32)]


class AliResult(C.Structure):
    _fields_ = [("score", C.c_int32), ("qs", C.c_int32), ("qe", C.c_int32), ("rs", C.c_int32),
                ("re", C.c_int32), ("diff_off", C.c_uint32), ("diff_len", C.c_uint32),
                ("task", C.c_uint32)]


SW_TASK_DTYPE = np.dtype([("read_off", "<u8"), ("ref_off", "<u8"), ("read_len", "<u4"),
                          ("ref_len", "<u4"), ("flags", "<u4"), ("reserved", "<u4")])
BAND_TASK_DTYPE = np.dtype([("read_off", "<u8"), ("ref_off", "<u8"), ("read_len", "<u4"),
                            ("ref_len", "<u4"), ("flags", "<u4"), ("l_edge", "<i4"),
                            ("r_edge", "<i4"), ("p_left", "<i4"), ("p_right", "<i4"),
                            ("u_left", "<i4"), ("u_right", "<i4"), ("minscore", "<i4"),
                            ("minscorlen", "<i4"), ("_pad", "<u4")])
SEED_INFO_DTYPE = np.dtype([("n_seeds", "<u4"), ("seed_rank", "<u4"), ("cover_deficit", "<u4"),
                            ("nhit_rank", "<u4"), ("nhit_tot", "<u4"), ("nhit_all", "<u4"),
                            ("status", "<u4"), ("err", "<i4")])
HIT_REQ_DTYPE = np.dtype([("lo", "<u8"), ("hi", "<u8"), ("read", "<u4"), ("nhit_max", "<u4"),
                          ("strand", "u1"), ("use_short", "u1"), ("reserved", "u1", (2,)), ("nhits_max", "<u4")])
INDEX_SEQ_DTYPE = np.dtype([("start", "<u8"), ("offs", "<u4"), ("n_k", "<u4"), ("tup_base", "<u4"), ("reserved", "<u4")])
INDEX_INFO_DTYPE = np.dtype([("npos", "<u4"), ("nwords", "<u4"), ("nkeys", "<u4"), ("kernel_ms", "<f4")])
ALI_RESULT_DTYPE = np.dtype([("score", "<i4"), ("qs", "<i4"), ("qe", "<i4"), ("rs", "<i4"),
                             ("re", "<i4"), ("diff_off", "<u4"), ("diff_len", "<u4"),
                             ("task", "<u4")])
BLOCK_JOB_DTYPE = np.dtype([("seed_read", "<u4"), ("niv", "<i4"), ("iv_first", "<u4"), ("min_cover", "<u4"),
                            ("min_swatscor", "<i4"), ("reserved", "<u4")])
BLOCK_IVAL_DTYPE = np.dtype([("lo", "<u8"), ("hi", "<u8"), ("seqidx", "<i4"), ("reserved", "<i4")])
BLOCK_PARAMS_DTYPE = np.dtype([("nhit_max", "<u4"), ("min_swatscor_below_max", "<i4"), ("target_depth", "<i4"),
                               ("max_depth", "<i4"), ("best", "u1"), ("sensitive", "u1"), ("termchar", "u1"),
                               ("cigar", "u1")])
BLOCK_READ_DTYPE = np.dtype([("errcode", "<i4"), ("reached_stats", "u1"), ("do_align", "u1"), ("reserved", "u1", (2,)),
                             ("nseg", "<i4"), ("nseg_tot", "<i4"), ("nhit", "<u4"), ("nhit_tot", "<u4"),
                             ("ncand", "<u4"), ("nscored", "<u4"), ("max1scor", "<i4"), ("max2scor", "<i4"),
                             ("min_swatscor", "<i4"), ("scorlen_min", "<i4"), ("bandwidth_min", "<i4"),
                             ("k3_first", "<u4"), ("nk3", "<u4"), ("reserved2", "<u4")])
BLOCK_CAND_DTYPE = np.dtype([("rs", "<u8"), ("sqidx", "<i4"), ("swscor", "<i4"), ("reflen", "<u4"), ("band_l", "<i4"),
                             ("band_r", "<i4"), ("reverse", "u1"), ("reserved", "u1", (3,))])
BLOCK_SIZES_DTYPE = np.dtype([(k, "<u8") for k in ("nhits", "ncand", "nk2", "nk2_band", "nk3", "nresults", "ndiffbytes",
                                                    "k2_cells", "k2_cells_ref", "k2_tasks_ref", "k3_cells")] +
                             [(k, "<f4") for k in ("ms_hits", "ms_cand", "ms_k2", "ms_k3")] +
                             [("launches", "<i4"), ("reserved", "<i4"), ("ncigarbytes", "<u8")])
CIGAR_ON, CIGAR_SOFTCLIP, CIGAR_XMISMATCH = 1, 2, 4
assert BLOCK_READ_DTYPE.itemsize == 64 and BLOCK_CAND_DTYPE.itemsize == 32 and BLOCK_JOB_DTYPE.itemsize == 24
assert SW_TASK_DTYPE.itemsize == C.sizeof(SwTask)
assert BAND_TASK_DTYPE.itemsize == C.sizeof(BandTask), (BAND_TASK_DTYPE.itemsize, C.sizeof(BandTask))
assert ALI_RESULT_DTYPE.itemsize == C.sizeof(AliResult)


def lib_path():
    return os.path.join(_HERE, "libsmalt_b200.so")


_lib = None


def load_library():
    """Loads libsmalt_b200.so; raises if it is missing - the CUDA library IS the product."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise ImportError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C smalt_b200/csrc` (no CPU fallback exists)" % p)
    lib = C.CDLL(p)
    lib.smb_version.restype = C.c_char_p
    lib.smb_last_error.restype = C.c_char_p
    lib.smb_last_error.argtypes = [C.c_void_p]
    lib.smb_last_kernel_ms.restype = C.c_float
    lib.smb_last_kernel_ms.argtypes = [C.c_void_p]
    lib.smb_last_kernel_launches.argtypes = [C.c_void_p]
    lib.smb_total_kernel_launches.restype = C.c_longlong
    lib.smb_total_kernel_launches.argtypes = [C.c_void_p]
    lib.smb_ctx_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    lib.smb_ctx_destroy.argtypes = [C.c_void_p]
    lib.smb_ctx_destroy.restype = None
    lib.smb_set_scoring.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.smb_int_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    lib.smb_arena_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.smb_refseq_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_void_p, C.c_int]
    lib.smb_sw_score_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.smb_band_score_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.smb_band_align_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t,
                                         C.POINTER(C.c_size_t), C.c_void_p, C.c_void_p, C.c_size_t,
                                         C.POINTER(C.c_size_t), C.c_void_p, C.POINTER(C.c_uint64)]
    lib.smb_index_upload.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32,
                                     C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.smb_seed_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_uint32,
                                   C.c_uint32, C.c_int, C.c_int] + [C.c_void_p] * 7
    lib.smb_hits_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.c_size_t,
                                   C.POINTER(C.c_size_t), C.c_void_p, C.c_void_p]
    lib.smb_hits_qmask.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.smb_index_build.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                    C.c_void_p]
    lib.smb_index_fetch.argtypes = [C.c_void_p] * 5
    lib.smb_seed_batch_tables.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int,
                                          C.c_void_p]
    lib.smb_block_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    lib.smb_block_fetch.argtypes = [C.c_void_p] * 7
    lib.smb_block_debug_cands.argtypes = [C.c_void_p] * 5 + [C.c_size_t]
    lib.smb_block_fetch_cigar.argtypes = [C.c_void_p] * 8
    lib.smb_cigar_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                    C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    _lib = lib
    return lib


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """One GPU context (one per process / per GPU): smb_ctx of include/smalt_b200.h."""

    def __init__(self, device=0, penalties=(1, -2, -4, -3)):
        self.lib = load_library()
        self._h = C.c_void_p()
        rc = self.lib.smb_ctx_create(C.byref(self._h), device)
        if rc:
            raise SmbError(rc, "smb_ctx_create failed (no CUDA device? there is no CPU fallback)")
        self._check(self.lib.smb_set_scoring(self._h, *penalties))
        self._keep = []

    def close(self):
        if self._h:
            self.lib.smb_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise SmbError(rc, self.lib.smb_last_error(self._h).decode())

    @property
    def last_kernel_ms(self):
        return float(self.lib.smb_last_kernel_ms(self._h))

    @property
    def last_kernel_launches(self):
        return int(self.lib.smb_last_kernel_launches(self._h))

    @property
    def total_kernel_launches(self):
        return int(self.lib.smb_total_kernel_launches(self._h))

    def int_peak(self):
        """-> (viaddmax, vimax3, add+max, viaddmax_s16x2, vimax3_s16x2) giga thread-instructions/s
        measured on this device"""
        g = (C.c_double * 5)()
        self._check(self.lib.smb_int_peak(self._h, g))
        return tuple(g)

    def arena_upload(self, codes):
        codes = np.ascontiguousarray(codes, np.uint8)
        self._check(self.lib.smb_arena_upload(self._h, _vp(codes), codes.size))

    def refseq_upload(self, words, nbases, seq_offs):
        words = np.ascontiguousarray(words, np.uint32)
        so = np.ascontiguousarray(seq_offs, np.uint64)
        self._check(self.lib.smb_refseq_upload(self._h, _vp(words), words.size, int(nbases), _vp(so),
                                               len(so) - 1))

    def index_upload(self, ix):
        """ix: dict as returned by smifile.read_smi / indexer.as_loaded(build_index(...))"""
        z = np.zeros(1, np.uint32)
        widx = ix["wordidx"] if ix["wordidx"] is not None else z
        pidx = ix["posidx"] if ix["posidx"] is not None else z
        self._check(self.lib.smb_index_upload(
            self._h, ix["typ"], ix["wordlen"], ix["nskip"], ix["nbits_key"], ix["nbits_lo"], ix["npos"],
            ix["nwords"], _vp(np.ascontiguousarray(ix["idx"], np.uint32)),
            _vp(np.ascontiguousarray(ix["pos"], np.uint32)), _vp(np.ascontiguousarray(widx, np.uint32)),
            _vp(np.ascontiguousarray(pidx, np.uint32))))

    def index_build(self, k, nskip, typ, nbits_key, nbits_lo, grid):
        """GPU index construction over the uploaded packed reference; grid: INDEX_SEQ_DTYPE array.
        -> dict(npos, nwords, nkeys, kernel_ms, idx, pos, wordidx, posidx)"""
        grid = np.ascontiguousarray(grid, INDEX_SEQ_DTYPE)
        info = np.zeros(1, INDEX_INFO_DTYPE)
        self._check(self.lib.smb_index_build(self._h, k, nskip, typ, nbits_key, nbits_lo, _vp(grid), len(grid), _vp(info)))
        npos, nwords, nkeys = int(info["npos"][0]), int(info["nwords"][0]), int(info["nkeys"][0])
        idx = np.zeros(nkeys + 1, np.uint32)
        pos = np.zeros(max(npos, 1), np.uint32)
        widx = np.zeros(nwords + 1, np.uint32)
        pidx = np.zeros(nwords + 1, np.uint32)
        self._check(self.lib.smb_index_fetch(self._h, _vp(idx), _vp(pos), _vp(widx), _vp(pidx)))
        return dict(npos=npos, nwords=nwords, nkeys=nkeys, kernel_ms=float(info["kernel_ms"][0]), idx=idx,
                    pos=pos[:npos], wordidx=widx if typ else None, posidx=pidx if typ else None)

    def seed_batch(self, read_off, read_len, qual=None, maxhit_per_tuple=10000, maxhit_total=16384,
                   basq_thresh=0, full=True, short_info=True):
        """-> (info[2*n] SEED_INFO_DTYPE, tables dict or None).  Table arrays have 2*sum(read_len)
        slots; read r strand s starts at 2*sum(read_len[:r]) + s*read_len[r]."""
        read_off = np.ascontiguousarray(read_off, np.uint64)
        read_len = np.ascontiguousarray(read_len, np.uint32)
        n = len(read_len)
        nslots = 2 * int(read_len.astype(np.int64).sum())
        info = np.zeros(2 * n, SEED_INFO_DTYPE)
        tabs = None
        ptrs = [None] * 6
        if full:
            tabs = {k: np.zeros(nslots, np.uint32) for k in ("posidx", "nhits", "qoffs", "sortkey", "sidx")}
            tabs["qmask"] = np.zeros(nslots, np.uint8)
            ptrs = [_vp(tabs[k]) for k in ("posidx", "nhits", "qoffs", "sortkey", "sidx", "qmask")]
        q = None if qual is None else _vp(np.ascontiguousarray(qual, np.uint8))
        self._check(self.lib.smb_seed_batch(self._h, _vp(read_off), _vp(read_len), n, q, maxhit_per_tuple,
                                            maxhit_total, basq_thresh, int(short_info), _vp(info), *ptrs))
        return info, tabs

    def seed_batch_tables(self, tables, read_table, read_off, read_len, qual=None, maxhit_per_tuple=0,
                          maxhit_total=0, basq_thresh=0, short_info=False):
        """Seed tables against per-read small perfect-hash indexes (dicts as made by indexer.build_index with
        typ 0, all of one k / nskip): read r uses tables[read_table[r]].  -> info[2*n]"""
        class Small(C.Structure):
            _fields_ = [("idx", C.c_void_p), ("pos", C.c_void_p), ("npos", C.c_uint32)]
        k, s = tables[0]["wordlen"], tables[0]["nskip"]
        keep, arr = [], (Small * len(tables))()
        for i, t in enumerate(tables):
            assert t["typ"] == 0 and t["wordlen"] == k and t["nskip"] == s
            idx = np.ascontiguousarray(t["idx"], np.uint32)
            pos = np.ascontiguousarray(t["pos"], np.uint32)
            keep += [idx, pos]
            arr[i].idx, arr[i].pos, arr[i].npos = idx.ctypes.data, pos.ctypes.data, int(t["npos"])
        read_table = np.ascontiguousarray(read_table, np.uint32)
        read_off = np.ascontiguousarray(read_off, np.uint64)
        read_len = np.ascontiguousarray(read_len, np.uint32)
        n = len(read_len)
        info = np.zeros(2 * n, SEED_INFO_DTYPE)
        q = None if qual is None else _vp(np.ascontiguousarray(qual, np.uint8))
        self._check(self.lib.smb_seed_batch_tables(self._h, k, s, C.cast(arr, C.c_void_p), len(tables), _vp(read_table),
                                                   _vp(read_off), _vp(read_len), n, q, maxhit_per_tuple, maxhit_total,
                                                   basq_thresh, int(short_info), _vp(info)))
        return info

    def hits_batch(self, req, nhits_alloc=0, max_hits=None):
        """-> (sqdat uint64, list_first[nreq+1], errs) for HIT_REQ_DTYPE requests"""
        req = np.ascontiguousarray(req, HIT_REQ_DTYPE)
        n = len(req)
        if max_hits is None:
            max_hits = 64 * n + 4096
        while True:
            sq = np.zeros(max_hits, np.uint64)
            first = np.zeros(n + 1, np.uint64)
            errs = np.zeros(n, np.int32)
            tot = C.c_size_t(0)
            rc = self.lib.smb_hits_batch(self._h, _vp(req), n, nhits_alloc, _vp(sq), max_hits, C.byref(tot),
                                         _vp(first), _vp(errs))
            if rc == SMB_ERR_CAPACITY and tot.value > max_hits:
                max_hits = tot.value
                continue
            self._check(rc)
            return sq[:tot.value], first, errs

    def hits_qmask(self, nreq, nbytes):
        """HITQUAL masks of the lists of the last hits_batch -> (qmask bytes, first[nreq+1])"""
        qm = np.zeros(max(1, nbytes), np.uint8)
        first = np.zeros(nreq + 1, np.uint64)
        self._check(self.lib.smb_hits_qmask(self._h, _vp(qm), nbytes, _vp(first)))
        return qm[:int(first[nreq])], first

    def sw_score(self, tasks):
        tasks = np.ascontiguousarray(tasks, SW_TASK_DTYPE)
        n = len(tasks)
        scores = np.zeros(n, np.int32)
        errs = np.zeros(n, np.int32)
        self._check(self.lib.smb_sw_score_batch(self._h, _vp(tasks), n, _vp(scores), _vp(errs)))
        return scores, errs

    def band_score(self, tasks):
        tasks = np.ascontiguousarray(tasks, BAND_TASK_DTYPE)
        n = len(tasks)
        scores = np.zeros(n, np.int32)
        errs = np.zeros(n, np.int32)
        self._check(self.lib.smb_band_score_batch(self._h, _vp(tasks), n, _vp(scores), _vp(errs)))
        return scores, errs

    def band_align(self, tasks, max_results=None, max_diff=None):
        """-> (results[ALI_RESULT_DTYPE], first_result[n+1], diffstr bytes, errs, ncells)"""
        tasks = np.ascontiguousarray(tasks, BAND_TASK_DTYPE)
        n = len(tasks)
        if max_results is None:
            max_results = 4 * n + 64
        if max_diff is None:
            max_diff = int(tasks["read_len"].sum() // 2 + 64 * n + 4096)
        while True:
            res = np.zeros(max_results, ALI_RESULT_DTYPE)
            first = np.zeros(n + 1, np.uint32)
            diff = np.zeros(max_diff, np.uint8)
            errs = np.zeros(n, np.int32)
            nres, ndiff, cells = C.c_size_t(0), C.c_size_t(0), C.c_uint64(0)
            rc = self.lib.smb_band_align_batch(self._h, _vp(tasks), n, _vp(res), max_results,
                                               C.byref(nres), _vp(first), _vp(diff), max_diff,
                                               C.byref(ndiff), _vp(errs), C.byref(cells))
            if rc == SMB_ERR_CAPACITY and (nres.value > max_results or ndiff.value > max_diff):
                max_results = max(max_results, nres.value)
                max_diff = max(max_diff, ndiff.value)
                continue
            self._check(rc)
            return res[:nres.value], first, diff[:ndiff.value], errs, cells.value


    # ---- resident block (hit lists -> candidates -> K2 -> replay -> K3 on the device) ----
    def block_run(self, jobs, ivals=None, nhit_max=10000, min_swatscor_below_max=-1, target_depth=200, max_depth=8000,
                  best=False, sensitive=False, termchar=False, cigar=0):
        """-> sizes (BLOCK_SIZES_DTYPE scalar) of the block run on the last seed batch; cigar = CIGAR_* flags"""
        jobs = np.ascontiguousarray(jobs, BLOCK_JOB_DTYPE)
        ivals = np.zeros(0, BLOCK_IVAL_DTYPE) if ivals is None else np.ascontiguousarray(ivals, BLOCK_IVAL_DTYPE)
        prm = np.zeros(1, BLOCK_PARAMS_DTYPE)
        prm[0] = (nhit_max, min_swatscor_below_max, target_depth, max_depth, int(best), int(sensitive), int(termchar), int(cigar))
        sizes = np.zeros(1, BLOCK_SIZES_DTYPE)
        self._check(self.lib.smb_block_run(self._h, _vp(prm), _vp(jobs), len(jobs), _vp(ivals), len(ivals), _vp(sizes)))
        self._block_n = (len(jobs), sizes[0])
        return sizes[0]

    def block_fetch(self):
        """-> (reads[BLOCK_READ_DTYPE], cands[BLOCK_CAND_DTYPE], errs, first_result, results, diffstr)"""
        n, sz = self._block_n
        nk3, nres, nd = int(sz["nk3"]), int(sz["nresults"]), int(sz["ndiffbytes"])
        reads = np.zeros(n, BLOCK_READ_DTYPE)
        cands = np.zeros(nk3, BLOCK_CAND_DTYPE)
        errs = np.zeros(nk3, np.int32)
        first = np.zeros(nk3 + 1, np.uint32)
        res = np.zeros(nres, ALI_RESULT_DTYPE)
        diff = np.zeros(max(nd, 1), np.uint8)
        self._check(self.lib.smb_block_fetch(self._h, _vp(reads), _vp(cands), _vp(errs), _vp(first), _vp(res), _vp(diff)))
        return reads, cands, errs, first, res, diff[:nd]

    def block_fetch_cigar(self):
        """block_fetch() + (cigar_first[nresults + 1], nm[nresults], text bytes) of the output stage"""
        n, sz = self._block_n
        nk3, nres, nd, nc = int(sz["nk3"]), int(sz["nresults"]), int(sz["ndiffbytes"]), int(sz["ncigarbytes"])
        reads = np.zeros(n, BLOCK_READ_DTYPE)
        cands = np.zeros(nk3, BLOCK_CAND_DTYPE)
        errs = np.zeros(nk3, np.int32)
        first = np.zeros(nk3 + 1, np.uint32)
        res = np.zeros(nres, ALI_RESULT_DTYPE)
        diff = np.zeros(max(nd, 1), np.uint8)
        blob = np.zeros((2 * nres + 1) * 4 + nc + 8, np.uint8)
        self._check(self.lib.smb_block_fetch_cigar(self._h, _vp(reads), _vp(cands), _vp(errs), _vp(first), _vp(res),
                                                   _vp(diff), _vp(blob)))
        cfirst, nm, text = _split_cigar_blob(blob, nres, nc)
        return reads, cands, errs, first, res, diff[:nd], cfirst, nm, text

    def cigar_batch(self, diffstr, diff_off, clip_start, clip_end, flags=0):
        """CIGAR text + edit distance of explicit alignment strings -> (cigar_first[n + 1], nm[n], text bytes)"""
        diffstr = np.ascontiguousarray(diffstr, np.uint8)
        diff_off = np.ascontiguousarray(diff_off, np.uint32)
        cs = np.ascontiguousarray(clip_start, np.uint32)
        ce = np.ascontiguousarray(clip_end, np.uint32)
        n = len(diff_off)
        cap = 64
        while True:
            blob = np.zeros((2 * n + 1) * 4 + cap + 8, np.uint8)
            nt = C.c_size_t(0)
            rc = self.lib.smb_cigar_batch(self._h, _vp(diffstr), diffstr.size, _vp(diff_off), _vp(cs), _vp(ce), n,
                                          int(flags), _vp(blob), cap, C.byref(nt))
            if rc == SMB_ERR_CAPACITY and nt.value > cap:
                cap = nt.value
                continue
            self._check(rc)
            return _split_cigar_blob(blob, n, nt.value)

    def block_debug_cands(self):
        """-> (cand_first[n+1], cands[BLOCK_CAND_DTYPE] with swscor = K2 score, cover, qs_qe[n, 2]) of ALL candidates"""
        n, sz = self._block_n
        nc = int(sz["ncand"])
        first = np.zeros(n + 1, np.uint64)
        cands = np.zeros(nc, BLOCK_CAND_DTYPE)
        cover = np.zeros(nc, np.uint32)
        qsqe = np.zeros((nc, 2), np.uint32)
        self._check(self.lib.smb_block_debug_cands(self._h, _vp(first), _vp(cands), _vp(cover), _vp(qsqe), nc))
        return first, cands, cover, qsqe


def _split_cigar_blob(blob, n, ntext):
    """the blob of the output stage (SMB_CIGAR_BLOB_BYTES) -> (first[n + 1], nm[n], text bytes)"""
    first = blob[:(n + 1) * 4].view(np.uint32).copy()
    nm = blob[(n + 1) * 4:(2 * n + 1) * 4].view(np.int32).copy()
    text = blob[(2 * n + 1) * 4:(2 * n + 1) * 4 + ntext].tobytes()
    return first, nm, text


def pack_sequences(seqs):
    """Concatenates code arrays into one arena; returns (arena, offsets)."""
    offs = np.zeros(len(seqs) + 1, np.uint64)
    if seqs:
        offs[1:] = np.cumsum([len(s) for s in seqs])
    arena = np.concatenate(seqs).astype(np.uint8) if seqs else np.zeros(0, np.uint8)
    return arena, offs
