"""ctypes binding of include/smalt_b200_map.h: the in-process `smalt_b200 map` driver
(libsmalt_b200_map.so = the reference's unmodified driver / candidate selection / results /
SAM writer objects around the B200 hot path of libsmalt_b200.so)."""
import ctypes as C
import os

from .capi import SmbError, load_library

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


class MapStats(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("wall_s", C.c_double),
                ("k1_ms", C.c_double), ("k2_ms", C.c_double), ("k3_ms", C.c_double),
                ("k2_tasks", C.c_uint64), ("k2_cells", C.c_uint64),
                ("k3_tasks", C.c_uint64), ("k3_cells", C.c_uint64),
                ("gpu_launches", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("host_stage_s", C.c_double * 12), ("host_cpu_s", C.c_double * 8), ("cand_ms", C.c_double),
                ("cigar_dev", C.c_uint64), ("cigar_host", C.c_uint64)]

    STAGES = ("staging", "seed", "hits", "candidates", "score", "replay", "align", "results", "parse",
              "results.add", "results.sort_filter", "results.emit")

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k not in ("host_stage_s", "host_cpu_s")}
        d["host_cpu_s"] = {k: self.host_cpu_s[i] for i, k in enumerate(self.STAGES[:8])}
        d["host_stage_s"] = {k: self.host_stage_s[i] for i, k in enumerate(self.STAGES)}
        return d


def map_lib_path():
    return os.path.join(_HERE, "libsmalt_b200_map.so")


def load_map_library():
    """Loads libsmalt_b200_map.so (and libsmalt_b200.so); raises if missing - no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    load_library()
    p = map_lib_path()
    if not os.path.exists(p):
        raise ImportError("%s not built: `make -C smalt_b200/hostc` needs the reference tree; on the GPU box "
                          "the prebuilt library travels with the repo snapshot" % p)
    lib = C.CDLL(p)
    lib.smbm_open.argtypes = [C.POINTER(C.c_void_p), C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_char_p)]
    lib.smbm_open_paired.argtypes = lib.smbm_open.argtypes
    lib.smbm_map_fastq_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                         C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(MapStats)]
    lib.smbm_map_fastq.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p),
                                   C.POINTER(C.c_size_t), C.POINTER(MapStats)]
    lib.smbm_sam_header.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    lib.smbm_free.argtypes = [C.c_void_p]
    lib.smbm_free.restype = None
    lib.smbm_close.argtypes = [C.c_void_p]
    _lib = lib
    return lib


def _as_pointer(text):
    """bytes / bytearray / memoryview / numpy uint8 array -> (void pointer, length, object to keep alive)"""
    if isinstance(text, bytes):
        return C.cast(C.c_char_p(text), C.c_void_p), len(text), text
    mv = memoryview(text).cast("B")
    if mv.readonly or not mv.contiguous:
        data = bytes(mv)
        return C.cast(C.c_char_p(data), C.c_void_p), len(data), data
    arr = (C.c_char * len(mv)).from_buffer(mv)
    return C.cast(arr, C.c_void_p), len(mv), (arr, mv)


class Mapper:
    """`smalt map -n nthreads -O [options] index_prefix` held open in this process."""

    def __init__(self, index_prefix, nthreads=1, options=(), paired=False):
        self.lib = load_map_library()
        self._h = C.c_void_p()
        opts = (C.c_char_p * max(1, len(options)))(*[o.encode() for o in options])
        openf = self.lib.smbm_open_paired if paired else self.lib.smbm_open
        rc = openf(C.byref(self._h), index_prefix.encode(), int(nthreads), len(options), opts)
        if rc:
            raise SmbError(rc, "smbm_open(%s) failed (index missing, unsupported option or no CUDA device; "
                               "there is no CPU fallback)" % index_prefix)
        self.stats = MapStats()

    def map_fastq(self, text):
        """text: bytes-like FASTQ/FASTA -> SAM records (bytes, no header) in input order."""
        p, nbytes, keep = _as_pointer(text)
        sam = C.c_void_p()
        n = C.c_size_t(0)
        rc = self.lib.smbm_map_fastq(self._h, p, nbytes, C.byref(sam), C.byref(n), C.byref(self.stats))
        del keep
        if rc:
            raise SmbError(rc, "smbm_map_fastq failed")
        return C.string_at(sam, n.value)

    def map_fastq_view(self, text, mates=None):
        """like map_fastq / map_fastq_pairs, but returns a memoryview of the SAM records in the mapper's own
        buffer (valid until the next call): no copy on the Python side"""
        sam = C.c_void_p()
        n = C.c_size_t(0)
        p, nbytes, keep = _as_pointer(text)
        if mates is None:
            rc = self.lib.smbm_map_fastq(self._h, p, nbytes, C.byref(sam), C.byref(n), C.byref(self.stats))
        else:
            p2, nbytes2, keep2 = _as_pointer(mates)
            rc = self.lib.smbm_map_fastq_pairs(self._h, p, nbytes, p2, nbytes2, C.byref(sam), C.byref(n), C.byref(self.stats))
            del keep2
        del keep
        if rc:
            raise SmbError(rc, "smbm_map_fastq failed")
        if not n.value:
            return memoryview(b"")
        return memoryview((C.c_char * n.value).from_address(sam.value)).cast("B")

    def map_fastq_nocopy(self, text):
        """like map_fastq but returns only the length of the SAM text (bench: no Python copy)"""
        sam = C.c_void_p()
        n = C.c_size_t(0)
        rc = self.lib.smbm_map_fastq(self._h, C.cast(C.c_char_p(text), C.c_void_p), len(text), C.byref(sam),
                                     C.byref(n), C.byref(self.stats))
        if rc:
            raise SmbError(rc, "smbm_map_fastq failed")
        return n.value

    def map_fastq_pairs(self, text, text_mates, copy=True):
        """record i of `text` and record i of `text_mates` are a pair (mapper opened with paired=True)
        -> SAM records of both mates of every pair (bytes), or their total length if not copy"""
        sam = C.c_void_p()
        n = C.c_size_t(0)
        rc = self.lib.smbm_map_fastq_pairs(self._h, C.cast(C.c_char_p(text), C.c_void_p), len(text),
                                           C.cast(C.c_char_p(text_mates), C.c_void_p), len(text_mates),
                                           C.byref(sam), C.byref(n), C.byref(self.stats))
        if rc:
            raise SmbError(rc, "smbm_map_fastq_pairs failed")
        return C.string_at(sam, n.value) if copy else n.value

    def sam_header(self):
        t = C.c_void_p()
        n = C.c_size_t(0)
        rc = self.lib.smbm_sam_header(self._h, C.byref(t), C.byref(n))
        if rc:
            raise SmbError(rc, "smbm_sam_header failed")
        s = C.string_at(t, n.value)
        self.lib.smbm_free(t)
        return s

    def close(self):
        if self._h:
            self.lib.smbm_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
