"""Multi-GPU read sharding: reads are independent, so rank r of N maps one contiguous
range of the input records with its own GPU (index + packed reference replicated per GPU) and
the per-rank SAM texts are merged in rank order - the input order, what the reference's OUTPUT
task restores with `-O` (smalt.c:966-1000) - by offset writes into one file (merge_to_file).  No
collective on the data path; the ranks exchange the lengths of their texts (8 bytes each).
The merged text equals a single-process run except where the reference draws among equally good
placements: every rank starts its own drand48 sequence (results.c:2298), like every run of the
reference with another seed."""
import ctypes as C
import os

import numpy as np


def split_points(text, nparts, lib=None):
    """Byte offsets cutting FASTQ/FASTA `text` into `nparts` contiguous ranges of whole records
    (the same boundary search the in-process driver uses for its blocks)."""
    from .mapper import load_map_library
    lib = lib or load_map_library()
    n = len(text)
    if nparts <= 1 or n == 0:
        return [0, n]
    chunk = max(1, (n + nparts - 1) // nparts)
    starts = (C.c_size_t * (nparts + 8))()
    ns, nrec = C.c_size_t(0), C.c_size_t(0)
    lib.smbm_split_blocks.argtypes = [C.c_char_p, C.c_size_t, C.c_size_t, C.POINTER(C.c_size_t), C.c_size_t,
                                      C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    rc = lib.smbm_split_blocks(text, n, chunk, starts, nparts + 8, C.byref(ns), C.byref(nrec))
    if rc:
        raise ValueError("read text is not whole FASTA / 4-line FASTQ records (error %d)" % rc)
    pts = [int(starts[i]) for i in range(ns.value)]
    pts += [n] * (nparts + 1 - len(pts))
    pts[0] = 0
    return pts[:nparts] + [n]


def shard_of(text, rank, world, lib=None):
    """the records rank `rank` of `world` maps"""
    pts = split_points(text, world, lib)
    return text[pts[rank]:pts[rank + 1]]


def pair_split_points(text1, text2, nparts):
    """Paired input: both 4-line FASTQ texts are cut at the same RECORD numbers (record i of one file
    is the mate of record i of the other) -> (offsets in text1, offsets in text2), nparts+1 each."""
    def line_starts(t):
        a = np.frombuffer(t, np.uint8)
        nl = np.flatnonzero(a == 10)
        starts = np.concatenate([[0], nl + 1])
        if len(t) and t[-1:] != b"\n":
            nlines = len(nl) + 1
        else:
            nlines = len(nl)
        return starts, nlines
    s1, n1 = line_starts(text1)
    s2, n2 = line_starts(text2)
    if n1 != n2 or n1 % 4:
        raise ValueError("mate files are not two 4-line FASTQ texts with the same number of records")
    nrec = n1 // 4
    cuts = [min(nrec, (nrec * r + nparts - 1) // nparts) for r in range(nparts + 1)]
    cuts[-1] = nrec
    def offs(starts, n, t):
        return [int(starts[4 * c]) if 4 * c < len(starts) and c < nrec else len(t) for c in cuts]
    return offs(s1, n1, text1), offs(s2, n2, text2)


def pair_shard_of(text1, text2, rank, world):
    """the pairs rank `rank` of `world` maps: (records of file 1, records of file 2)"""
    p1, p2 = pair_split_points(text1, text2, world)
    return text1[p1[rank]:p1[rank + 1]], text2[p2[rank]:p2[rank + 1]]


_MERGE_MAPS = {}   # path -> (fd, mmap, size): the shared mapping of a merge target is kept between merges


def _merge_map(path, need):
    """shared read-write mapping of `path` with room for `need` bytes (grown in 256 MB steps, kept open)"""
    import mmap
    ent = _MERGE_MAPS.get(path)
    if ent is not None and ent[2] >= need:
        return ent[1]
    if ent is not None:
        ent[1].close()
        os.close(ent[0])
    size = max(1 << 20, (need + (1 << 28) - 1) >> 28 << 28)
    fd = os.open(path, os.O_RDWR | os.O_CREAT, 0o644)
    if os.fstat(fd).st_size < size:
        os.ftruncate(fd, size)       # sparse: pages appear when written
    mm = mmap.mmap(fd, size, mmap.MAP_SHARED, mmap.PROT_READ | mmap.PROT_WRITE)
    _MERGE_MAPS[path] = (fd, mm, size)
    return mm


def merge_to_file(dist, local_bytes, path, header=b"", writers=8, final_size=True):
    """Host merge of the per-rank outputs in rank (= input) order, as the reference's OUTPUT task
    orders its blocks by read number (smalt.c:966-1000): every rank copies its text to its offset
    in ONE file - `header` (rank 0), then rank 0's records, rank 1's, ... - through a shared mapping of
    the file; the offsets are the exclusive scan of the text lengths, the only thing the ranks
    exchange (8 bytes each).  Put `path` on /dev/shm for a merge through shared memory; the mapping is
    kept between calls, so that repeated merges into the same path are plain memory copies shared by
    `writers` threads.  final_size: rank 0 cuts the file to the merged length (after the barrier).
    Ends with a barrier; -> total bytes."""
    import ctypes
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    mine = len(local_bytes) + (len(header) if rank == 0 else 0)
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    sizes[rank] = mine
    dist.all_reduce(sizes)
    sizes = [int(x) for x in sizes.cpu().tolist()]
    total = sum(sizes)
    off = sum(sizes[:rank])
    mm = _merge_map(path, total)
    dst = (ctypes.c_char * len(mm)).from_buffer(mm)
    base = ctypes.addressof(dst)
    if rank == 0 and header:
        ctypes.memmove(base, header, len(header))
        off = len(header)
    mv = memoryview(local_bytes).cast("B")
    n = len(mv)
    if n:
        src = (ctypes.c_char * n).from_buffer_copy(mv) if mv.readonly and not isinstance(local_bytes, (bytes, bytearray)) else None
        if isinstance(local_bytes, bytes):
            src_addr = ctypes.cast(ctypes.c_char_p(local_bytes), ctypes.c_void_p).value
        elif src is not None:
            src_addr = ctypes.addressof(src)
        else:
            src_addr = ctypes.addressof((ctypes.c_char * n).from_buffer(mv))
        nthr = max(1, min(writers, n >> 24))     # a writer per 16 MB: the copy is the whole cost
        if nthr == 1:
            ctypes.memmove(base + off, src_addr, n)
        else:
            import threading
            step = (n + nthr - 1) // nthr
            ths = [threading.Thread(target=ctypes.memmove, args=(base + off + k * step, src_addr + k * step,
                                                                 min(step, n - k * step))) for k in range(nthr)]
            for t in ths:
                t.start()
            for t in ths:
                t.join()
    del dst
    dist.barrier()
    if final_size and rank == 0:
        os.truncate(path, total)
        ent = _MERGE_MAPS.pop(path)
        ent[1].close()
        os.close(ent[0])
    elif final_size:
        ent = _MERGE_MAPS.pop(path)
        ent[1].close()
        os.close(ent[0])
    return total


def every_nth_record(text, n, phase=0):
    """records phase, phase + n, phase + 2n, ... of a 4-line FASTQ text"""
    a = np.frombuffer(text, np.uint8)
    nl = np.flatnonzero(a == 10)
    starts = np.concatenate([[0], nl + 1])
    nrec = (len(nl) + (0 if not len(text) or text[-1:] == b"\n" else 1)) // 4
    out = []
    for r in range(phase, nrec, n):
        b = int(starts[4 * r])
        e = int(starts[4 * r + 4]) if 4 * r + 4 < len(starts) else len(text)
        out.append(text[b:e])
    return b"".join(out)


def sample_insert_sizes(dist, exe, index_prefix, text1, text2, skip, hist_path, threads, env=None):
    """Insert-size estimation for sharded pairs.  The reference samples every skip-th pair of the input and
    builds ONE histogram from the sampled insert sizes (smalt.c:838-878, :1288-1300: the only cross-read state
    of the path).  Here every rank contributes every skip-th pair of its shard, the subsamples are merged in
    rank order on the host (merge_to_file) and rank 0 runs `sample -u 1` on them; all ranks then read the
    same histogram file.  -> hist_path"""
    import subprocess
    rank = dist.get_rank()
    f1, f2 = hist_path + ".s1.fq", hist_path + ".s2.fq"
    merge_to_file(dist, every_nth_record(text1, skip), f1)
    merge_to_file(dist, every_nth_record(text2, skip), f2)
    err = b""
    if rank == 0:
        r = subprocess.run([exe, "sample", "-u", "1", "-n", str(threads), "-o", hist_path, index_prefix, f1, f2],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
        if r.returncode != 0:
            err = r.stderr[-300:]
        os.unlink(f1)
        os.unlink(f2)
    dist.barrier()
    if not os.path.exists(hist_path):
        raise RuntimeError("sample failed on rank 0: " + err.decode(errors="replace"))
    return hist_path
