"""Multi-GPU read sharding: reads are independent, so rank r of N maps one contiguous
range of the input records with its own GPU (index + packed reference replicated per GPU) and
the per-rank SAM texts are concatenated in rank order - the input order, exactly what the
reference's OUTPUT task restores with `-O` (smalt.c:966-1000).  No collective on the data
path; the only communication is the final gather of the SAM text on the host side."""
import ctypes as C

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


def gather_in_order(dist, local_bytes, dst=0):
    """Concatenates the per-rank outputs in rank order on `dst` (None elsewhere).  Works with any
    torch.distributed backend: sizes by all_gather, payloads as uint8 tensors."""
    import torch
    world = dist.get_world_size()
    rank = dist.get_rank()
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    size = torch.tensor([len(local_bytes)], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, size)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes + [1])
    buf = torch.zeros(mx, dtype=torch.uint8, device=dev)
    if local_bytes:
        buf[:len(local_bytes)] = torch.frombuffer(bytearray(local_bytes), dtype=torch.uint8).to(dev)
    bufs = [torch.zeros(mx, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(bufs, buf)
    if rank != dst:
        return None
    return b"".join(bytes(np.asarray(b[:s].cpu().numpy()).tobytes()) for b, s in zip(bufs, sizes))
