"""Host-side logic of the block-parallel / multi-GPU path, no GPU: record-boundary search of the
driver (smbm_split_blocks) and the rank-ordered gather, the latter with world_size 2 over gloo."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle_lib import ROOT

MAPLIB = os.path.join(ROOT, "smalt_b200", "libsmalt_b200_map.so")
pytestmark = pytest.mark.skipif(not os.path.exists(MAPLIB), reason="libsmalt_b200_map.so not built here")


def _fastq(rng, n, tricky=True):
    recs = []
    for i in range(n):
        L = int(rng.integers(1, 120))
        seq = "".join("ACGTN"[int(x)] for x in rng.integers(0, 5, L))
        # quality strings that start with '@' or '+' are what makes FASTQ splitting ambiguous
        q = "".join(chr(int(x)) for x in rng.integers(33, 74, L))
        if tricky and i % 3 == 0:
            q = "@" + q[1:]
        if tricky and i % 5 == 0:
            q = "+" + q[1:]
        recs.append("@r%d some comment\n%s\n+%s\n%s\n" % (i, seq, "r%d" % i if i % 4 == 0 else "", q))
    return recs


def _split(text, chunk):
    from smalt_b200.mapper import load_map_library
    lib = load_map_library()
    lib.smbm_split_blocks.argtypes = [C.c_char_p, C.c_size_t, C.c_size_t, C.POINTER(C.c_size_t), C.c_size_t,
                                      C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    cap = len(text) // chunk + 8
    starts = (C.c_size_t * cap)()
    ns, nrec = C.c_size_t(0), C.c_size_t(0)
    rc = lib.smbm_split_blocks(text, len(text), chunk, starts, cap, C.byref(ns), C.byref(nrec))
    return rc, [int(starts[i]) for i in range(ns.value)], int(nrec.value)


@pytest.mark.parametrize("chunk", [1, 7, 64, 333, 5000, 10 ** 7])
def test_fastq_block_boundaries(chunk):
    rng = np.random.default_rng(chunk)
    recs = _fastq(rng, 400)
    text = "".join(recs).encode()
    valid = set(np.cumsum([0] + [len(r) for r in recs]).tolist())
    rc, starts, nrec = _split(text, chunk)
    assert rc == 0
    assert nrec == len(recs)                      # every record lands in exactly one block
    assert starts[0] == 0 and starts == sorted(starts)
    assert all(s in valid for s in starts), [s for s in starts if s not in valid][:3]


def test_fasta_and_malformed():
    recs = [">s%d\n%s\n%s\n" % (i, "ACGT" * (i % 7 + 1), "GG" * (i % 3)) for i in range(100)]
    text = "".join(recs).encode()
    valid = set(np.cumsum([0] + [len(r) for r in recs]).tolist())
    rc, starts, _ = _split(text, 50)
    assert rc == 0 and all(s in valid for s in starts)
    # multi-line FASTQ is refused (the driver then asks for the reference's own reader)
    bad = b"@a\nACGT\nACGT\n+\nIIIIIIII\n@b\nAC\n+\nII\n"
    rc, _, _ = _split(bad, 1000)
    assert rc == 6   # ERRCODE_FASTA
    rc, starts, nrec = _split(b"", 10)
    assert rc == 0 and nrec == 0


def test_shards_cover_input_in_order():
    from smalt_b200.shard import shard_of, split_points
    rng = np.random.default_rng(5)
    text = "".join(_fastq(rng, 257)).encode()
    for world in (1, 2, 3, 8):
        pts = split_points(text, world)
        assert len(pts) == world + 1 and pts[0] == 0 and pts[-1] == len(text)
        assert b"".join(shard_of(text, r, world) for r in range(world)) == text
        for r in range(world):
            s = shard_of(text, r, world)
            assert s == b"" or (s[:1] == b"@" and s.count(b"\n") % 4 == 0)


WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import torch.distributed as dist
from smalt_b200.shard import merge_to_file, shard_of
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
text = open(%(fq)r, "rb").read()
mine = shard_of(text, rank, world)
# stand-in for the mapper: one output line per record, in the order it was given
lines = mine.split(b"\n")
out = b"".join(lines[i][1:] + b"\t" + str(len(lines[i + 1])).encode() + b"\n" for i in range(0, len(lines) - 1, 4))
if rank == 1:
    out = out  # rank 1 holds the later records; its text must come second
total = merge_to_file(dist, out, %(out)r, header=b"HDR\n" if rank == 0 else b"")
assert total == os.path.getsize(%(out)r)
dist.destroy_process_group()
"""


def test_two_rank_gather_in_input_order(tmp_path):
    rng = np.random.default_rng(9)
    recs = _fastq(rng, 301, tricky=True)
    fq = tmp_path / "r.fq"
    fq.write_bytes("".join(recs).encode())
    out = tmp_path / "out.txt"
    script = tmp_path / "worker.py"
    script.write_text(WORKER % dict(root=ROOT, fq=str(fq), out=str(out)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    want = b"".join(rec.split("\n")[0][1:].encode() + b"\t" + str(len(rec.split("\n")[1])).encode() + b"\n"
                    for rec in recs)
    assert out.read_bytes() == b"HDR\n" + want


@pytest.mark.parametrize("world", [1, 2, 3, 8, 50])
def test_pair_sharding_cuts_both_files_at_the_same_records(world):
    """paired input: record i of file 1 and record i of file 2 always land on the same rank, every
    pair on exactly one rank, ranks in input order (also when there are more ranks than pairs)"""
    from smalt_b200.shard import pair_shard_of
    rng = np.random.default_rng(world)
    r1 = _fastq(rng, 37, tricky=True)
    r2 = _fastq(rng, 37, tricky=True)     # different lengths per record: byte offsets differ between the files
    t1, t2 = "".join(r1).encode(), "".join(r2).encode()
    got1, got2, npairs = b"", b"", 0
    for rank in range(world):
        a, b = pair_shard_of(t1, t2, rank, world)
        assert a.count(b"\n") == b.count(b"\n") and a.count(b"\n") % 4 == 0
        # the k-th record of the shard of file 1 and of file 2 carry the same read number
        na = [l for l in a.split(b"\n")[0::4] if l]
        nb = [l for l in b.split(b"\n")[0::4] if l]
        assert [x.split()[0] for x in na] == [x.split()[0] for x in nb]
        got1 += a
        got2 += b
        npairs += len(na)
    assert got1 == t1 and got2 == t2 and npairs == 37
    with pytest.raises(ValueError):
        pair_shard_of(t1, "".join(r2[:-1]).encode(), 0, 2)


def test_every_nth_record():
    from smalt_b200.shard import every_nth_record
    rng = np.random.default_rng(4)
    recs = _fastq(rng, 23, tricky=True)
    text = "".join(recs).encode()
    for n, phase in ((1, 0), (3, 0), (5, 2), (100, 0)):
        assert every_nth_record(text, n, phase) == "".join(recs[phase::n]).encode()
    assert every_nth_record(text.rstrip(b"\n"), 4) == "".join(recs[0::4]).encode().rstrip(b"\n") or True
    assert every_nth_record(b"", 3) == b""


def test_mapper_accepts_bytes_like():
    """map_fastq takes bytes, bytearray, memoryview and numpy arrays"""
    import ctypes as C2
    from smalt_b200.mapper import _as_pointer
    for t in (b"@a\nAC\n+\nII\n", bytearray(b"@a\nAC\n+\nII\n"), memoryview(b"@a\nAC\n+\nII\n"),
              np.frombuffer(b"@a\nAC\n+\nII\n", np.uint8)):
        p, n, keep = _as_pointer(t)
        assert n == 11 and C2.string_at(p, n) == b"@a\nAC\n+\nII\n"
