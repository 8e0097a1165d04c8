"""Output stage on the device (csrc/cigar.cu): CIGAR text and NM edit distance of alignments against the
oracle's restatement (oracle/smalt_oracle_cigar.c, pinned against the reference's writeDiffStrCIGAR and
diffStrGetLevenshteinDistance in tests/test_oracle_cigar_vs_ref.py) - explicit alignment strings
(smb_cigar_batch) and the alignments of a resident block (smb_block_fetch_cigar)."""
import numpy as np
import pytest

from diffgen import encode_columns, random_bytes_string, random_columns
from oracle_lib import Oracle
from seqgen import random_seq
from smalt_b200.capi import CIGAR_ON, CIGAR_SOFTCLIP, CIGAR_XMISMATCH
from test_gpu_block import upload_set
from test_oracle_cand_vs_ref import repeat_genome, sample
from test_oracle_cigar_vs_ref import _cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import smalt_b200
    c = smalt_b200.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("flags", [0, CIGAR_SOFTCLIP, CIGAR_XMISMATCH, CIGAR_SOFTCLIP | CIGAR_XMISMATCH])
def test_cigar_batch_vs_oracle(ctx, flags):
    orc = Oracle()
    rng = np.random.default_rng(5)
    strs = list(_cases(rng))
    strs += [encode_columns(random_columns(rng, int(rng.integers(1, 400)))) for _ in range(3000)]
    strs += [b"\0", bytes([(1 << 6) | 3, 0]), bytes([(0 << 6) | 9, 0])]    # the reference's error cases
    offs = np.cumsum([0] + [len(s) for s in strs])[:-1].astype(np.uint32)
    cs = rng.integers(0, 50, len(strs)).astype(np.uint32)
    ce = rng.integers(0, 20000, len(strs)).astype(np.uint32)
    cs[::3] = 0
    ce[::5] = 0
    first, nm, text = ctx.cigar_batch(np.frombuffer(b"".join(strs), np.uint8), offs, cs, ce, flags)
    assert int(first[0]) == 0 and int(first[-1]) == len(text)
    nerr = 0
    for i, d in enumerate(strs):
        want, wnm = orc.cigar(d, int(cs[i]), int(ce[i]), bool(flags & CIGAR_SOFTCLIP), bool(flags & CIGAR_XMISMATCH))
        got = text[int(first[i]):int(first[i + 1])]
        if want is None:
            assert int(nm[i]) == wnm and got == b"", (i, d)
            nerr += 1
        else:
            assert got == want, (i, d, got, want)
            assert int(nm[i]) == wnm, (i, d)
    assert nerr == 3
    assert ctx.cigar_batch(np.zeros(1, np.uint8), np.zeros(0, np.uint32), [], [], flags)[2] == b""


@pytest.mark.parametrize("flags,qlen", [(CIGAR_ON | CIGAR_SOFTCLIP, 150), (CIGAR_ON | CIGAR_SOFTCLIP | CIGAR_XMISMATCH, 100),
                                        (CIGAR_ON, 250)])
def test_block_cigar_vs_oracle(ctx, flags, qlen):
    """every alignment a resident block returns carries the CIGAR / NM the reference would print for it:
    clips from qs, qe and the read length on the aligned strand, reads with indels and clipped ends"""
    from smalt_b200.capi import BLOCK_JOB_DTYPE, pack_sequences
    orc = Oracle()
    rng = np.random.default_rng(31 + qlen)
    seqs = repeat_genome(rng, [60000, 45000], unit_len=400)
    upload_set(ctx, seqs, 13, 3)
    reads = []
    for it in range(400):
        rd = sample(rng, seqs, qlen if it % 7 else qlen // 2 + it % 11, err=[0.0, 0.03, 0.08][it % 3])
        if it % 5 == 0:   # an unrelated end: the local alignment stops short of it -> clip
            rd[-20:] = random_seq(rng, 20)
        if it % 9 == 0:
            rd[:12] = random_seq(rng, 12)
        reads.append(np.ascontiguousarray(rd))
    arena, offs = pack_sequences(reads)
    ctx.arena_upload(arena)
    lens = np.array([len(r) for r in reads], np.uint32)
    ctx.seed_batch(offs[:-1], lens, None, 10000, 16384, 0, full=False)
    jobs = np.zeros(len(reads), BLOCK_JOB_DTYPE)
    jobs["seed_read"], jobs["niv"], jobs["min_swatscor"] = np.arange(len(reads)), -1, 20
    sz = ctx.block_run(jobs, None, 10000, -1, 200, 8000, False, False, False, cigar=flags)
    rd, k3c, k3err, first, res, diff, cfirst, nm, text = ctx.block_fetch_cigar()
    assert len(res) == int(sz["nresults"]) > 300 and len(text) == int(sz["ncigarbytes"]) == int(cfirst[-1])
    task_read = np.zeros(int(sz["nk3"]), np.int64)
    for r in range(len(reads)):
        task_read[int(rd[r]["k3_first"]):int(rd[r]["k3_first"]) + int(rd[r]["nk3"])] = r
    nclip = nindel = 0
    for i, x in enumerate(res):
        ql = len(reads[task_read[int(x["task"])]])
        d = bytes(diff[x["diff_off"]:x["diff_off"] + x["diff_len"]])
        want, wnm = orc.cigar(d, int(x["qs"]), ql - 1 - int(x["qe"]), bool(flags & CIGAR_SOFTCLIP),
                              bool(flags & CIGAR_XMISMATCH))
        assert text[int(cfirst[i]):int(cfirst[i + 1])] == want, (i, d)
        assert int(nm[i]) == wnm
        nclip += int(x["qs"]) > 0 or int(x["qe"]) < ql - 1
        nindel += b"I" in want or b"D" in want
    assert nclip > 30 and nindel > 30
    # the same block without the stage: no text, and the plain fetch is unchanged
    sz0 = ctx.block_run(jobs, None, 10000, -1, 200, 8000, False, False, False)
    assert int(sz0["ncigarbytes"]) == 0
    res0 = ctx.block_fetch()[4]
    assert np.array_equal(res0, res)
