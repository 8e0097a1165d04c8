"""Parity of the device-resident block pipeline (smb_block_run / smb_block_fetch: hit lists ->
candidate selection -> K2 -> score replay -> K3 without leaving the GPU) with the oracle's
restatement of segment.c / rmap.c (itself pinned against the reference's own segment.c,
tests/test_oracle_cand_vs_ref.py) and of the DP kernels."""
import numpy as np
import pytest

from oracle_lib import Oracle
from seqgen import random_seq, revcomp
from smalt_b200 import indexer
from smalt_b200.seqpack import pack3
from test_oracle_cand_vs_ref import repeat_genome, sample

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import smalt_b200
    c = smalt_b200.Context(0)
    yield c
    c.close()


def upload_set(ctx, seqs, k, nskip):
    ix = indexer.as_loaded(indexer.build_index(seqs, k, nskip))
    ctx.index_upload(ix)
    codes = np.concatenate([np.asarray(s, np.uint8) for s in seqs] + [np.array([7], np.uint8)])
    soffs = np.concatenate([[0], np.cumsum([len(s) for s in seqs])]).astype(np.uint64)
    ctx.refseq_upload(pack3(codes), len(codes), soffs)
    return ix, soffs


def check_block(ctx, orc, oix, seqs, soffs, reads, k, nskip, mode, min_swatscor=20):
    from smalt_b200.capi import BLOCK_JOB_DTYPE, pack_sequences
    arena, offs = pack_sequences(reads)
    ctx.arena_upload(arena)
    lens = np.array([len(r) for r in reads], np.uint32)
    nhit_max = mode.get("nhit_max", 10000)
    ctx.seed_batch(offs[:-1], lens, None, nhit_max, mode.get("maxhit_total", 16384), 0, full=False)
    jobs = np.zeros(len(reads), BLOCK_JOB_DTYPE)
    jobs["seed_read"] = np.arange(len(reads))
    jobs["niv"] = -1
    jobs["min_cover"] = mode.get("min_cover", 0)
    jobs["min_swatscor"] = min_swatscor
    best = mode.get("best", False)
    below = mode.get("min_swatscor_below_max", -1)
    sz = ctx.block_run(jobs, None, nhit_max, below, mode.get("target_depth", 200), mode.get("max_depth", 8000), best,
                       mode.get("sensitive", False), False)
    rd, k3c, k3err, first, res, diff = ctx.block_fetch()
    cfirst, cands, cover, qsqe = ctx.block_debug_cands()
    nk3_seen = nmulti = nres_seen = 0
    omode = {kk: v for kk, v in mode.items()}
    for r, read in enumerate(reads):
        e, st, want = orc.candidates(oix, read, soffs, termchar=0, **omode)
        assert int(rd[r]["errcode"]) == e, (r, int(rd[r]["errcode"]), e)
        if e:
            continue
        assert (int(rd[r]["nseg"]), int(rd[r]["nseg_tot"]), int(rd[r]["nhit"]), int(rd[r]["nhit_tot"])) == \
            (st["n_sort"], st["n_mincover"], st["nhit"], st["nhit_tot"]), (r, rd[r], st)
        c0, c1 = int(cfirst[r]), int(cfirst[r + 1])
        assert c1 - c0 == len(want) == int(rd[r]["ncand"]), (r, c1 - c0, len(want))
        nmulti += len(want) > 1
        scores, bl, br = [], [], []
        for c, w in enumerate(want):
            g = cands[c0 + c]
            got = (int(g["rs"]), int(g["reflen"]), int(g["band_l"]), int(g["band_r"]), int(g["sqidx"]),
                   int(g["reverse"]), int(cover[c0 + c]), int(qsqe[c0 + c, 0]), int(qsqe[c0 + c, 1]))
            exp = (w["rs"], w["re"] - w["rs"] + 1, w["band_l"], w["band_r"], w["sqidx"], w["flags"] & 1, w["cover"],
                   w["qs"], w["qe"])
            assert got == exp, (r, c, got, exp)
            # K2 score of the candidate (the SIMD predicate of rmap.c:715-718 selects the kernel)
            prof = revcomp(read) if w["flags"] & 1 else read
            win = np.ascontiguousarray(seqs[w["sqidx"]][w["rs"]:w["re"] + 1])
            simd = len(read) >= 32 and (w["band_r"] - w["band_l"]) * 48 > len(read) and w["qs"] == 0 and w["qe"] >= len(read) - 1
            if simd:
                es, sc = orc.sw_striped(np.ascontiguousarray(prof), win)
            else:
                es, sc, _ = orc.band_fast(np.ascontiguousarray(prof), win, w["band_l"], w["band_r"], w["qs"], w["qe"], 0,
                                          len(win) - 1)
            assert es == 0
            scores.append(sc)
            bl.append(w["band_l"])
            br.append(w["band_r"])
        nsc = int(rd[r]["nscored"])
        e2, rp = orc.score_replay([w["cover"] for w in want], [w["flags"] & 1 for w in want], scores, bl, br,
                                  st["cover_deficit"], len(read), k, nskip, min_swatscor, below, best)
        assert e2 == 0
        assert (nsc, int(rd[r]["max1scor"]), int(rd[r]["max2scor"])) == (rp["nscored"], rp["max1"], rp["max2"]), (r, rd[r], rp)
        # scores of the candidates the reference scores (later ones are computed but never looked at)
        for c in range(min(nsc + 1, len(want))):
            assert int(cands[c0 + c]["swscor"]) == scores[c], (r, c)
        if rp["max1"] < 1:
            assert not rd[r]["do_align"] and int(rd[r]["nk3"]) == 0
            continue
        assert rd[r]["do_align"]
        assert (int(rd[r]["min_swatscor"]), int(rd[r]["scorlen_min"]), int(rd[r]["bandwidth_min"])) == \
            (rp["min_swatscor"], rp["scorlen_min"], rp["bandwidth_min"]), (r, rd[r], rp)
        sel = [c for c in range(len(want)) if rp["align"][c]]
        k0 = int(rd[r]["k3_first"])
        assert int(rd[r]["nk3"]) == len(sel), (r, rd[r], sel)
        for i, c in enumerate(sel):
            w = want[c]
            g = k3c[k0 + i]
            assert (int(g["rs"]), int(g["reflen"]), int(g["sqidx"]), int(g["reverse"]), int(g["swscor"]), int(g["band_l"]),
                    int(g["band_r"])) == (w["rs"], w["re"] - w["rs"] + 1, w["sqidx"], w["flags"] & 1, scores[c],
                                          int(rp["band_l"][c]), int(rp["band_r"][c])), (r, c)
            prof = np.ascontiguousarray(revcomp(read) if w["flags"] & 1 else read)
            win = np.ascontiguousarray(seqs[w["sqidx"]][w["rs"]:w["re"] + 1])
            ea, exp_res, _ = orc.band_align(prof, win, int(rp["band_l"][c]), int(rp["band_r"][c]), w["qs"], w["qe"], 0,
                                            len(win) - 1, rp["min_swatscor"], rp["scorlen_min"])
            assert int(k3err[k0 + i]) == ea, (r, c, int(k3err[k0 + i]), ea)
            got_res = [((int(x["score"]), int(x["qs"]), int(x["qe"]), int(x["rs"]), int(x["re"])),
                        bytes(diff[x["diff_off"]:x["diff_off"] + x["diff_len"]]))
                       for x in res[first[k0 + i]:first[k0 + i + 1]]]
            assert got_res == exp_res, (r, c, got_res, exp_res)
            assert all(int(x["task"]) == k0 + i for x in res[first[k0 + i]:first[k0 + i + 1]])
            nres_seen += len(got_res)
        nk3_seen += len(sel)
    assert int(sz["nk3"]) == nk3_seen and int(sz["nresults"]) == nres_seen
    return nmulti, nk3_seen


MODES = [dict(), dict(best=True, min_swatscor_below_max=0), dict(min_swatscor_below_max=12),
         dict(target_depth=2, max_depth=5), dict(target_depth=2, sensitive=True), dict(nhit_max=40, maxhit_total=400)]


@pytest.mark.parametrize("k,nskip,lens,qlen", [(13, 6, [60000, 45000], 150), (11, 3, [30011, 20007, 999], 100),
                                               (13, 2, [50000], 250), (7, 1, [3000, 2999], 36), (20, 13, [90000, 70000], 150)])
def test_block_vs_oracle(ctx, k, nskip, lens, qlen):
    rng = np.random.default_rng(2000 + k * 10 + nskip)
    seqs = repeat_genome(rng, lens, unit_len=min(400, min(lens) // 4))
    ix, soffs = upload_set(ctx, seqs, k, nskip)
    orc = Oracle()
    oix = orc.make_index(ix)
    tot_multi = tot_k3 = 0
    for m, mode in enumerate(MODES):
        reads = []
        for it in range(160):
            rd = sample(rng, seqs, qlen, err=[0.0, 0.02, 0.06][it % 3])
            if it % 17 == 5:
                rd = random_seq(rng, qlen)
            if it % 23 == 7:
                rd = sample(rng, seqs, max(k, qlen // 2 + it % 9), 0.02)    # ragged lengths in one block
            if it % 41 == 11:
                rd[int(rng.integers(0, len(rd)))] = 5
            reads.append(np.ascontiguousarray(rd))
        reads.append(random_seq(rng, k - 1))    # ERRCODE_SHORTSEQ
        nm, n3 = check_block(ctx, orc, oix, seqs, soffs, reads, k, nskip, mode)
        tot_multi += nm
        tot_k3 += n3
    assert tot_multi > 20 and tot_k3 > 200


def test_block_empty_and_unmappable(ctx):
    from smalt_b200.capi import BLOCK_JOB_DTYPE, pack_sequences
    rng = np.random.default_rng(77)
    seqs = [random_seq(rng, 20000)]
    ix, soffs = upload_set(ctx, seqs, 13, 6)
    sz = ctx.block_run(np.zeros(0, BLOCK_JOB_DTYPE))
    assert int(sz["nk3"]) == 0
    reads = [random_seq(rng, 100) for _ in range(5)]   # nothing maps: no hits at all
    arena, offs = pack_sequences(reads)
    ctx.arena_upload(arena)
    ctx.seed_batch(offs[:-1], np.full(5, 100, np.uint32), None, 10000, 16384, 0, full=False)
    jobs = np.zeros(5, BLOCK_JOB_DTYPE)
    jobs["seed_read"] = np.arange(5)
    jobs["niv"] = -1
    sz = ctx.block_run(jobs)
    rd, k3c, k3err, first, res, diff = ctx.block_fetch()
    assert int(sz["nk3"]) == 0 and len(res) == 0
    assert all(int(x["errcode"]) == 0 and int(x["ncand"]) == 0 and x["reached_stats"] for x in rd)
