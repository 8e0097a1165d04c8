"""Parity of the CUDA seed kernel (K1, smb_seed_batch through the C ABI) with the oracle:
seed tables, the unstable-sort order, seed_rank, cover deficit and hit statistics."""
import numpy as np
import pytest

from oracle_lib import Oracle
from seqgen import random_seq
from smalt_b200 import indexer
from test_oracle_k1_vs_ref import make_genome, sample_read

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import smalt_b200
    c = smalt_b200.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("k,nskip,lens", [(13, 6, [150000]), (11, 3, [30011, 20007, 999]),
                                          (7, 1, [3000, 2999]), (20, 13, [40000, 30000])])
def test_seed_batch_vs_oracle(ctx, k, nskip, lens):
    from smalt_b200.capi import pack_sequences
    rng = np.random.default_rng(1000 + k)
    seqs = make_genome(rng, lens)
    ix = indexer.as_loaded(indexer.build_index(seqs, k, nskip))
    orc = Oracle()
    oix = orc.make_index(ix)
    ctx.index_upload(ix)
    reads = [sample_read(rng, seqs, int(rng.integers(max(k, 25), 300))) for _ in range(400)]
    reads.append(random_seq(rng, k - 1))      # too short: ERRCODE_SHORTSEQ
    reads.append(random_seq(rng, k))          # exactly one k-mer
    reads.append(np.zeros(120, np.uint8))     # poly-A: tandem repeat filter
    reads.append(np.full(60, 5, np.uint8))    # all N
    arena, offs = pack_sequences(reads)
    lens_r = np.array([len(r) for r in reads], np.uint32)
    qual = (33 + rng.integers(2, 41, len(arena))).astype(np.uint8)
    ctx.arena_upload(arena)
    for use_qual, basq, maxhit in ((False, 0, 10000), (True, 10, 10000), (False, 0, 5), (False, 0, 0)):
        info, tabs = ctx.seed_batch(offs[:-1], lens_r, qual if use_qual else None, maxhit, 16384, basq)
        assert ctx.last_kernel_launches == 1
        slot = 0
        h = {0: None, 1: None}
        for r, rd in enumerate(reads):
            q = qual[int(offs[r]):int(offs[r + 1])] if use_qual else None
            for s in (0, 1):
                e, want, h[s] = orc.hitinfo(oix, rd, q, s, 1, maxhit, 16384, basq, h=h[s])
                got = info[2 * r + s]
                assert int(got["err"]) == e, (r, s)
                if e == 0:
                    for key in ("n_seeds", "seed_rank", "status", "cover_deficit", "nhit_rank", "nhit_tot",
                                "nhit_all"):
                        assert int(got[key]) == want[key], (key, r, s, maxhit)
                    n = want["n_seeds"]
                    for key in ("posidx", "nhits", "qoffs", "sortkey", "sidx"):
                        assert np.array_equal(tabs[key][slot:slot + n], want[key]), (key, r, s)
                    assert np.array_equal(tabs["qmask"][slot:slot + len(rd)], want["qmask"]), (r, s)
                slot += len(rd)


def _repeat_genome(rng):
    """300 diverged copies of a 400 bp unit: reads from it collect > 32768 hits, which
    exercises the ALLOCBOUNDARY retry of hashCollectHitsForSegment (hashhit.c:1739-1741)."""
    from seqgen import mutate
    unit = random_seq(rng, 400)
    parts = []
    for _ in range(300):
        parts.append(mutate(rng, unit, p_sub=0.01, p_ins=0, p_del=0)[:400])
        parts.append(random_seq(rng, 50))
    return [np.concatenate(parts), random_seq(rng, 20000)], unit


@pytest.mark.parametrize("k,nskip,kind", [(13, 6, "plain"), (11, 3, "multi"), (11, 3, "repeat")])
def test_hits_batch_vs_oracle(ctx, k, nskip, kind):
    from smalt_b200.capi import HIT_REQ_DTYPE, pack_sequences
    rng = np.random.default_rng(2000 + k + len(kind))
    unit = None
    if kind == "plain":
        seqs = make_genome(rng, [150000])
    elif kind == "multi":
        seqs = make_genome(rng, [30011, 20007, 999])
    else:
        seqs, unit = _repeat_genome(rng)
    ix = indexer.as_loaded(indexer.build_index(seqs, k, nskip))
    orc = Oracle()
    oix = orc.make_index(ix)
    ctx.index_upload(ix)
    reads = [sample_read(rng, seqs, 150) for _ in range(150)]
    if unit is not None:
        reads += [np.ascontiguousarray(unit[i:i + 150]) for i in range(0, 200, 20)]
    arena, offs = pack_sequences(reads)
    lens_r = np.array([len(r) for r in reads], np.uint32)
    ctx.arena_upload(arena)
    soffs = np.concatenate([[0], np.cumsum([len(x) for x in seqs])])
    for nhit_max in (10000, 40, 0):
        info, tabs = ctx.seed_batch(offs[:-1], lens_r, None, 10000, 16384, 0)
        req = np.zeros(len(reads) * 2 * len(seqs), HIT_REQ_DTYPE)
        n = 0
        for r in range(len(reads)):
            for s in (0, 1):
                for sx in range(len(seqs)):
                    req[n]["lo"], req[n]["hi"] = soffs[sx], soffs[sx + 1]
                    req[n]["read"], req[n]["nhit_max"], req[n]["strand"], req[n]["use_short"] = r, nhit_max, s, 1
                    n += 1
        sq, first, errs = ctx.hits_batch(req, nhits_alloc=32768)
        assert ctx.last_kernel_launches in (3, 5)   # COUNT, device scan of the list sizes (one launch for short arrays, else three), FILL
        n = 0
        big = 0
        for r, rd in enumerate(reads):
            for s in (0, 1):
                e, _, h = orc.hitinfo(oix, rd, None, s, 1, 10000, 16384, 0)
                hl = orc.lib.so_hitlist_create(32768)
                for sx in range(len(seqs)):
                    if e == 0:
                        e2, want, hl = orc.hitlist_segment(oix, h, soffs[sx], soffs[sx + 1], nhit_max, 1, hl)
                        got = sq[int(first[n]):int(first[n + 1])]
                        assert e2 == 0
                        assert np.array_equal(got, want), (r, s, sx, nhit_max, len(got), len(want))
                        big = max(big, len(want))
                    n += 1
                orc.lib.so_hitlist_delete(hl)
                orc.lib.so_hitinfo_delete(h)
        if kind == "repeat" and nhit_max == 0:
            assert big > 10000


@pytest.mark.parametrize("kind", ["multi", "repeat"])
def test_hits_cutoff_vs_oracle(ctx, kind):
    """whole-set hit lists (mode 2 = hashCollectHitsUsingCutoff, the path for >= 512 reference
    sequences): packed hits, the list's HITQUAL mask, the ceiling / halving retry"""
    from smalt_b200.capi import HIT_REQ_DTYPE, pack_sequences
    rng = np.random.default_rng(3000 + len(kind))
    unit = None
    if kind == "multi":
        seqs = make_genome(rng, [30011, 20007, 999])
    else:
        seqs, unit = _repeat_genome(rng)
    k, nskip = 11, 3
    ix = indexer.as_loaded(indexer.build_index(seqs, k, nskip))
    orc = Oracle()
    oix = orc.make_index(ix)
    ctx.index_upload(ix)
    reads = [sample_read(rng, seqs, int(rng.integers(40, 200))) for _ in range(120)]
    if unit is not None:
        reads += [np.ascontiguousarray(unit[i:i + 150]) for i in range(0, 200, 20)]
    arena, offs = pack_sequences(reads)
    lens_r = np.array([len(r) for r in reads], np.uint32)
    ctx.arena_upload(arena)
    for nhit_max in (10000, 40, 0):
        info, tabs = ctx.seed_batch(offs[:-1], lens_r, None, 10000, 16384, 0)
        req = np.zeros(len(reads) * 2, HIT_REQ_DTYPE)
        req["read"] = np.repeat(np.arange(len(reads), dtype=np.uint32), 2)
        req["strand"] = np.tile(np.array([0, 1], np.uint8), len(reads))
        req["nhit_max"], req["use_short"] = nhit_max, 2
        sq, first, errs = ctx.hits_batch(req)
        qm, qfirst = ctx.hits_qmask(len(req), int(2 * lens_r.sum()))
        n, big = 0, 0
        for r, rd in enumerate(reads):
            for s in (0, 1):
                e, _, h = orc.hitinfo(oix, rd, None, s, 1, 10000, 16384, 0)
                if e == 0:
                    e2, want, wantmask, hl = orc.hitlist_cutoff(oix, h, nhit_max)
                    assert e2 == 0 and int(errs[n]) == 0
                    got = sq[int(first[n]):int(first[n + 1])]
                    assert np.array_equal(got, want), (r, s, nhit_max, len(got), len(want))
                    assert np.array_equal(qm[int(qfirst[n]):int(qfirst[n + 1])], wantmask), (r, s, nhit_max)
                    big = max(big, len(want))
                    orc.lib.so_hitlist_delete(hl)
                orc.lib.so_hitinfo_delete(h)
                n += 1
        if kind == "repeat" and nhit_max == 0:
            assert big > 5000


def test_seed_batch_tables_vs_oracle(ctx):
    """per-read on-the-fly indexes (k=5 s=1 perfect hash, rmap.c:495-517): seed statistics of every
    read against its own table, then restricted hit lists (full seed table, hashhit.c:1691-1769)"""
    from smalt_b200.capi import HIT_REQ_DTYPE, pack_sequences
    rng = np.random.default_rng(77)
    orc = Oracle()
    tables, oixs, gens = [], [], []
    for t in range(9):
        g = [random_seq(rng, int(rng.integers(900, 2600)))]
        ixd = indexer.build_index(g, 5, 1)
        assert ixd["typ"] == 0
        tables.append(ixd)
        oixs.append(orc.make_index(indexer.as_loaded(ixd)))
        gens.append(g)
    reads, rt = [], []
    for r in range(60):
        t = int(rng.integers(0, len(tables)))
        reads.append(sample_read(rng, gens[t], int(rng.integers(30, 200))))
        rt.append(t)
    reads.append(random_seq(rng, 4)); rt.append(0)     # shorter than k
    arena, offs = pack_sequences(reads)
    lens_r = np.array([len(r) for r in reads], np.uint32)
    ctx.arena_upload(arena)
    info = ctx.seed_batch_tables(tables, rt, offs[:-1], lens_r)
    req = np.zeros(2 * len(reads), HIT_REQ_DTYPE)
    for r in range(len(reads)):
        L = len(gens[rt[r]][0])
        for s in (0, 1):
            req[2 * r + s]["lo"], req[2 * r + s]["hi"] = L // 5, L - L // 7
            req[2 * r + s]["read"], req[2 * r + s]["strand"] = r, s
            req[2 * r + s]["nhit_max"], req[2 * r + s]["use_short"] = 10000, 0
    sq, first, errs = ctx.hits_batch(req)
    nlists = 0
    for r, rd in enumerate(reads):
        for s in (0, 1):
            e, want, h = orc.hitinfo(oixs[rt[r]], rd, None, s, 0, 0, 0, 0)
            got = info[2 * r + s]
            assert int(got["err"]) == e, (r, s)
            if e == 0:
                for key in ("n_seeds", "cover_deficit", "nhit_tot", "nhit_all"):
                    assert int(got[key]) == want[key], (key, r, s)
                q = req[2 * r + s]
                e2, wl, hl = orc.hitlist_segment(oixs[rt[r]], h, int(q["lo"]), int(q["hi"]), 10000, 0)
                assert e2 == 0 and int(errs[2 * r + s]) == 0
                assert np.array_equal(sq[int(first[2 * r + s]):int(first[2 * r + s + 1])], wl), (r, s)
                nlists += len(wl) > 0
                orc.lib.so_hitlist_delete(hl)
            orc.lib.so_hitinfo_delete(h)
    assert nlists > 60
