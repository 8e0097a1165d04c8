"""Pins the K1 part of the C restatement (index lookup, seed table, ranking, hit lists)
against the UNMODIFIED reference (oracle/_ref/libsmalt_ref.so)."""
import numpy as np
import pytest

from oracle_lib import Oracle, RefLib, have_ref
from seqgen import mutate, random_seq, revcomp
from smalt_b200 import indexer

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")


def make_genome(rng, lens, repeats=True):
    seqs = [random_seq(rng, n, p_n=0.0005) for n in lens]
    if repeats:  # diverged repeat copies -> multi-hit seeds
        unit = random_seq(rng, 300)
        for s in seqs:
            for _ in range(6):
                p = int(rng.integers(0, len(s) - 300))
                s[p:p + 300] = mutate(rng, unit, p_sub=0.02, p_ins=0, p_del=0)[:300]
        seqs[0][100:160] = 0  # poly-A: tandem-repeat filter and a very frequent word
        seqs[0][500:520] = np.tile([0, 1], 10)
    return seqs


def sample_read(rng, seqs, qlen, err=0.03):
    s = seqs[int(rng.integers(0, len(seqs)))]
    st = int(rng.integers(0, len(s) - qlen))
    rd = mutate(rng, s[st:st + qlen].copy(), p_sub=err, p_ins=err / 4, p_del=err / 4)
    if rng.random() < 0.5:
        rd = revcomp(rd)
    if rng.random() < 0.2:
        rd[rng.integers(0, len(rd))] = 5
    return np.ascontiguousarray(rd)


@pytest.fixture(scope="module", params=[(11, 3, [30011, 20007, 999]), (13, 6, [120000]),
                                        (7, 1, [3000, 2999]), (20, 13, [40000, 30000])])
def setup(request, tmp_path_factory):
    k, nskip, lens = request.param
    rng = np.random.default_rng(k * 100 + nskip)
    seqs = make_genome(rng, lens)
    ix = indexer.build_index(seqs, k, nskip)
    pref = str(tmp_path_factory.mktemp("ix") / "g")
    indexer.write_smi(pref, ix)
    indexer.write_sma(pref, ["s%d" % i for i in range(len(seqs))], seqs)
    ref = RefLib()
    ref.index_load(pref)
    orc = Oracle()
    oix = orc.make_index(indexer.as_loaded(ix))
    return dict(k=k, nskip=nskip, seqs=seqs, ix=ix, ref=ref, orc=orc, oix=oix, rng=rng)


def test_lookup(setup):
    s = setup
    rng = s["rng"]
    k = s["k"]
    words = []
    g = s["seqs"][0]
    for _ in range(300):
        p = int(rng.integers(0, len(g) - k))
        w = 0
        for c in g[p:p + k]:
            w = (w << 2) | (int(c) & 3)
        if rng.random() < 0.3:
            w ^= 1 << int(rng.integers(0, 2 * k))
        words.append(w | (int(rng.integers(0, 4)) << (2 * k)) if k < 30 else w)  # junk above 2k bits is masked
    nh_r, px_r = s["ref"].lookup(np.array(words, np.uint64))
    nh_o, px_o = s["orc"].lookup(s["oix"], words)
    assert np.array_equal(nh_r, nh_o)
    hit = nh_r > 0
    assert np.array_equal(px_r[hit], px_o[hit])
    assert hit.sum() > 20


KEYS = ("n_seeds", "seed_rank", "status", "cover_deficit", "nhit_rank", "nhit_tot", "nhit_all")
ARRS = ("posidx", "nhits", "qoffs", "sortkey", "sidx", "qmask")


def _cmp_info(a, b, ctx):
    for key in KEYS:
        assert a[key] == b[key], (key, ctx)
    for key in ARRS:
        assert np.array_equal(a[key], b[key]), (key, ctx)


def test_hitinfo_and_hitlists(setup):
    s = setup
    rng, ref, orc, oix = s["rng"], s["ref"], s["orc"], s["oix"]
    nseq = len(s["seqs"])
    soffs = np.concatenate([[0], np.cumsum([len(x) for x in s["seqs"]])])
    nhits_seen = 0
    h = {0: None, 1: None}
    hl = None
    for it in range(120):
        qlen = int(rng.integers(max(s["k"], 25), 260))
        rd = sample_read(rng, s["seqs"], qlen)
        qual = None
        if it % 3 == 0:
            qual = (33 + rng.integers(2, 41, len(rd))).astype(np.uint8)
        basq = 10 if it % 6 == 0 else 0
        maxhit = [10000, 10000, 4, 0][it % 4]
        for is_short in (1, 0):
            for strand in (0, 1):
                er, ir = ref.hitinfo(rd, qual, strand, is_short, maxhit, 16384, basq)
                eo, io, h[strand] = orc.hitinfo(oix, rd, qual, strand, is_short, maxhit, 16384, basq, h=h[strand])
                assert er == eo, (it, strand)
                if er:
                    continue
                _cmp_info(ir, io, (it, strand, is_short))
                if ir["n_seeds"] == 0:
                    continue
                # hit lists: per reference sequence in order (collectHits SEQBYSEQ, rmap.c:283-318) ...
                nmax = [10000, 50, 0][it % 3]
                if is_short:
                    for sx in range(nseq):
                        e1, d1, _ = ref.hitlist(strand, sx, nmax, 1, len(rd))
                        e2, d2, hl = orc.hitlist_segment(oix, h[strand], soffs[sx], soffs[sx + 1], nmax, 1, hl)
                        assert e1 == e2, (it, strand, sx)
                        assert e1 != 0 or np.array_equal(d1, d2), (it, strand, sx)
                        nhits_seen += len(d1)
                # ... and over the whole set (hashCollectHitsUsingCutoff)
                e1, d1, q1 = ref.hitlist(strand, -1, nmax, 1, len(rd))
                e2, d2, q2, hl = orc.hitlist_cutoff(oix, h[strand], nmax, hl)
                assert e1 == e2, (it, strand, "cutoff")
                if e1 == 0:
                    assert np.array_equal(d1, d2), (it, strand, "cutoff")
                    assert np.array_equal(q1, q2)
    assert nhits_seen > 300


def test_short_read_error(setup):
    s = setup
    rd = random_seq(s["rng"], s["k"] - 1)
    er, _ = s["ref"].hitinfo(rd, None, 0, 1)
    eo, _, _ = s["orc"].hitinfo(s["oix"], rd, None, 0, 1)
    assert er == eo == 30  # ERRCODE_SHORTSEQ
