"""Parity of the CUDA DP kernels (through the C ABI, libsmalt_b200.so) with the oracle:
K2 smb_sw_score_batch, K2' smb_band_score_batch, K3 smb_band_align_batch.
Bit-exact: scores, coordinates, DiffStr bytes, result order, error codes."""
import numpy as np
import pytest

from golden_io import load_bam_cigar, load_trace, parse_record
from oracle_lib import Oracle
from seqgen import random_seq, read_window_pair, revcomp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import smalt_b200
    c = smalt_b200.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def orc():
    return Oracle()


def _arena(pairs):
    from smalt_b200.capi import pack_sequences
    seqs = []
    for rd, win in pairs:
        seqs += [rd, win]
    arena, offs = pack_sequences(seqs)
    return arena, offs


def _sw_tasks(pairs, offs, flags=None):
    from smalt_b200.capi import SW_TASK_DTYPE
    t = np.zeros(len(pairs), SW_TASK_DTYPE)
    for i, (rd, win) in enumerate(pairs):
        t[i] = (offs[2 * i], offs[2 * i + 1], len(rd), len(win), 0 if flags is None else flags[i], 0)
    return t


def _band_tasks(pairs, offs, args, minscore=None, minscorlen=None):
    from smalt_b200.capi import BAND_TASK_DTYPE
    t = np.zeros(len(pairs), BAND_TASK_DTYPE)
    for i, (rd, win) in enumerate(pairs):
        t[i]["read_off"], t[i]["ref_off"] = offs[2 * i], offs[2 * i + 1]
        t[i]["read_len"], t[i]["ref_len"] = len(rd), len(win)
        (t[i]["l_edge"], t[i]["r_edge"], t[i]["p_left"], t[i]["p_right"], t[i]["u_left"],
         t[i]["u_right"]) = args[i]
        if minscore is not None:
            t[i]["minscore"], t[i]["minscorlen"] = minscore[i], minscorlen[i]
    return t


def _unpack(res, first, diff, i):
    out = []
    for r in res[first[i]:first[i + 1]]:
        d = bytes(diff[r["diff_off"]:r["diff_off"] + r["diff_len"]])
        out.append(((int(r["score"]), int(r["qs"]), int(r["qe"]), int(r["rs"]), int(r["re"])), d))
    return out


def _rand_pairs(rng, n, qmin, qmax, nonstd=True):
    pairs = []
    for it in range(n):
        qlen = int(rng.integers(qmin, qmax))
        mut = [dict(p_sub=0.02, p_ins=0.005, p_del=0.005), dict(p_sub=0.1, p_ins=0.04, p_del=0.04),
               dict(p_sub=0.0, p_ins=0.0, p_del=0.0)][it % 3]
        rd, win, lf = read_window_pair(rng, qlen, with_flank=True, **mut)
        if nonstd and it % 6 == 5:
            rd[rng.integers(0, len(rd))] = 5
            win[rng.integers(0, len(win))] = 5
            win[rng.integers(0, len(win))] = 4
            rd[rng.integers(0, len(rd))] = 4
        if it % 9 == 8:
            win = random_seq(rng, len(win))  # unrelated
        pairs.append((rd, win, lf))
    return pairs


def _band_args(rng, qlen, rlen, lf):
    style = rng.integers(0, 6)
    if style == 0 or style > 3:
        off = -lf + int(rng.integers(-6, 7))
        w = int(rng.integers(2, 40))
        return off - w, off + w, 0, qlen - 1, 0, rlen - 1
    if style == 1:
        pl, pr = int(rng.integers(0, qlen // 2)), int(rng.integers(qlen // 2, qlen + 3))
        ul, ur = int(rng.integers(0, rlen // 2)), int(rng.integers(rlen // 2, rlen + 3))
        l = int(rng.integers(-rlen, qlen))
        return l, l + int(rng.integers(0, 60)), pl, pr, ul, ur
    if style == 2:
        l = int(rng.integers(-50, 50))
        return l, l - int(rng.integers(0, 5)), 0, qlen - 1, 0, rlen - 1
    l = int(rng.integers(-2 * rlen, 2 * qlen))
    return (l, l + int(rng.integers(0, 2 * qlen)), int(rng.integers(-2, qlen)), int(rng.integers(-2, qlen + 2)),
            int(rng.integers(-2, rlen)), int(rng.integers(-2, rlen + 2)))


def test_sw_score_vs_oracle(ctx, orc):
    rng = np.random.default_rng(101)
    # every columns-per-lane class (1..8), the multi-block path (> 256), ragged lengths
    trip = _rand_pairs(rng, 400, 8, 300) + _rand_pairs(rng, 40, 257, 1200) + _rand_pairs(rng, 6, 2000, 4000)
    pairs = [(a, b) for a, b, _ in trip]
    arena, offs = _arena(pairs)
    ctx.arena_upload(arena)
    scores, errs = ctx.sw_score(_sw_tasks(pairs, offs))
    assert ctx.last_kernel_launches >= 8
    for i, (rd, win) in enumerate(pairs):
        e, s = orc.sw_striped(rd, win)
        assert (int(errs[i]), int(scores[i])) == (e, s), (i, len(rd), len(win))
    # reverse-complement flag == profiling the reverse complement
    flags = [1] * len(pairs)
    scores_rc, _ = ctx.sw_score(_sw_tasks(pairs, offs, flags))
    for i in range(0, len(pairs), 7):
        rd, win = pairs[i]
        assert int(scores_rc[i]) == orc.sw_striped(revcomp(rd), win)[1]


def test_sw_score_n_bases_and_ragged_pairs(ctx, orc):
    """two-tasks-per-warp kernel: N in reads and windows (no X: stays on the PRMT path, N rows are
    masked to score 0 against every column, N against N included), tasks of very different window
    and read lengths sharing a warp (padding rows / columns), windows up to the staging limit"""
    rng = np.random.default_rng(111)
    pairs = []
    for it in range(600):
        qlen = int(rng.integers(20, 257))
        rd, win, _ = read_window_pair(rng, qlen, with_flank=True, p_sub=0.03, p_ins=0.01, p_del=0.01)
        if it % 5 == 0:
            win = np.concatenate([win, random_seq(rng, int(rng.integers(1, 500 - len(win))))]) if len(win) < 480 else win
        pn = (0.0, 0.02, 0.15, 0.6)[it % 4]
        rd[rng.random(len(rd)) < pn] = 5
        win[rng.random(len(win)) < pn] = 5
        if it % 50 == 0:
            rd[:] = 5
        pairs.append((np.ascontiguousarray(rd), np.ascontiguousarray(win)))
    arena, offs = _arena(pairs)
    ctx.arena_upload(arena)
    scores, errs = ctx.sw_score(_sw_tasks(pairs, offs))
    for i, (rd, win) in enumerate(pairs):
        e, s = orc.sw_striped(rd, win)
        assert (int(errs[i]), int(scores[i])) == (e, s), (i, len(rd), len(win))


def test_sw_score_packed_reference(ctx, orc):
    """windows read from the 3-bit packed reference store (the .sma layout)"""
    from smalt_b200.capi import SW_TASK_DTYPE, pack_sequences
    from smalt_b200.seqpack import concat_set, pack3
    rng = np.random.default_rng(102)
    genome = [random_seq(rng, 5000, p_n=0.002), random_seq(rng, 3333)]
    codes, soffs = concat_set(genome)
    ctx.refseq_upload(pack3(codes), len(codes), soffs)
    reads, tasks = [], []
    for i in range(200):
        s = int(rng.integers(0, 2))
        qlen = int(rng.integers(30, 200))
        st = int(rng.integers(0, len(genome[s]) - qlen - 40))
        rd = genome[s][st + 10:st + 10 + qlen].copy()
        rd[rng.integers(0, qlen, 3)] = rng.integers(0, 4, 3)
        reads.append(rd)
        tasks.append((st, qlen + 30, s))
    arena, offs = pack_sequences(reads)
    ctx.arena_upload(arena)
    t = np.zeros(len(reads), SW_TASK_DTYPE)
    for i, (st, wl, s) in enumerate(tasks):
        t[i] = (offs[i], int(soffs[s]) + st, len(reads[i]), wl, 2, 0)
    scores, errs = ctx.sw_score(t)
    for i, (st, wl, s) in enumerate(tasks):
        assert (int(errs[i]), int(scores[i])) == orc.sw_striped(reads[i], np.ascontiguousarray(genome[s][st:st + wl]))


def test_sw_score_edge_cases(ctx, orc):
    from smalt_b200.capi import SW_TASK_DTYPE
    scores, errs = ctx.sw_score(np.zeros(0, SW_TASK_DTYPE))
    assert len(scores) == 0
    pairs = [(np.array([0], np.uint8), np.array([0], np.uint8)),
             (np.array([0, 1, 2], np.uint8), np.array([3], np.uint8)),
             (np.full(40, 5, np.uint8), np.full(50, 5, np.uint8)),
             (np.full(33, 2, np.uint8), np.full(70, 2, np.uint8))]
    arena, offs = _arena(pairs)
    ctx.arena_upload(arena)
    scores, errs = ctx.sw_score(_sw_tasks(pairs, offs))
    for i, (rd, win) in enumerate(pairs):
        assert (int(errs[i]), int(scores[i])) == orc.sw_striped(rd, win)


def test_band_score_vs_oracle(ctx, orc):
    rng = np.random.default_rng(103)
    trip = _rand_pairs(rng, 500, 12, 220)
    pairs = [(a, b) for a, b, _ in trip]
    args = [_band_args(rng, len(a), len(b), lf) for a, b, lf in trip]
    arena, offs = _arena(pairs)
    ctx.arena_upload(arena)
    scores, errs = ctx.band_score(_band_tasks(pairs, offs, args))
    nok = 0
    for i, (rd, win) in enumerate(pairs):
        e, s, _ = orc.band_fast(rd, win, *args[i])
        assert int(errs[i]) == e, (i, args[i])
        if e == 0:
            assert int(scores[i]) == s, (i, args[i])
            nok += 1
    assert nok > 200


def test_band_align_vs_oracle(ctx, orc):
    rng = np.random.default_rng(104)
    trip = _rand_pairs(rng, 600, 20, 260)
    pairs, args = [], []
    for k, (a, b, lf) in enumerate(trip):
        if k % 7 == 0:  # two copies of the target -> recursion produces several results
            b = np.concatenate([b, random_seq(rng, 9), b])
        pairs.append((a, b))
        args.append(_band_args(rng, len(a), len(b), lf))
    minscore = [int(x) for x in rng.integers(1, 40, len(pairs))]
    minscorlen = [int(x) for x in rng.integers(5, 30, len(pairs))]
    arena, offs = _arena(pairs)
    ctx.arena_upload(arena)
    res, first, diff, errs, cells = ctx.band_align(_band_tasks(pairs, offs, args, minscore, minscorlen))
    tot, ocells = 0, 0
    for i, (rd, win) in enumerate(pairs):
        e, want, c = orc.band_align(rd, win, *args[i], minscore[i], minscorlen[i])
        ocells += c
        assert int(errs[i]) == e, (i, args[i])
        assert _unpack(res, first, diff, i) == want, (i, args[i], minscore[i], minscorlen[i])
        tot += len(want)
    assert tot > 200
    assert cells == ocells


@pytest.mark.parametrize("name", ["dp_trace_c1.txt", "dp_trace_hard.txt"])
def test_golden_traces(ctx, name):
    """DP boundary calls recorded from the reference's own `smalt map` runs"""
    recs = load_trace(name)
    _run_records(ctx, recs)


def test_reference_known_answers(ctx):
    g = load_bam_cigar()
    _run_records(ctx, [parse_record(l) for l in g["trace"]["cigar"]])


def _run_records(ctx, recs):
    for kind in ("SW", "BF", "BA"):
        sel = [r for r in recs if r["kind"] == kind]
        if not sel:
            continue
        pairs = [(r["read"], r["ref"]) for r in sel]
        arena, offs = _arena(pairs)
        ctx.arena_upload(arena)
        if kind == "SW":
            scores, errs = ctx.sw_score(_sw_tasks(pairs, offs))
            for i, r in enumerate(sel):
                assert (int(errs[i]), int(scores[i])) == (r["err"], r["score"])
        elif kind == "BF":
            scores, errs = ctx.band_score(_band_tasks(pairs, offs, [r["args"] for r in sel]))
            for i, r in enumerate(sel):
                assert int(errs[i]) == r["err"]
                if r["err"] == 0:
                    assert int(scores[i]) == r["score"]
        else:
            t = _band_tasks(pairs, offs, [r["args"] for r in sel], [r["minscore"] for r in sel],
                            [r["minscorlen"] for r in sel])
            res, first, diff, errs, _ = ctx.band_align(t)
            for i, r in enumerate(sel):
                assert int(errs[i]) == r["err"]
                assert _unpack(res, first, diff, i) == r["results"], i


@pytest.mark.parametrize("pack8", [False, True])
def test_band_align_packed_kernel_geometry(ctx, orc, pack8, monkeypatch):
    """Geometries at the limits of the four-tasks-per-warp kernel (band_pack.cu): window lengths
    around multiples of 32 up to 256 rows, bands of 1..32 diagonals, very different task sizes
    paired in one half-warp, odd task counts and batches of one.  pack8: bands of at most 24
    diagonals go to the <8 lanes, 3 diagonals> instance of the kernel (opt-in, SMB_PACK8)."""
    if pack8:
        monkeypatch.setenv("SMB_PACK8", "1")
    else:
        monkeypatch.delenv("SMB_PACK8", raising=False)
    rng = np.random.default_rng(105)
    pairs, args = [], []
    for rows in (31, 32, 33, 63, 64, 65, 159, 160, 161, 191, 192, 193, 223, 224, 255, 256):
        for bw in (1, 2, 7, 16, 19, 31, 32):
            qlen = int(rng.integers(20, min(rows, 250) + 1))
            rd = random_seq(rng, qlen)
            off = int(rng.integers(0, max(1, rows - qlen)))
            win = random_seq(rng, rows)
            from seqgen import mutate
            m = mutate(rng, rd.copy(), p_sub=0.03, p_ins=0.01, p_del=0.01)[:rows - off]
            win[off:off + len(m)] = m
            if (rows + bw) % 5 == 0:
                win[int(rng.integers(0, rows))] = 5
                rd[int(rng.integers(0, qlen))] = 5
            if (rows + bw) % 11 == 0:
                win[int(rng.integers(0, rows))] = 4
            pairs.append((rd, win))
            l = -off - bw // 2
            args.append((l, l + bw - 1, 0, qlen - 1, 0, rows - 1))
    order = rng.permutation(len(pairs))          # unlike sizes next to each other
    pairs = [pairs[i] for i in order]
    args = [args[i] for i in order]
    minscore = [int(x) for x in rng.integers(1, 25, len(pairs))]
    minscorlen = [int(x) for x in rng.integers(5, 20, len(pairs))]
    arena, offs = _arena(pairs)
    ctx.arena_upload(arena)
    for sel in (slice(None), slice(0, 1), slice(1, 4), slice(0, len(pairs) - 1)):
        idx = list(range(len(pairs)))[sel]
        sub_pairs = [pairs[i] for i in idx]
        sub_offs = np.concatenate([[offs[2 * i], offs[2 * i + 1]] for i in idx] + [[0]])
        t = _band_tasks(sub_pairs, sub_offs, [args[i] for i in idx], [minscore[i] for i in idx],
                        [minscorlen[i] for i in idx])
        res, first, diff, errs, cells = ctx.band_align(t)
        ocells = 0
        for k, i in enumerate(idx):
            rd, win = pairs[i]
            e, want, c = orc.band_align(rd, win, *args[i], minscore[i], minscorlen[i])
            ocells += c
            assert int(errs[k]) == e, (i, args[i])
            assert _unpack(res, first, diff, k) == want, (i, len(rd), len(win), args[i], minscore[i], minscorlen[i])
        assert cells == ocells


def test_band_align_wide_kernel_geometry(ctx, orc):
    """Geometries of the four-diagonals-per-lane kernel (band_wide.cu): bands of 33..128 diagonals
    (also beyond: thread-per-task kernel), windows of up to 512 rows, reads of up to 512 bases,
    several local alignments per window (recursion), batches of one."""
    from seqgen import mutate
    rng = np.random.default_rng(106)
    pairs, args = [], []
    for rows in (120, 255, 256, 257, 300, 383, 384, 385, 500, 511, 512, 513):
        for bw in (33, 64, 65, 66, 97, 127, 128, 129):
            qlen = int(rng.integers(60, min(rows, 512) + 1))
            if rows % 7 == 0:
                qlen = min(rows, 512)
            rd = random_seq(rng, qlen)
            off = int(rng.integers(0, max(1, rows - qlen)))
            win = random_seq(rng, rows)
            m = mutate(rng, rd.copy(), p_sub=0.04, p_ins=0.02, p_del=0.02)[:rows - off]
            win[off:off + len(m)] = m
            if (rows + bw) % 3 == 0:          # two separate pieces: several results, recursion left/right
                cut = len(m) // 2
                win[off + cut:off + cut + 25] = random_seq(rng, min(25, rows - off - cut))
            if (rows + bw) % 5 == 0:
                win[int(rng.integers(0, rows))] = 5
                rd[int(rng.integers(0, qlen))] = 5
            pairs.append((rd, win))
            l = -off - bw // 2 + int(rng.integers(-6, 7))
            args.append((l, l + bw - 1, 0, qlen - 1, 0, rows - 1))
    order = rng.permutation(len(pairs))
    pairs = [pairs[i] for i in order]
    args = [args[i] for i in order]
    minscore = [int(x) for x in rng.integers(1, 30, len(pairs))]
    minscorlen = [int(x) for x in rng.integers(5, 25, len(pairs))]
    arena, offs = _arena(pairs)
    ctx.arena_upload(arena)
    nmulti = 0
    for sel in (slice(None), slice(0, 1), slice(2, 7)):
        idx = list(range(len(pairs)))[sel]
        sub_pairs = [pairs[i] for i in idx]
        sub_offs = np.concatenate([[offs[2 * i], offs[2 * i + 1]] for i in idx] + [[0]])
        t = _band_tasks(sub_pairs, sub_offs, [args[i] for i in idx], [minscore[i] for i in idx],
                        [minscorlen[i] for i in idx])
        res, first, diff, errs, cells = ctx.band_align(t)
        ocells = 0
        for k, i in enumerate(idx):
            rd, win = pairs[i]
            e, want, c = orc.band_align(rd, win, *args[i], minscore[i], minscorlen[i])
            ocells += c
            assert int(errs[k]) == e, (i, args[i])
            assert _unpack(res, first, diff, k) == want, (i, len(rd), len(win), args[i], minscore[i], minscorlen[i])
            nmulti += len(want) > 1
        assert cells == ocells
    assert nmulti > 5


@pytest.mark.parametrize("packed", [False, True])
def test_band_align_long_kernel_geometry(ctx, orc, packed):
    """(packed: the windows lie in the 3-bit packed reference store, which the kernel stages by bulk async
    copies in chunks of 1024 rows; else in the byte arena.)  Geometries of the CTA-per-task long-read kernel (band_long.cu): bands of 129..4096 diagonals (16 or 32
    per thread; beyond: thread-per-task kernel), windows of 513..4000 rows, reads of up to 4000 bases with
    indel-rich errors, several local alignments per window (recursion), sub-ranges of read and window."""
    from seqgen import mutate
    rng = np.random.default_rng(107)
    pairs, args = [], []
    for rows, bw in ((600, 129), (513, 300), (700, 64), (1500, 700), (2000, 2047), (2000, 2048), (2100, 2049),
                     (1800, 3000), (2600, 4096), (1000, 4097), (3000, 1500), (4000, 900), (900, 130), (520, 2500)):
        qlen = int(rng.integers(rows // 2, rows - 20))
        rd = random_seq(rng, qlen)
        off = int(rng.integers(0, max(1, rows - qlen)))
        win = random_seq(rng, rows)
        m = mutate(rng, rd.copy(), p_sub=0.03, p_ins=0.04, p_del=0.04)[:rows - off]
        win[off:off + len(m)] = m
        if (rows + bw) % 3 == 0:          # two separate pieces: several results, recursion left/right
            cut = len(m) // 2
            win[off + cut:off + cut + 40] = random_seq(rng, min(40, rows - off - cut))
        if (rows + bw) % 5 == 0:
            win[int(rng.integers(0, rows))] = 5
            rd[int(rng.integers(0, qlen))] = 5
        pairs.append((rd, win))
        l = -off - bw // 2 + int(rng.integers(-6, 7))
        if rows == 3000:                  # sub-ranges of the read segment and of the window
            args.append((l, l + bw - 1, 50, qlen - 80, 30, rows - 100))
        else:
            args.append((l, l + bw - 1, 0, qlen - 1, 0, rows - 1))
    minscore = [int(x) for x in rng.integers(20, 60, len(pairs))]
    minscorlen = [int(x) for x in rng.integers(10, 40, len(pairs))]
    arena, offs = _arena(pairs)
    ctx.arena_upload(arena)
    if packed:
        from smalt_b200.capi import SMB_TASK_REF_PACKED
        from smalt_b200.seqpack import pack3
        woff, parts, pos = [], [], 0
        for i, (rd, win) in enumerate(pairs):        # windows back to back behind gaps of 0..12 bases: every word phase
            gap = random_seq(rng, (7 * i) % 13)
            parts += [gap, win]
            woff.append(pos + len(gap))
            pos += len(gap) + len(win)
        codes = np.concatenate(parts)
        ctx.refseq_upload(pack3(codes), len(codes), np.array([0, len(codes)], np.uint64))
    nmulti = 0
    for sel in (slice(None), slice(3, 4)):
        idx = list(range(len(pairs)))[sel]
        sub_pairs = [pairs[i] for i in idx]
        sub_offs = np.concatenate([[offs[2 * i], offs[2 * i + 1]] for i in idx] + [[0]])
        t = _band_tasks(sub_pairs, sub_offs, [args[i] for i in idx], [minscore[i] for i in idx],
                        [minscorlen[i] for i in idx])
        if packed:
            for k, i in enumerate(idx):
                t[k]["ref_off"] = woff[i]
                t[k]["flags"] = SMB_TASK_REF_PACKED
        res, first, diff, errs, cells = ctx.band_align(t)
        ocells = 0
        for k, i in enumerate(idx):
            rd, win = pairs[i]
            e, want, c = orc.band_align(rd, win, *args[i], minscore[i], minscorlen[i])
            ocells += c
            assert int(errs[k]) == e, (i, args[i])
            assert _unpack(res, first, diff, k) == want, (i, len(rd), len(win), args[i], minscore[i], minscorlen[i])
            nmulti += len(want) > 1
        assert cells == ocells
    assert nmulti >= 1


@pytest.mark.parametrize("pen", [(2, -1, -1, -1), (1, -3, -9, -1), (3, -2, -5, -4), (1, -1, -2, 0)])
def test_dp_kernels_other_penalties(pen):
    """K2 and K3 under other penalty sets than the default (match, mismatch, gap open, gap extension):
    cheap gaps, where gap states reach far into the padding the packed K3 kernel no longer masks and
    the gap_init bias of the packed K2 kernel is small; expensive ones; a free extension."""
    import smalt_b200
    orc = Oracle(pen)
    ctx = smalt_b200.Context(0, penalties=pen)
    try:
        rng = np.random.default_rng(700 + pen[0] * 7 - pen[2])
        qmax = 255 // pen[0]          # packed kernels: scores < 256
        trip = _rand_pairs(rng, 240, 20, min(160, qmax))
        pairs, args = [], []
        for k, (a, b, lf) in enumerate(trip):
            if k % 7 == 0:
                b = np.concatenate([b, random_seq(rng, 9), b])
            pairs.append((a, b))
            args.append(_band_args(rng, len(a), len(b), lf))
        minscore = [int(x) for x in rng.integers(1, 40, len(pairs))]
        minscorlen = [int(x) for x in rng.integers(5, 30, len(pairs))]
        arena, offs = _arena(pairs)
        ctx.arena_upload(arena)
        scores, errs = ctx.sw_score(_sw_tasks(pairs, offs))
        for i, (rd, win) in enumerate(pairs):
            assert (int(errs[i]), int(scores[i])) == orc.sw_striped(rd, win), (i, pen)
        res, first, diff, errs, cells = ctx.band_align(_band_tasks(pairs, offs, args, minscore, minscorlen))
        ocells, most = 0, 0
        for i, (rd, win) in enumerate(pairs):
            e, want, c = orc.band_align(rd, win, *args[i], minscore[i], minscorlen[i])
            ocells += c
            most = max(most, len(want))
            assert int(errs[i]) == e, (i, pen, args[i])
            assert _unpack(res, first, diff, i) == want, (i, pen, args[i], minscore[i], minscorlen[i])
        # (a task with more results than the first attempt has slots for is run again with more slots:
        # the cell counter then includes the abandoned attempt)
        assert cells == ocells if most <= 4 else cells >= ocells
    finally:
        ctx.close()
