"""Pins the restatement of the candidate selection (oracle/smalt_oracle_cand.c: segLstFillHits,
segAliCandsAddFast, segAliCandsStats, segAliCandsCalcSegmentOffsets of segment.c) against the
UNMODIFIED reference (oracle/_ref/libsmalt_ref.so running its own segment.c on its own hit lists)."""
import numpy as np
import pytest

from oracle_lib import Oracle, RefLib, have_ref
from seqgen import mutate, random_seq, revcomp
from smalt_b200 import indexer

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")


def repeat_genome(rng, lens, nrep=12, unit_len=400, div=0.03):
    """sequences with diverged copies of a repeat unit (several candidates per read, tied covers)"""
    seqs = [random_seq(rng, n, p_n=0.0003) for n in lens]
    unit = random_seq(rng, unit_len)
    for s in seqs:
        for _ in range(nrep):
            p = int(rng.integers(0, len(s) - unit_len))
            s[p:p + unit_len] = mutate(rng, unit, p_sub=div, p_ins=0, p_del=0)[:unit_len]
    # tandem duplication: overlapping candidate segments with different shifts
    seqs[0][2000:2300] = seqs[0][1700:2000]
    return seqs


def sample(rng, seqs, qlen, err):
    s = seqs[int(rng.integers(0, len(seqs)))]
    st = int(rng.integers(0, len(s) - qlen - 40))
    rd = mutate(rng, s[st:st + qlen + 30].copy(), p_sub=err, p_ins=err / 3, p_del=err / 3)[:qlen]
    if rng.random() < 0.5:
        rd = revcomp(rd)
    return np.ascontiguousarray(rd)


CASES = [(13, 6, [60000, 45000], 150), (11, 3, [30011, 20007, 999], 100), (13, 2, [50000], 250),
         (7, 1, [3000, 2999], 36), (20, 13, [90000, 70000], 150)]


@pytest.fixture(scope="module", params=CASES)
def setup(request, tmp_path_factory):
    k, nskip, lens, qlen = request.param
    rng = np.random.default_rng(1000 + k * 10 + nskip)
    seqs = repeat_genome(rng, lens, unit_len=min(400, min(lens) // 4))
    ix = indexer.build_index(seqs, k, nskip)
    pref = str(tmp_path_factory.mktemp("ixc") / "g")
    indexer.write_smi(pref, ix)
    indexer.write_sma(pref, ["s%d" % i for i in range(len(seqs))], seqs)
    ref = RefLib()
    ref.index_load(pref)
    orc = Oracle()
    oix = orc.make_index(indexer.as_loaded(ix))
    # `smalt index` stores the sequences back to back, no terminator in between (sequence.c:2448-2519)
    soffs = np.concatenate([[0], np.cumsum([len(s) for s in seqs])]).astype(np.uint64)
    return dict(k=k, nskip=nskip, seqs=seqs, ref=ref, orc=orc, oix=oix, rng=rng, qlen=qlen, soffs=soffs)


@pytest.mark.parametrize("mode", [dict(), dict(best=True, min_swatscor_below_max=0), dict(min_swatscor_below_max=12),
                                  dict(target_depth=2, max_depth=5), dict(target_depth=2, sensitive=True),
                                  dict(min_cover=60), dict(nhit_max=40, maxhit_total=400)])
def test_candidates_vs_reference(setup, mode):
    s = setup
    rng = s["rng"]
    nmulti = 0
    for it in range(120):
        rd = sample(rng, s["seqs"], s["qlen"], err=[0.0, 0.02, 0.06][it % 3])
        if it % 17 == 5:
            rd = random_seq(rng, s["qlen"])   # maps nowhere
        er, st_r, cd_r = s["ref"].candidates(rd, **mode)
        eo, st_o, cd_o = s["orc"].candidates(s["oix"], rd, s["soffs"], termchar=0, **mode)
        assert er == eo, (it, er, eo)
        if er:
            continue
        for key in ("n_sort", "n_mincover", "max_cover", "max2nd_cover", "cover_deficit", "nhit", "nhit_tot"):
            assert st_r[key] == st_o[key], (it, key, st_r, st_o)
        assert cd_r == cd_o, (it, cd_r, cd_o)
        nmulti += len(cd_r) > 1
    assert nmulti > 5 or mode.get("min_cover")   # the repeats do produce candidate lists with several entries
