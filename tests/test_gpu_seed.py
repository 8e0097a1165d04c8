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
