"""Quick device-time probe of the three kernels on C2-shaped synthetic tasks (not the bench)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smalt_b200
from smalt_b200.capi import SW_TASK_DTYPE, BAND_TASK_DTYPE
from smalt_b200 import indexer
from smalt_b200.seqpack import pack3

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
rng = np.random.default_rng(0)
G = 5_000_000
genome = rng.integers(0, 4, G).astype(np.uint8)
qlen, wl = 150, 171
starts = rng.integers(0, G - wl - 1, n)
reads = np.stack([genome[s + 10:s + 10 + qlen] for s in starts])
mut = rng.random(reads.shape) < 0.02
reads[mut] = (reads[mut] + rng.integers(1, 4, int(mut.sum()))) & 3
ctx = smalt_b200.Context(0)
t0 = time.time(); ix = indexer.as_loaded(indexer.build_index([genome], 13, 6)); print("index build %.1fs" % (time.time() - t0))
ctx.index_upload(ix)
words = pack3(np.concatenate([genome, [7]]).astype(np.uint8))
ctx.refseq_upload(words, G + 1, np.array([0, G], np.uint64))
ctx.arena_upload(reads.reshape(-1))
sw = np.zeros(n, SW_TASK_DTYPE)
sw["read_off"] = np.arange(n, dtype=np.uint64) * qlen
sw["ref_off"] = starts; sw["read_len"] = qlen; sw["ref_len"] = wl; sw["flags"] = 2
for rep in range(3):
    sc, er = ctx.sw_score(sw)
    ms = ctx.last_kernel_ms
    print("K2: %d tasks %.2f ms  %.1f GCUPS  mean score %.1f" % (n, ms, n * qlen * wl / ms / 1e6, sc.mean()))
bt = np.zeros(n, BAND_TASK_DTYPE)
bt["read_off"] = sw["read_off"]; bt["ref_off"] = starts; bt["read_len"] = qlen; bt["ref_len"] = wl; bt["flags"] = 2
bt["l_edge"] = -10 - 9; bt["r_edge"] = -10 + 9; bt["p_left"] = 0; bt["p_right"] = qlen - 1
bt["u_left"] = 0; bt["u_right"] = wl - 1; bt["minscore"] = 50; bt["minscorlen"] = 30
for rep in range(3):
    t0 = time.time()
    res, first, diff, errs, cells = ctx.band_align(bt)
    ms = ctx.last_kernel_ms
    print("K3: %d tasks %.2f ms  %.2f GCUPS (cells %d) results %d  wall %.2fs" % (n, ms, cells / ms / 1e6, cells, len(res), time.time() - t0))
for rep in range(3):
    t0 = time.time()
    info, _ = ctx.seed_batch(sw["read_off"], sw["read_len"], full=False)
    ms = ctx.last_kernel_ms
    print("K1: %d reads %.2f ms  %.1f Mlookups/s  mean seeds %.1f rank %.1f  wall %.2fs" % (
        n, ms, n * 2 * (qlen - 12) / ms / 1e3, info["n_seeds"].mean(), info["seed_rank"].mean(), time.time() - t0))
