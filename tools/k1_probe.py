"""K1 on an index that does not fit L2: device times of the seed tables, the hit lists and the candidate
selection for one batch of simulated reads, per stage (CUDA events of the C ABI).
usage: python tools/k1_probe.py <config c2|c3|c4> [reads] [genome scale]"""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from smalt_b200 import indexer
from smalt_b200.capi import Context, BLOCK_JOB_DTYPE, HIT_REQ_DTYPE
from smalt_b200.seqpack import pack3

cfgname = sys.argv[1] if len(sys.argv) > 1 else "c4"
nreads = int(sys.argv[2]) if len(sys.argv) > 2 else 32000
scale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
cfg = dict(bench.CONFIGS[cfgname])
cfg["seqs"] = [max(1000, int(n * scale)) for n in cfg["seqs"]]
t0 = time.time()
seqs = bench.make_genome(cfg)
words = pack3(np.concatenate(list(seqs) + [np.array([7], np.uint8)]))
ctx = Context(0)
ix = indexer.build_index_gpu(ctx, seqs, cfg["k"], cfg["s"], words=words)
ctx.index_upload(indexer.as_loaded(ix))
print("index: %.1f s, npos %d nwords %d nkeys %d" % (time.time() - t0, ix["npos"], ix["nwords"], ix["nkeys"]), flush=True)
# reads: substrings of the genome with 2 % substitutions, random strand (the seed path does not care about pairing)
rng = np.random.default_rng(5)
qlen = cfg["qlen"] or 150
reads = np.zeros((nreads, qlen), np.uint8)
for i in range(nreads):
    s = seqs[int(rng.integers(0, len(seqs)))]
    st = int(rng.integers(0, len(s) - qlen))
    reads[i] = s[st:st + qlen]
sub = rng.random((nreads, qlen)) < 0.02
reads[sub] = (reads[sub] + 1 + rng.integers(0, 3, int(sub.sum()))) & 3
rc = rng.integers(0, 2, nreads).astype(bool)
reads[rc] = 3 - reads[rc][:, ::-1]
arena = reads.reshape(-1)
offs = np.arange(nreads, dtype=np.uint64) * qlen
lens = np.full(nreads, qlen, np.uint32)
ctx.arena_upload(arena)
out = {"config": cfgname, "reads": nreads, "index": {k: int(ix[k]) for k in ("npos", "nwords", "nkeys")}}
for rep in range(3):
    info, _ = ctx.seed_batch(offs, lens, None, 10000, 16384, 0, full=False)
    out["seed_ms"] = ctx.last_kernel_ms
nseq = len(seqs)
soffs = np.concatenate([[0], np.cumsum([len(s) for s in seqs])]).astype(np.uint64)
req = np.zeros(nreads * 2 * nseq, HIT_REQ_DTYPE)
req["read"] = np.repeat(np.arange(nreads, dtype=np.uint32), 2 * nseq)
req["strand"] = np.tile(np.repeat(np.array([0, 1], np.uint8), nseq), nreads)
req["lo"] = np.tile(np.tile(soffs[:-1], 2), nreads)
req["hi"] = np.tile(np.tile(soffs[1:], 2), nreads)
req["nhit_max"], req["use_short"] = 10000, 1
for rep in range(2):
    sq, first, errs = ctx.hits_batch(req)
    out["hits_ms"] = ctx.last_kernel_ms
out["hits"] = int(len(sq))
out["requests"] = int(len(req))
out["nonempty_requests"] = int((np.diff(first.astype(np.int64)) > 0).sum())
jobs = np.zeros(nreads, BLOCK_JOB_DTYPE)
jobs["seed_read"] = np.arange(nreads)
jobs["niv"] = -1
jobs["min_swatscor"] = 20
for rep in range(2):
    sz = ctx.block_run(jobs, None, 10000, 0, 200, 8000, True, False, False)
out["block"] = {k: float(sz[k]) for k in ("ms_hits", "ms_cand", "ms_k2", "ms_k3", "nhits", "ncand", "nk3")}
# algorithmic bytes of the lookups
bucket = ix["nwords"] / ix["nkeys"] if ix["typ"] else 1.0
probes = float(np.ceil(np.log2(bucket + 1))) if ix["typ"] else 0.0
nlook = 2 * nreads * (qlen - cfg["k"] + 1)
out["lookup_bytes"] = 8 + 4 * probes + 8
out["seed_gbs"] = nlook * out["lookup_bytes"] / (out["seed_ms"] * 1e-3) / 1e9
print(json.dumps(out))
ctx.close()
