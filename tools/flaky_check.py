"""which side is not reproducible? runs reference and smalt_b200 N times on an e2e test workload"""
import os, subprocess, sys, hashlib, collections
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import test_gpu_e2e_sam as T
from smalt_b200 import indexer
import pathlib, tempfile
case = sys.argv[1] if len(sys.argv) > 1 else "fasta_noisy"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 12
name, seed, lens, k, s, nreads, qlen, err, repeats, fasta, threads = [c for c in T.CASES if c[0] == case][0]
tmp = pathlib.Path(tempfile.mkdtemp())
rng = np.random.default_rng(seed)
seqs = T._genome(rng, lens, repeats)
pref = str(tmp / "idx")
indexer.write_smi(pref, indexer.build_index(seqs, k, s))
indexer.write_sma(pref, ["chr%d" % i for i in range(len(seqs))], seqs)
reads = T._reads(rng, seqs, nreads, qlen, err)
fq = str(tmp / ("reads.fa" if fasta else "reads.fq"))
T._write_fastq(fq, reads, fasta)
res = {}
for tag, exe in (("ref", T.ref_binary("smalt")), ("b200", T.B200)):
    hs = collections.Counter()
    outs = {}
    for it in range(N):
        out = str(tmp / ("%s_%d.sam" % (tag, it)))
        env = dict(os.environ, SMALT_B200_BLOCK="1024")
        r = subprocess.run([exe, "map"] + sys.argv[3:] + ["-o", out, pref, fq], capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr[-500:]
        lines = [l for l in open(out).read().splitlines() if not l.startswith("@PG")]
        h = hashlib.md5("\n".join(lines).encode()).hexdigest()
        hs[h] += 1
        outs[h] = lines
    print(tag, dict(hs))
    res[tag] = outs
    if len(outs) > 1:
        a, b = list(outs.values())[:2]
        d = [(x, y) for x, y in zip(a, b) if x != y]
        print(" ", len(d), "lines differ between two runs of", tag)
        for x, y in d[:3]:
            print("   ", x[:150]); print("   ", y[:150])
