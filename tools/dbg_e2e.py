"""debug helper: reproduce an e2e case, isolate differing reads, show reference DP trace vs wave debug output"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import test_gpu_e2e_sam as T
from smalt_b200 import indexer
case = [c for c in T.CASES if c[0] == sys.argv[1]][0]
name, seed, lens, k, s, nreads, qlen, err, repeats, fasta, threads = case
tmp = "/tmp/dbg_e2e"; os.makedirs(tmp, exist_ok=True)
rng = np.random.default_rng(seed)
seqs = T._genome(rng, lens, repeats)
pref = tmp + "/idx"
indexer.write_smi(pref, indexer.build_index(seqs, k, s)); indexer.write_sma(pref, ["chr%d" % i for i in range(len(seqs))], seqs)
reads = T._reads(rng, seqs, nreads, qlen, err)
fq = tmp + "/reads.fq"; T._write_fastq(fq, reads, fasta)
REF = T.ref_binary("smalt")
def run(exe, out, fqp, env=None):
    r = subprocess.run([exe, "map", "-o", out, pref, fqp], capture_output=True, text=True, env=dict(os.environ, **(env or {})))
    return r
run(REF, tmp + "/ref.sam", fq); run(T.B200, tmp + "/b.sam", fq)
a, b = T._sam(tmp + "/ref.sam"), T._sam(tmp + "/b.sam")
bad = [x.split("\t")[0] for x, y in zip(a, b) if x != y]
print("differing reads:", bad)
for rn in bad[:2]:
    i = int(rn[1:])
    fq1 = tmp + "/one.fq"; T._write_fastq(fq1, [reads[i]], fasta)
    run(REF, tmp + "/ref1.sam", fq1); r = run(T.B200, tmp + "/b1.sam", fq1, {"SMALT_B200_DEBUG": "1"})
    print("REF :", [l for l in T._sam(tmp + "/ref1.sam") if not l.startswith("@")])
    print("B200:", [l for l in T._sam(tmp + "/b1.sam") if not l.startswith("@")])
    print(r.stderr[-3000:])
    tr = tmp + "/trace.txt"
    subprocess.run([T.ref_binary("smalt_trace"), "map", "-o", tmp + "/t.sam", pref, fq1], capture_output=True, env=dict(os.environ, SMALT_TRACE=tr))
    for ln in open(tr):
        f = ln.split()
        if f[0] == "SW": print("REF SW score", f[2], "qlen", f[3], "rlen", f[4])
        elif f[0] == "BA": print("REF BA band", f[2], f[3], "minscore", f[8], "minscorlen", f[9], "rlen", f[11], "nres", f[14], f[15:])
