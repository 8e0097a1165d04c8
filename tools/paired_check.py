"""Paired-end / fiber-path parity and timing check against the reference (needs a GPU):
python tools/paired_check.py [npairs] [genome_mb] [nseq]"""
import json, os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from seqgen import mutate, random_seq, revcomp
from smalt_b200 import indexer
REF = os.path.join(ROOT, "oracle", "_ref", "smalt")
B200 = os.path.join(ROOT, "smalt_b200", "bin", "smalt_b200")
LET = np.frombuffer(b"ACGTNN", np.uint8)


def make(tmp, npairs, gmb, nseq, seed=5, repeats=True, k=13, s=6):
    rng = np.random.default_rng(seed)
    g = [random_seq(rng, int(gmb * 1e6 / nseq)) for _ in range(nseq)]
    if repeats:
        unit = random_seq(rng, 700)
        for q in g:
            for _ in range(max(2, len(q) // 50000)):
                p = int(rng.integers(0, len(q) - 700))
                q[p:p + 700] = mutate(rng, unit, p_sub=0.01, p_ins=0, p_del=0)[:700]
    pref = os.path.join(tmp, "idx")
    indexer.write_smi(pref, indexer.build_index(g, k, s))
    indexer.write_sma(pref, ["chr%d" % i for i in range(len(g))], g)
    f1, f2 = os.path.join(tmp, "r1.fq"), os.path.join(tmp, "r2.fq")
    with open(f1, "w") as a, open(f2, "w") as b:
        for i in range(npairs):
            q = g[int(rng.integers(0, nseq))]
            ins = max(160, int(rng.normal(400, 40)))
            if ins + 2 >= len(q):
                ins = len(q) - 2
            st = int(rng.integers(0, len(q) - ins - 1))
            frag = q[st:st + ins]
            x = mutate(rng, frag[:150].copy(), p_sub=0.02, p_ins=0.002, p_del=0.002)
            y = revcomp(mutate(rng, frag[-150:].copy(), p_sub=0.02, p_ins=0.002, p_del=0.002))
            if i % 37 == 0:
                y = random_seq(rng, 150)
            if i % 41 == 0:
                x = random_seq(rng, 150)
            if i % 2:
                x, y = y, x
            for f, r, t in ((a, x, 1), (b, y, 2)):
                sq = LET[r].tobytes().decode()
                f.write("@p%d/%d\n%s\n+\n%s\n" % (i, t, sq, "I" * len(sq)))
    return pref, f1, f2


def run(exe, args, out, env=None):
    t0 = time.time()
    r = subprocess.run([exe, "map", "-r", "7", "-o", out] + args, capture_output=True, text=True,
                       env=dict(os.environ, **(env or {})), timeout=3000)
    dt = time.time() - t0
    if r.returncode:
        print(r.stderr[-3000:])
        raise SystemExit("%s failed" % exe)
    return [l for l in open(out).read().splitlines() if not l.startswith("@PG")], dt


def main():
    npairs = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    gmb = float(sys.argv[2]) if len(sys.argv) > 2 else 4.0
    nseq = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    extra = sys.argv[4:]
    tmp = os.environ.get("TMPDIR", "/tmp") + "/paired_check"
    os.makedirs(tmp, exist_ok=True)
    pref, f1, f2 = make(tmp, npairs, gmb, nseq)
    args = extra + ["-i", "600", "-j", "200", pref, f1, f2]
    ncpu = os.cpu_count()
    ref, t_ref = run(REF, ["-n", "1"] + args, tmp + "/ref.sam")
    stats = tmp + "/stats.json"
    got, t_b = run(B200, ["-n", "1"] + args, tmp + "/b200.sam", {"SMALT_B200_STATS": stats})
    diff = [(x, y) for x, y in zip(ref, got) if x != y]
    print("pairs %d  ref -n 1: %.2f s   b200 -n 1: %.2f s   lines %d/%d  differing %d" %
          (npairs, t_ref, t_b, len(ref), len(got), len(diff)))
    print(open(stats).read())
    for x, y in diff[:3]:
        print("REF ", x[:300]); print("B200", y[:300])
    refn, t_refn = run(REF, ["-n", str(ncpu), "-O"] + args, tmp + "/refn.sam")
    gotn, t_bn = run(B200, ["-n", str(ncpu), "-O"] + args, tmp + "/b200n.sam", {"SMALT_B200_STATS": stats})
    dn = [(x, y) for x, y in zip(refn, gotn) if x != y and (x.startswith("@") or int(x.split("\t")[4]) > 6 or int(y.split("\t")[4]) > 6)]
    print("-n %d: ref %.2f s (%.0f reads/s)  b200 %.2f s (%.0f reads/s)  differing (MAPQ>6) %d" %
          (ncpu, t_refn, 2 * npairs / t_refn, t_bn, 2 * npairs / t_bn, len(dn)))
    print(open(stats).read())
    for x, y in dn[:3]:
        print("REFn ", x[:200]); print("B200n", y[:200])
    return 1 if diff or dn or len(ref) != len(got) else 0


if __name__ == "__main__":
    sys.exit(main())
