"""`smalt_b200 index` (GPU) against the reference's `smalt index` on a C3-sized genome:
python tools/index_timing.py [megabases] [nseq] [k] [s]"""
import os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
mb = float(sys.argv[1]) if len(sys.argv) > 1 else 100.0
nseq = int(sys.argv[2]) if len(sys.argv) > 2 else 4
k = sys.argv[3] if len(sys.argv) > 3 else "13"
s = sys.argv[4] if len(sys.argv) > 4 else "6"
tmp = "/tmp/idxt"; os.makedirs(tmp, exist_ok=True)
rng = np.random.default_rng(9)
let = np.frombuffer(b"ACGT", np.uint8)
fa = tmp + "/g.fa"
with open(fa, "wb") as f:
    for i in range(nseq):
        n = int(mb * 1e6 / nseq)
        t = let[rng.integers(0, 4, n)]
        pad = (-n) % 60
        rows = np.concatenate([t, np.zeros(pad, np.uint8)]).reshape(-1, 60)
        rows = np.concatenate([rows, np.full((len(rows), 1), 10, np.uint8)], axis=1).reshape(-1)
        f.write(b">chr%d\n" % i + rows[rows != 0].tobytes())
out = {}
for tag, exe in (("b200", ROOT + "/smalt_b200/bin/smalt_b200"), ("ref", ROOT + "/oracle/_ref/smalt")):
    t0 = time.time()
    r = subprocess.run([exe, "index", "-k", k, "-s", s, tmp + "/" + tag, fa], capture_output=True, text=True,
                       env=dict(os.environ, SMB_INDEX_DEBUG="1"))
    out[tag] = time.time() - t0
    print(tag, "%.2f s" % out[tag], r.returncode, "\n".join(l for l in r.stderr.splitlines() if "GPU" in l or "elapsed" in l or "index_build" in l or "smalt_b200 timing" in l))
same = all(open(tmp + "/b200" + e, "rb").read() == open(tmp + "/ref" + e, "rb").read() for e in (".smi", ".sma"))
print("genome %.0f Mb x %d sequences, k=%s s=%s: identical files: %s, speed-up of the whole program %.1fx" % (mb, nseq, k, s, same, out["ref"] / out["b200"]))
