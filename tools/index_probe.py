"""Times smb_index_build (csrc/index_build.cu) stage by stage on a random genome:
    python tools/index_probe.py [Mb] [k] [nskip]
SMB_INDEX_DEBUG=1 makes the library print the stage times; the second build is the warm one.  The table is
checked with smalt_b200/indexcheck.py (structure of every array, sampled grid positions found under their words)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SMB_INDEX_DEBUG", "1")
import smalt_b200  # noqa: E402
from smalt_b200 import indexer  # noqa: E402

mb = float(sys.argv[1]) if len(sys.argv) > 1 else 100
k = int(sys.argv[2]) if len(sys.argv) > 2 else 13
nskip = int(sys.argv[3]) if len(sys.argv) > 3 else 6
rng = np.random.default_rng(3)
n = int(mb * 1e6)
seqs = [rng.integers(0, 4, n // 4, dtype=np.uint8) for _ in range(4)]
ctx = smalt_b200.Context(0)
for it in range(2):
    t = time.time()
    ix = indexer.build_index_gpu(ctx, seqs, k, nskip)
    print("build %d: %.3f s  npos %d nwords %d" % (it, time.time() - t, ix["npos"], ix["nwords"]), flush=True)
from smalt_b200 import indexcheck  # noqa: E402
t = time.time()
indexcheck.check_structure(ix)
n = indexcheck.check_samples(ix, seqs, k, nskip, nsample=20000)
print("structure ok, %d sampled grid positions found under their words (%.1f s)" % (n, time.time() - t))
