"""times the smalt_b200 driver against the reference smalt on a C2-shaped workload"""
import os, subprocess, sys, tempfile, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
threads = int(sys.argv[2]) if len(sys.argv) > 2 else bench.host_threads()
block = sys.argv[3] if len(sys.argv) > 3 else "8192"
do_ref = len(sys.argv) <= 4 or sys.argv[4] != "noref"
genome = bench.make_genome()
reads, pos, strand, span = bench.simulate_reads(genome, n, seed=43)
with tempfile.TemporaryDirectory() as tmp:
    pref, fq, ix = bench.write_workload_files(tmp, genome, reads)
    exe = os.path.join(bench.ROOT, "smalt_b200", "bin", "smalt_b200")
    ref = os.path.join(bench.ROOT, "oracle", "_ref", "smalt")
    for tag, e in (("b200", exe), ("ref", ref)):
        if tag == "ref" and not do_ref:
            continue
        out = os.path.join(tmp, tag + ".sam")
        env = dict(os.environ, SMALT_B200_BLOCK=block, SMALT_B200_STATS=os.path.join(tmp, "stats.json"), SMALT_B200_TIMING=os.environ.get("TIMING", ""))
        t0 = time.time()
        r = subprocess.run([e, "map", "-n", str(threads), "-O", "-o", out, pref, fq], capture_output=True, text=True, env=env)
        dt = time.time() - t0
        print(tag, "rc", r.returncode, "%.2f s  %.0f reads/s" % (dt, n / dt), r.stderr[-300:] if r.returncode else "")
        if tag == "b200" and os.environ.get("TIMING"): print("\n".join(l for l in r.stderr.splitlines() if "timing" in l or "Time" in l)[:3000])
        if tag == "b200" and os.path.exists(os.path.join(tmp, "stats.json")):
            print(open(os.path.join(tmp, "stats.json")).read())
    if do_ref:
        a = [l for l in open(os.path.join(tmp, "b200.sam")) if not l.startswith("@PG")]
        b = [l for l in open(os.path.join(tmp, "ref.sam")) if not l.startswith("@PG")]
        d = [(x, y) for x, y in zip(a, b) if x != y]
        d6 = [(x, y) for x, y in d if int(x.split("\t")[4]) > 6 or int(y.split("\t")[4]) > 6]
        print("SAM lines", len(a), len(b), "differing", len(d), "differing with MAPQ>6", len(d6))
