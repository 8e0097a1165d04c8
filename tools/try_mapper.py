"""quick GPU check of the in-process mapper against the reference binary"""
import os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from smalt_b200.mapper import Mapper
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
threads = int(sys.argv[2]) if len(sys.argv) > 2 else bench.host_threads()
genome = bench.make_genome()
reads, pos, strand, span = bench.simulate_reads(genome, n, seed=43)
with tempfile.TemporaryDirectory() as tmp:
    pref, fq, ix = bench.write_workload_files(tmp, genome, reads)
    text = open(fq, "rb").read()
    t0 = time.time(); m = Mapper(pref, threads); print("open %.2f s" % (time.time() - t0))
    for it in range(4):
        t0 = time.time(); sam = m.map_fastq(text); dt = time.time() - t0
        st = m.stats.as_dict()
        print("map %d reads: %.3f s  %.0f reads/s  kernels %.0f ms  stages %s" % (n, dt, n / dt, st["k1_ms"] + st["k2_ms"] + st["k3_ms"],
              {k: round(v, 3) for k, v in st["host_stage_s"].items()}))
    m.close()
    ref = os.path.join(bench.ROOT, "oracle", "_ref", "smalt")
    out = os.path.join(tmp, "ref.sam")
    t0 = time.time(); subprocess.run([ref, "map", "-n", str(threads), "-O", "-o", out, pref, fq], check=True, capture_output=True); print("ref %.2f s" % (time.time() - t0))
    a = sam.decode().splitlines()
    b = [l for l in open(out).read().splitlines() if not l.startswith("@")]
    d = [(x, y) for x, y in zip(a, b) if x != y]
    d6 = [(x, y) for x, y in d if int(x.split("\t")[4]) > 6 or int(y.split("\t")[4]) > 6]
    print("SAM lines", len(a), len(b), "differing", len(d), "differing with MAPQ>6", len(d6))
    exe = os.path.join(bench.ROOT, "smalt_b200", "bin", "smalt_b200")
    for t in (threads,):
        t0 = time.time(); r = subprocess.run([exe, "map", "-n", str(t), "-O", "-o", os.path.join(tmp, "b.sam"), pref, fq], capture_output=True, text=True,
                                             env=dict(os.environ, SMALT_B200_STATS=os.path.join(tmp, "st.json"))); dt = time.time() - t0
        print("cli -n %d: rc %d %.2f s %.0f reads/s" % (t, r.returncode, dt, n / dt), r.stderr[-300:] if r.returncode else "")
        print(open(os.path.join(tmp, "st.json")).read())
