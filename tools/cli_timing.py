"""whole-program timing of `smalt_b200 map` with the stage timers on (SMALT_B200_TIMING)"""
import os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
threads = int(sys.argv[2]) if len(sys.argv) > 2 else bench.host_threads()
genome = bench.make_genome()
reads, pos, strand, span = bench.simulate_reads(genome, n, seed=43)
with tempfile.TemporaryDirectory() as tmp:
    pref, fq, ix = bench.write_workload_files(tmp, genome, reads)
    exe = os.path.join(bench.ROOT, "smalt_b200", "bin", "smalt_b200")
    for rep in range(2):
        t0 = time.time()
        r = subprocess.run([exe, "map", "-n", str(threads), "-O", "-o", os.path.join(tmp, "b.sam"), pref, fq], capture_output=True, text=True,
                           env=dict(os.environ, SMALT_B200_TIMING="1", SMALT_B200_STATS=os.path.join(tmp, "st.json"), **({"SMALT_B200_BLOCK": sys.argv[3]} if len(sys.argv) > 3 else {})))
        dt = time.time() - t0
        print("cli -n %d: rc %d %.2f s %.0f reads/s" % (threads, r.returncode, dt, n / dt))
        lines = [l for l in r.stderr.splitlines() if "timing" in l]
        print("\n".join(lines[:40]))
        print("...")
        print("\n".join(lines[-6:]))
        print(open(os.path.join(tmp, "st.json")).read())
