"""whole-program timing of `smalt_b200 map` on the C2 workload: device candidate path vs SMALT_B200_HOSTCAND"""
import os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
cfg = bench.CONFIGS["c2"]
with tempfile.TemporaryDirectory() as tmp:
    wl = bench.Workload(tmp, cfg, n)
    for mode in ({}, {"SMALT_B200_HOSTCAND": "1"}, {}):
        t0 = time.time()
        r = subprocess.run([bench.SMALT_B200, "map", "-n", "16", "-O", "-o", os.path.join(tmp, "b.sam"), wl.pref] + wl.files,
                           capture_output=True, text=True, env=dict(os.environ, SMALT_B200_TIMING="1", **mode))
        dt = time.time() - t0
        print("cli %s: rc %d %.2f s %.0f reads/s" % (mode, r.returncode, dt, n / dt))
        lines = [l for l in r.stderr.splitlines() if "timing" in l]
        print("\n".join(lines[:12]))
        print("...")
        print("\n".join(lines[-8:]))
